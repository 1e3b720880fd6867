"""Oracle (test infrastructure only) for SURVEY.md section 8 rows a9 (region level-set
energy), a10 (length regulariser), a11 (local consistency module), a16 (DiscoBox
mean-field CRF) and the DiscoBox dice/MIL form of a6.  Plain torch on CPU; float32 or
float64 according to the inputs.  Never imported by the product.
"""
import torch
import torch.nn.functional as F

from .boxinst import _shifted


# ----------------------------------------------------------------------------------------
# a9 / a10
# ----------------------------------------------------------------------------------------
def region_levelset(scores2, target):
    """Chan-Vese region energy per instance, [n].

    scores2 [n,2,h,w] (inside / outside memberships), target [n,C,h,w].
    c_k = <S_k,T>/max(sum S_k, 1e-5);  E = sum_{c,p} (T-c_0)^2 S_0 + (T-c_1)^2 S_1, / C.
    mmdet/models/losses/levelset_loss.py:21-44.
    """
    n, c = target.shape[:2]
    energy = target.new_zeros(n)
    for k in range(2):
        s = scores2[:, k:k + 1]
        mean_k = (s * target).sum((2, 3)) / s.sum((2, 3)).clamp(min=1e-5)      # [n,C]
        dev = target - mean_k[:, :, None, None]
        energy = energy + (dev * dev * s).sum((1, 2, 3))
    return energy / c


def levelset_loss(scores2, target, pixel_num, loss_weight=1.0):
    """levelset_loss.py:13-18."""
    return loss_weight * region_levelset(scores2, target) / pixel_num


def length_regularization(scores):
    """Total variation sum|dy|+sum|dx| per instance.  levelset_loss.py:47-60 (unused by
    the heads; kept for API completeness)."""
    gy = (scores[:, :, 1:, :] - scores[:, :, :-1, :]).abs().sum((1, 2, 3))
    gx = (scores[:, :, :, 1:] - scores[:, :, :, :-1]).abs().sum((1, 2, 3))
    return gy + gx


# ----------------------------------------------------------------------------------------
# a11
# ----------------------------------------------------------------------------------------
_LCM_TAPS = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]


def _replicate_neighbours(x, dilation):
    """[b,c,h,w] -> [b,c,8,h,w]; neighbour k = x at (y+dy*d, x+dx*d), edge-replicated."""
    h, w = x.shape[-2:]
    ys = torch.arange(h)
    xs = torch.arange(w)
    outs = []
    for dy, dx in _LCM_TAPS:
        yy = (ys + dy * dilation).clamp(0, h - 1)
        xx = (xs + dx * dilation).clamp(0, w - 1)
        outs.append(x[:, :, yy][:, :, :, xx])
    return torch.stack(outs, dim=2)


def lcm_affinity(imgs, dilation=2, alpha=0.3):
    """[b,1,8,h,w] softmax affinity.  levelset_loss.py:108-118."""
    nb = _replicate_neighbours(imgs, dilation)
    dev = (nb - imgs[:, :, None]).abs()
    std = nb.std(dim=2, keepdim=True)               # unbiased over the 8 neighbours
    a = -((dev / (std + 1e-8) / alpha) ** 2)
    a = a.mean(dim=1, keepdim=True)
    return torch.softmax(a, dim=2)


def lcm_refine(imgs, phis, num_iter=10, dilation=2):
    aff = lcm_affinity(imgs, dilation)
    for _ in range(num_iter):
        phis = (_replicate_neighbours(phis, dilation) * aff).sum(2)
    return phis


def lcm_loss(imgs, phis, box_targets, num_iter=10, dilation=2):
    """sum|phi_T - phi_0| * box / max(sum box, 1).  levelset_loss.py:64-71."""
    ref = lcm_refine(imgs, phis, num_iter, dilation)
    return ((ref - phis).abs() * box_targets).sum() / box_targets.sum().clamp(min=1)


# ----------------------------------------------------------------------------------------
# DiscoBox: dice / MIL (a6 variant) and mean field (a16)
# ----------------------------------------------------------------------------------------
def disco_dice_loss(x, t):
    """1 - 2a/((b+1e-3)+(c+1e-3)).  mmdet/models/dense_heads/discobox_head.py:542-550."""
    x = x.flatten(1).float()
    t = t.flatten(1).float()
    return 1 - 2 * (x * t).sum(1) / ((x * x).sum(1) + 0.001 + (t * t).sum(1) + 0.001)


def disco_mil_loss(x, t):
    """x,t [n,h,w]: dice on the column profile + dice on the row profile.
    discobox_head.py:552-562."""
    return disco_dice_loss(x.max(2)[0], t.max(2)[0]) + disco_dice_loss(x.max(1)[0], t.max(1)[0])


def meanfield_kernel(feature_map, kernel_size=3, theta0=0.5, theta1=30.0, alpha0=3.0):
    """[b, k*k, h*w] bilateral kernel incl. the centre tap.  discobox_head.py:590-610.

    The +10 offset makes zero-padded border taps dissimilar.
    """
    b, c, h, w = feature_map.shape
    fm = feature_map + 10
    r = kernel_size // 2
    taps = []
    for j in range(kernel_size * kernel_size):
        dy, dx = j // kernel_size - r, j % kernel_size - r
        d = _shifted(fm, dy, dx, 0.0) - fm
        app = -(d * d).sum(1) / (2 * theta0 ** 2)
        spa = -float(dy * dy + dx * dx) / (2 * theta1 ** 2)
        taps.append(alpha0 * torch.exp(app + spa))
    return torch.stack(taps, 1).flatten(2)


def meanfield_forward(kernel, x, targets, kernel_size=3, num_iter=20, base=0.45, inter=None, gamma=0.01):
    """x, targets [n,1,h,w]; kernel [1,k*k,h*w] -> (binary pseudo label [n,1,h,w], valid [n]).

    discobox_head.py:616-651; ``inter`` [n,2,h,w] is corr_loss's inter_img_mask (background, foreground): added to the
    two potentials, times gamma, before the target product (:643-644).
    """
    n, _, h, w = x.shape
    r = kernel_size // 2
    q = ((x * targets) > 0.5).to(x.dtype) * (1 - 2 * base) + base
    u = torch.cat([1 - q, q], 1)                                   # [n,2,h,w]
    kern = kernel.view(kernel.shape[0], 1, kernel_size * kernel_size, h, w)
    for _ in range(num_iter):
        e = -torch.log(u)
        agg = torch.zeros_like(u)
        for j in range(kernel_size * kernel_size):
            dy, dx = j // kernel_size - r, j % kernel_size - r
            agg = agg + _shifted(e, dy, dx, 0.0) * kern[:, :, j]
        f = torch.exp(-agg)
        if inter is not None:
            f = f + inter * gamma
        f = torch.cat([f[:, :1], f[:, 1:] * targets], 1) + 1e-6
        f = f / f.sum(1, keepdim=True)
        u = (f > 0.5).to(x.dtype) * (1 - 2 * base) + base
    ret = (u[:, 1:] > 0.5).to(x.dtype)
    cnt = ret.flatten(1).sum(1)
    valid = ((cnt >= h * w * 0.05) & (cnt <= h * w * 0.95)).to(x.dtype)
    return ret, valid

"""GPU parity of the glue kernels (rows a14, a17, a18): bilinear resize against ATen's own kernel (same arithmetic:
expected bit-level agreement up to FMA contraction), the fused tree edge weights against the reference's torch
formulation, and the one-launch level-set assembly against the oracle composition that follows the heads' lines.
Tolerance 1e-3 relative is the bar; the assertions are tighter where the arithmetic is the same."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


# ------------------------------------------------------------------ a18
@pytest.mark.parametrize('shape,size,align', [((3, 1, 256, 256), (96, 96), False), ((2, 2, 96, 96), (256, 256), False),
                                              ((1, 3, 1024, 1024), (256, 256), False), ((5, 1, 100, 128), (200, 256), False),
                                              ((2, 3, 37, 53), (41, 17), False), ((2, 3, 200, 256), (100, 128), True),
                                              ((1, 1, 7, 9), (1, 1), False), ((1, 2, 1, 1), (5, 4), True),
                                              ((4, 1, 50, 64), (200, 256), False)])
def test_bilinear_resize_matches_aten(shape, size, align):
    from boxinstseg_b200.ops.resize import bilinear_resize
    gen = torch.Generator().manual_seed(shape[2] * 7 + size[0])
    x = torch.randn(shape, generator=gen).to(DEV).requires_grad_(True)
    g = torch.randn(shape[0], shape[1], *size, generator=gen).to(DEV)
    ref = F.interpolate(x, size=size, mode='bilinear', align_corners=align)
    out = bilinear_resize(x, size, align_corners=align)
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)
    (gref,) = torch.autograd.grad((ref * g).sum(), x)
    (gx,) = torch.autograd.grad((out * g).sum(), x)
    assert torch.allclose(gx, gref, rtol=1e-5, atol=1e-5) and rel_err(gx, gref) <= 1e-6
    (gx2,) = torch.autograd.grad((bilinear_resize(x, size, align_corners=align) * g).sum(), x)
    assert torch.equal(gx, gx2)                                     # deterministic (ATen's backward is an atomic scatter)


def test_scale_target_is_the_reference_helper():
    from boxinstseg_b200.ops.resize import scale_target
    t = torch.rand(4, 64, 80, device=DEV)
    ref = F.interpolate(t.unsqueeze(1), size=(96, 96), mode='bilinear', align_corners=False)     # utils/misc.py:75-86
    assert torch.allclose(scale_target(t), ref, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ a14
@pytest.mark.parametrize('B,C,h,w,groups,low', [(2, 3, 24, 30, 1, True), (3, 5, 17, 9, 1, False), (1, 5, 200, 256, 1, False),
                                                (2, 4, 20, 12, 2, False), (2, 3, 96, 96, 1, True)])
def test_tree_edge_weight_matches_torch_formulation(B, C, h, w, groups, low):
    from boxinstseg_b200.ops.tree_filter import MinimumSpanningTree, TreeFilter2D, bfs
    gen = torch.Generator().manual_seed(h + w)
    guide = torch.randn(B, 3, h, w, generator=gen).to(DEV)
    embed = (torch.randn(B, C, h, w, generator=gen) * (0.05 if low else 0.4)).to(DEV)
    tree = MinimumSpanningTree(TreeFilter2D.norm2_distance)(guide)
    idx, par, chd = bfs(tree, 4)
    tf = TreeFilter2D(groups=groups)
    e1 = embed.clone().requires_grad_(True)
    e2 = embed.clone().requires_grad_(True)
    w_fused = tf.build_edge_weight(e1, idx, par, low, chd)                    # kernels
    w_ref = tf.build_edge_weight(e2, idx, par, low)                           # tree_filter.py:91-108 in torch
    assert w_fused.shape == w_ref.shape
    assert torch.allclose(w_fused, w_ref, rtol=1e-5, atol=1e-7)
    g = torch.randn(w_ref.shape, generator=gen).to(DEV)
    (g1,) = torch.autograd.grad((w_fused * g).sum(), e1)
    (g2,) = torch.autograd.grad((w_ref * g).sum(), e2)
    assert rel_err(g1, g2) <= 1e-5
    (g3,) = torch.autograd.grad((tf.build_edge_weight(e1, idx, par, low, chd) * g).sum(), e1)
    assert torch.equal(g1, g3)


# ------------------------------------------------------------------ a17
def _boxes(n, h, w):
    m = torch.zeros(n, h, w)
    for i in range(n):
        m[i, h // 6 + i: h // 6 + i + h // 2, w // 5: w // 5 + w // 2 + i] = 1
    return m


@pytest.mark.parametrize('n,C,h,w', [(4, 3, 24, 32), (16, 3, 200, 256), (8, 2, 256, 256), (3, 5, 50, 64), (2, 1, 9, 7)])
def test_levelset_assembly_vs_oracle(n, C, h, w):
    from boxinstseg_b200.models.losses import LevelsetLoss, levelset_assembly
    from oracle import levelset as ol
    gen = torch.Generator().manual_seed(n * h)
    logits = torch.randn(n, h, w, generator=gen) * 2
    box = _boxes(n, h, w)
    if n > 2:
        box[1] = 0                                                           # empty box: clamp(sum box, 1) and the eps clamps are active
    T = torch.randn(n, C, h, w, generator=gen)
    gl = torch.rand(n, generator=gen) + 0.5
    # oracle composition, float64, following box2mask_head.py:305-312
    x64 = logits.double().requires_grad_(True)
    t64 = T.double().requires_grad_(True)
    s = torch.sigmoid(x64.unsqueeze(1))
    b = box.double().unsqueeze(1)
    phi = torch.cat((s, 1 - s), 1) * b
    pix = b.sum((1, 2, 3)).clamp(min=1)
    ref = ol.levelset_loss(phi, t64 * b, pix) * 1.5
    gx_ref, gt_ref = torch.autograd.grad((ref * gl.double()).sum(), [x64, t64])
    # fused assembly
    xg = logits.to(DEV).requires_grad_(True)
    tg = T.to(DEV).requires_grad_(True)
    out = levelset_assembly(xg, box.to(DEV), tg, loss_weight=1.5)
    gx, gt = torch.autograd.grad((out * gl.to(DEV)).sum(), [xg, tg])
    assert torch.allclose(out.cpu().double(), ref.detach(), rtol=1e-4, atol=1e-7)
    assert rel_err(gx.cpu(), gx_ref) <= 1e-4 and rel_err(gt.cpu(), gt_ref) <= 1e-4
    # the dense LevelsetLoss signature (mode 0) through the same kernels
    phi_g = (torch.cat((torch.sigmoid(xg.unsqueeze(1)), 1 - torch.sigmoid(xg.unsqueeze(1))), 1) * box.to(DEV).unsqueeze(1)).detach().requires_grad_(True)
    tm = (tg * box.to(DEV).unsqueeze(1)).detach().requires_grad_(True)
    dense = LevelsetLoss(loss_weight=1.5)(phi_g, tm, pix.float().to(DEV))
    assert torch.allclose(dense.cpu().double(), ref.detach(), rtol=1e-4, atol=1e-7)
    gphi, gtm = torch.autograd.grad((dense * gl.to(DEV)).sum(), [phi_g, tm])
    phi64 = phi.detach().requires_grad_(True)
    tm64 = (t64 * b).detach().requires_grad_(True)
    gphi_ref, gtm_ref = torch.autograd.grad((ol.levelset_loss(phi64, tm64, pix) * 1.5 * gl.double()).sum(), [phi64, tm64])
    assert rel_err(gphi.cpu(), gphi_ref) <= 1e-4 and rel_err(gtm.cpu(), gtm_ref) <= 1e-4
    # determinism
    out2 = levelset_assembly(xg, box.to(DEV), tg, loss_weight=1.5)
    assert torch.equal(out, out2)

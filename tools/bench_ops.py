"""Secondary-op timings on the GPU box (configs C/D/E shapes of SURVEY section 8d): dynamic conv (tcgen05),
CondInst head, tree filter, level set, LCM, mean field, targets.  Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
dev = torch.device('cuda:0')
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0


def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


res = {}
g = torch.Generator(device=dev).manual_seed(0)
from boxinstseg_b200.ops.dynconv import dynconv1x1
for name, (B, C, h, w, I) in dict(box2mask_E=(1, 256, 256, 256, 100), discobox_C=(1, 256, 200, 256, 16),
                                  boxsolo_D_allgrid=(2, 256, 200, 256, 1600)).items():
    feats = [torch.randn(B, C, h, w, device=dev, generator=g) for _ in range(4)]       # 4 x 67 MB rotation > L2
    kern = torch.randn(B, I, C, device=dev, generator=g) * 0.05
    it = [0]
    def run():
        it[0] += 1
        return dynconv1x1(feats[it[0] % 4], kern)
    us = timeit(run)
    byts = B * (C * h * w * 4 + I * h * w * 4 + I * C * 4)
    flops = 2.0 * B * I * C * h * w
    ref_us = timeit(lambda: torch.einsum('bic,bchw->bihw', kern, feats[it[0] % 4]))
    res['dynconv_' + name] = dict(us=us, algo_mb=byts / 1e6, gbs=byts / us / 1e3, frac_hbm=byts / us / 1e3 / PEAK,
                                 tflops=flops / us / 1e6, cublas_fp32_einsum_us=ref_us)

from boxinstseg_b200.models import build_head
head = build_head(dict(type='CondInstMaskHead', in_channels=16, in_stride=8, out_stride=4, topk_per_img=64, max_proposals=-1,
                       boxinst_enabled=True)).to(dev)
feat = torch.randn(2, 16, 100, 128, device=dev, generator=g).requires_grad_(True)
params = (torch.randn(128, 233, device=dev, generator=g) * 0.3).requires_grad_(True)
coors = torch.rand(128, 2, device=dev, generator=g) * torch.tensor([1024.0, 800.0], device=dev)
levels = torch.randint(0, 5, (128,), device=dev, generator=g)
img_inds = torch.arange(128, device=dev) // 64
res['condinst_head_fwd_us'] = timeit(lambda: head(feat, params, coors, levels, img_inds))
out = head(feat, params, coors, levels, img_inds)
gout = torch.randn_like(out)
res['condinst_head_bwd_us'] = timeit(lambda: torch.autograd.grad(out, [feat, params], gout, retain_graph=True))

from boxinstseg_b200.ops.tree_filter import MinimumSpanningTree, TreeFilter2D, bfs
mst, tf = MinimumSpanningTree(TreeFilter2D.norm2_distance), TreeFilter2D()
for name, (n, h, w) in dict(levelset_D_200x256=(16, 200, 256), box2mask_E_96x96=(8, 96, 96)).items():
    guide = F.interpolate(torch.randn(n, 3, h // 8, w // 8, device=dev, generator=g), size=(h, w), mode='bilinear') + \
        0.05 * torch.randn(n, 3, h, w, device=dev, generator=g)
    feat1 = torch.rand(n, 1, h, w, device=dev, generator=g).requires_grad_(True)
    res[f'mst_{name}_us'] = timeit(lambda: mst(guide), reps=5, warm=1)
    tree = mst(guide)
    res[f'bfs_{name}_us'] = timeit(lambda: bfs(tree, 4), reps=5, warm=1)
    res[f'treefilter_fwd_{name}_us'] = timeit(lambda: tf(feat1, guide, tree), reps=5, warm=1)
    o = tf(feat1, guide, tree, low_tree=False)
    res[f'treefilter_bwd_{name}_us'] = timeit(lambda: torch.autograd.grad(o.sum(), feat1, retain_graph=True), reps=5, warm=1)
    idx, par, chd = bfs(tree, 4)
    res[f'tree_levels_{name}'] = int(getattr(idx, '_bxs_levels')[1].max())

from boxinstseg_b200.models.losses import LCM, LevelsetLoss, projection_losses
s = torch.rand(16, 1, 200, 256, device=dev, generator=g).requires_grad_(True)
t = torch.randn(16, 3, 200, 256, device=dev, generator=g)
m = torch.zeros(16, 1, 200, 256, device=dev); m[:, :, 40:150, 60:200] = 1
pix = m.sum((1, 2, 3))
ls = LevelsetLoss()
res['levelset_fwd_bwd_us'] = timeit(lambda: torch.autograd.grad(ls(torch.cat([s, 1 - s], 1) * m, t * m, pix).sum(), s))
res['projection_fwd_bwd_us'] = timeit(lambda: torch.autograd.grad(projection_losses(s, m).sum(), s))
phi = torch.rand(8, 1, 96, 96, device=dev, generator=g).requires_grad_(True)
img96 = torch.rand(8, 3, 96, 96, device=dev, generator=g)
box96 = torch.zeros(8, 1, 96, 96, device=dev); box96[:, :, 20:70, 10:80] = 1
res['lcm_fwd_bwd_us'] = timeit(lambda: torch.autograd.grad(LCM(img96, phi, box96), phi))
from boxinstseg_b200.models.dense_heads import MeanField
cf = torch.randn(1, 3, 200, 256, device=dev, generator=g)
res['meanfield_kernel_us'] = timeit(lambda: MeanField(cf, alpha0=2, theta0=0.5, theta1=30, iter=10, base=0.1))
mf = MeanField(cf, alpha0=2, theta0=0.5, theta1=30, iter=10, base=0.1)
x = torch.rand(16, 1, 200, 256, device=dev, generator=g)
tg = torch.zeros(16, 1, 200, 256, device=dev); tg[:, :, 30:160, 50:220] = 1
res['meanfield_10it_16obj_us'] = timeit(lambda: mf(x, tg))
from bench import synthetic_case
from boxinstseg_b200.ops.boxinst import boxinst_targets
case = synthetic_case(1234)
img = case['img'].to(dev); boxes = [b.to(dev) for b in case['gt_bboxes']]
res['boxinst_targets_us'] = timeit(lambda: boxinst_targets(img, case['metas'], boxes))
print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}, indent=1))

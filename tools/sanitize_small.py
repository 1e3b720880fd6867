"""Small shapes of the kernels with hand-rolled synchronisation, for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from boxinstseg_b200.ops.tree_filter import MinimumSpanningTree, TreeFilter2D, bfs
from boxinstseg_b200.ops.tree_filter.functions.refine import refine
from boxinstseg_b200.ops.dynconv import dynconv1x1
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
n, h, w = 2, 40, 48
guide = F.interpolate(torch.randn(n, 3, h // 8, w // 8, device=dev, generator=g), size=(h, w), mode='bilinear') + 0.05 * torch.randn(n, 3, h, w, device=dev, generator=g)
tree = MinimumSpanningTree(TreeFilter2D.norm2_distance)(guide)
idx, par, chd = bfs(tree, 4)
tf = TreeFilter2D()
feat = torch.rand(n, 2, h * w, device=dev, generator=g).requires_grad_(True)
emb = guide.clone().requires_grad_(True)
ew = tf.build_edge_weight(emb, idx, par, False, chd)
out = refine(feat, ew, idx, par, chd, False)
gf, ge = torch.autograd.grad(out.sum(), [feat, emb])
f = torch.randn(1, 64, 16, 24, device=dev, generator=g).requires_grad_(True)
k = (torch.randn(1, 40, 64, device=dev, generator=g) * 0.1).requires_grad_(True)
o = dynconv1x1(f, k)
torch.autograd.grad(o.sum(), [f, k])
torch.cuda.synchronize()
print('ok', float(out.sum()), float(gf.abs().sum()), float(ge.abs().sum()), float(o.abs().sum()))

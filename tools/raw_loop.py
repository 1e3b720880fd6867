"""Raw C-ABI loop of the fused BoxInst loss (config A) for ncu / timing (diagnostic, GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_case, N_INST, H, W
from boxinstseg_b200 import _lib as L
from boxinstseg_b200.ops.boxinst import boxinst_targets
dev = torch.device('cuda:0')
case = synthetic_case(1234)
t = boxinst_targets(case['img'].to(dev), case['metas'], [b.to(dev) for b in case['gt_bboxes']])
it = torch.tensor([10000.0], device=dev)
lib = L.lib()
xs = [torch.randn(N_INST, 1, H, W, device=dev) * 2 for _ in range(8)]
gls = [torch.empty_like(xs[0]) for _ in range(8)]
inst_gt = case['gt_inds'].to(dev).to(torch.int32)
ws = torch.empty(lib.bxs_boxinst_loss_workspace_bytes(N_INST, H, W), dtype=torch.uint8, device=dev)
out = torch.empty(4, device=dev); g = torch.ones(2, device=dev)
st = L.stream()
def raw(i):
    x, gl = xs[i % 8], gls[i % 8]
    lib.bxs_boxinst_loss_forward(L.ptr(x), L.ptr(t.edge_bits), L.ptr(t.rects), L.ptr(inst_gt), L.ptr(t.gt_img), L.ptr(it), 10000.0, L.ptr(ws), L.ptr(out), N_INST, H, W, 2, st)
    lib.bxs_boxinst_loss_backward(L.ptr(x), L.ptr(t.edge_bits), L.ptr(t.rects), L.ptr(inst_gt), L.ptr(t.gt_img), L.ptr(ws), L.ptr(g), L.ptr(gl), N_INST, H, W, 2, st)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for i in range(8): raw(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n): raw(i)
e1.record(); torch.cuda.synchronize()
print(f'raw C ABI rotating 8 sets: {e0.elapsed_time(e1)/n*1e3:.1f} us/step', out.tolist())

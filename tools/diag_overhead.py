"""Where does a fused-loss step spend host time?  (diagnostic, run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_case, N_INST, H, W, ROTATE, ClockSampler
from boxinstseg_b200 import _lib as L
from boxinstseg_b200.ops.boxinst import boxinst_mask_loss, boxinst_targets

dev = torch.device('cuda:0')
case = synthetic_case(1234)
t = boxinst_targets(case['img'].to(dev), case['metas'], [b.to(dev) for b in case['gt_bboxes']])
gt_inds = case['gt_inds'].to(dev)
it = torch.tensor([10000.0], device=dev)
xs = [(torch.randn(N_INST, 1, H, W, device=dev) * 2).requires_grad_(True) for _ in range(ROTATE)]
ones = torch.ones((), device=dev)
ring = [None] * ROTATE

def step(i):
    x = xs[i % ROTATE]
    prj, pair = boxinst_mask_loss(x, t, gt_inds, it)
    torch.autograd.backward([prj, pair], [ones, ones])
    ring[i % ROTATE], x.grad = x.grad, None

def timed(label, n=50, sampler=False):
    for i in range(5): step(i)
    torch.cuda.synchronize()
    s = ClockSampler(0) if sampler else None
    if s: s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): step(i)
    e1.record(); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    if s: print(s.stop())
    print(f'{label}: gpu {e0.elapsed_time(e1)/n*1e3:.1f} us/step, host issue {t_issue/n*1e6:.1f} us/step, wall {t_all/n*1e6:.1f} us/step')

timed('no sampler')
timed('with sampler', sampler=True)
timed('no sampler again')

# raw C calls, no autograd
lib = L.lib()
x = xs[0].detach()
inst_gt = gt_inds.to(torch.int32)
ws = torch.empty(lib.bxs_boxinst_loss_workspace_bytes(N_INST, H, W), dtype=torch.uint8, device=dev)
out = torch.empty(4, device=dev); g = torch.ones(2, device=dev); gl = torch.empty_like(x)
st = L.stream()
def raw():
    lib.bxs_boxinst_loss_forward(L.ptr(x), L.ptr(t.edge_bits), L.ptr(t.rects), L.ptr(inst_gt), L.ptr(t.gt_img), L.ptr(it), 10000.0, L.ptr(ws), L.ptr(out), N_INST, H, W, 2, st)
    lib.bxs_boxinst_loss_backward(L.ptr(x), L.ptr(t.edge_bits), L.ptr(t.rects), L.ptr(inst_gt), L.ptr(t.gt_img), L.ptr(ws), L.ptr(g), L.ptr(gl), N_INST, H, W, 2, st)
for _ in range(5): raw()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(100): raw()
e1.record(); ti = time.perf_counter() - t0; torch.cuda.synchronize()
print(f'raw C ABI (L2-warm): gpu {e0.elapsed_time(e1)/100*1e3:.1f} us/step, host issue {ti/100*1e6:.1f} us/step')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(30): step(i)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)

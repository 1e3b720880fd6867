"""Dump the CUDA graph of two public-op steps (forward + autograd backward) as DOT and summarise nodes/edges."""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_case, N_INST, H, W
from boxinstseg_b200.ops.boxinst import boxinst_mask_loss, boxinst_targets
dev = torch.device('cuda:0')
case = synthetic_case(1234)
t = boxinst_targets(case['img'].to(dev), case['metas'], [b.to(dev) for b in case['gt_bboxes']])
it = torch.tensor([10000.0], device=dev)
gi = case['gt_inds'].to(dev).to(torch.int32)
ones = torch.ones((), device=dev)
xs = [(torch.randn(N_INST, 1, H, W, device=dev) * 2).requires_grad_(True) for _ in range(2)]
def step(x):
    prj, pair = boxinst_mask_loss(x, t, gi, it)
    torch.autograd.backward([prj, pair], [ones, ones])
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for x in xs: step(x); x.grad = None
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
g.enable_debug_mode()
with torch.cuda.graph(g):
    for x in xs: step(x)
os.makedirs('gpurun_out', exist_ok=True)
g.debug_dump('gpurun_out/op_graph.dot')
txt = open('gpurun_out/op_graph.dot').read()
nodes = re.findall(r'label="\{?\s*([^|\n"]+)', txt)
print('nodes:', len(nodes))
for n in nodes: print('  ', n.strip()[:100])
print('edges:', len(re.findall(r'->', txt)))

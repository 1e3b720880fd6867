"""Config A of BASELINE.json at FULL size (B=2 images of 3x800x1024, 8 GT each, N=128 instances, loss grid 200x256,
seed 1234 = the bench inputs): the fused BoxInst mask loss and its gradient against

  * the CPU oracle (oracle/boxinst.py: restated CondInstMaskHead.loss, float64), and
  * the REFERENCE'S OWN compiled CUDA pairwise op (oracle/_ref/pairwise_ext_ref, built by oracle/Makefile from
    mmdet/ops/pairwise/csrc unmodified) inside a restatement of condinst_head.py:1288-1343 that materialises what the
    reference materialises.

Tolerance: 1e-3 relative on loss and gradient (BASELINE.json north_star); observed ~1e-6."""
import glob
import importlib.util
import os

import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = 'cuda:0'
TOL = 1e-3


@pytest.fixture(scope='module')
def case():
    import bench
    return bench.synthetic_case(1234)


@pytest.fixture(scope='module')
def ours(case):
    """(loss_prj, loss_pairwise, gradient) of the CUDA path for upstream gradients (1, 1), with and without a plan."""
    from boxinstseg_b200.ops.boxinst import boxinst_loss_plan, boxinst_mask_loss, boxinst_targets
    dev = torch.device(DEV)
    t = boxinst_targets(case['img'].to(dev), case['metas'], [b.to(dev) for b in case['gt_bboxes']], want_similarity=True)
    it = torch.tensor([10000.0], device=dev)
    gi = case['gt_inds'].to(dev)
    res = []
    for planned in (True, False):
        x = case['logits'].to(dev).requires_grad_(True)
        plan = boxinst_loss_plan(t, gi, x.shape[2], x.shape[3], 2) if planned else None
        prj, pair = boxinst_mask_loss(x, t, gi, it, plan=plan)
        (gx,) = torch.autograd.grad(prj + pair, x)
        res.append((prj.detach(), pair.detach(), gx))
    # the plan is only a precomputed work list: same bits with and without it
    assert torch.equal(res[0][2], res[1][2]) and res[0][0].item() == res[1][0].item() and res[0][1].item() == res[1][1].item()
    return t, res[0]


def test_full_size_against_cpu_oracle(case, ours):
    from oracle import boxinst as ob
    _, (prj, pair, gx) = ours
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    sim, bms = ob.boxinst_targets(case['img'], case['metas'], case['gt_bboxes'])
    x = case['logits'].double().requires_grad_(True)
    bm = torch.cat(bms)[case['gt_inds']][:, None].double()
    p64, q64 = ob.boxinst_mask_loss(x, sim[case['img_inds']].double(), bm, warmup_factor=1.0)
    (g64,) = torch.autograd.grad(p64 + q64, x)
    assert abs(prj.item() - p64.item()) <= TOL * abs(p64.item()), (prj.item(), p64.item())
    assert abs(pair.item() - q64.item()) <= TOL * abs(q64.item()), (pair.item(), q64.item())
    assert rel_err(gx.cpu(), g64) <= TOL
    # element-wise as well: arg-max ties do not occur on continuous random logits
    assert torch.allclose(gx.cpu().double(), g64, rtol=TOL, atol=1e-9)


def test_full_size_against_reference_cuda_extension(case, ours):
    hits = glob.glob(os.path.join(ROOT, 'oracle', '_ref', 'pairwise_ext_ref*.so'))
    assert hits, 'oracle/_ref/pairwise_ext_ref*.so is missing: run `make -C oracle` where /root/reference exists'
    spec = importlib.util.spec_from_file_location('pairwise_ext_ref', hits[0])
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    t, (prj, pair, gx) = ours
    dev = torch.device(DEV)

    class RefPairwise(torch.autograd.Function):            # mmdet/ops/pairwise/pairwise.py:6-26
        @staticmethod
        def forward(ctx, logits, size, dilation):
            pw = ext.pairwise_nlog_forward(size, dilation, logits)
            ctx.save_for_backward(logits, pw)
            ctx.cfg = (size, dilation)
            return pw

        @staticmethod
        def backward(ctx, g):
            logits, pw = ctx.saved_tensors
            return ext.pairwise_nlog_backward(ctx.cfg[0], ctx.cfg[1], logits, pw, g.contiguous()), None, None

    gi = case['gt_inds'].to(dev)
    sim = t.similarity[case['img_inds'].to(dev)]                       # the G-fold gather of condinst_head.py:1316
    bm = torch.cat(t.bitmasks())[gi][:, None]
    x = case['logits'].to(dev).requires_grad_(True)

    def dice(a, b):                                                    # condinst_head.py:117-131
        a, b = a.flatten(1), b.flatten(1)
        return 1.0 - 2.0 * (a * b).sum(1) / ((a * a).sum(1) + (b * b).sum(1) + 1e-5)

    scores = x.sigmoid()
    prj_ref = (dice(scores.max(dim=2, keepdim=True)[0], bm.max(dim=2, keepdim=True)[0]) +
               dice(scores.max(dim=3, keepdim=True)[0], bm.max(dim=3, keepdim=True)[0])).mean()      # :134-143
    w = (sim >= 0.3).float() * bm
    pair_ref = (RefPairwise.apply(x, 3, 2) * w).sum() / w.sum().clamp(min=1.0)                        # :1318-1332
    (g_ref,) = torch.autograd.grad(prj_ref + pair_ref, x)
    assert abs(prj.item() - prj_ref.item()) <= TOL * abs(prj_ref.item())
    assert abs(pair.item() - pair_ref.item()) <= TOL * abs(pair_ref.item())
    assert rel_err(gx, g_ref) <= TOL

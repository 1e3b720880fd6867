"""GPU: BoxMatchingCost against the golden vector minted from the reference class and against
the reference recipe upsample -> sigmoid -> max -> pairwise dice (match_cost.py:386-425, box2mask_head.py:157-161)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = 'cuda:0'


def test_match_cost_golden(golden):
    from boxinstseg_b200.core import BoxMatchingCost
    g = golden('projection')
    cost = BoxMatchingCost(weight=2.0, pred_act=False, eps=1.0)(T(g['scores']).to(DEV), T(g['targets']).to(DEV))
    assert torch.allclose(cost.cpu(), T(g['match_cost_w2_eps1']), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('Q,G,h,w,H,W', [(20, 5, 32, 40, 128, 160), (7, 3, 25, 31, 100, 90), (100, 8, 64, 64, 256, 256),
                                         (3, 2, 16, 300, 64, 1300), (4, 2, 20, 24, 20, 24)])
def test_match_cost_fused_upsample(Q, G, h, w, H, W):
    from boxinstseg_b200.core import BoxMatchingCost, projection_profiles
    gen = torch.Generator().manual_seed(Q)
    pred = torch.randn(Q, h, w, generator=gen) * 3
    gt = torch.zeros(G, H, W)
    for i in range(G):
        gt[i, H // 8 + i: H // 2 + 3 * i, W // 6: W // 6 + W // 3 + i] = 1
    # the reference resizes in float32 (source coordinates rounded to float32): compare with exactly that
    up = F.interpolate(pred.to(DEV).unsqueeze(1), (H, W), mode='bilinear', align_corners=False).cpu().double()
    row, col = projection_profiles(pred.to(DEV), (H, W), sigmoid=True)
    assert torch.allclose(row.cpu().double(), up.sigmoid().amax(3)[:, 0], rtol=1e-5, atol=2e-6)
    assert torch.allclose(col.cpu().double(), up.sigmoid().amax(2)[:, 0], rtol=1e-5, atol=2e-6)
    mc = BoxMatchingCost(weight=5.0, pred_act=True, eps=1.0)
    fused = mc.cost_from_lowres(pred.to(DEV), gt.to(DEV))
    two_step = mc(F.interpolate(pred.to(DEV).unsqueeze(1), (H, W), mode='bilinear', align_corners=False), gt.to(DEV).unsqueeze(1))
    s = up.sigmoid()
    def dice(p, g_):
        return 1 - (2 * p @ g_.t() + 1.0) / (p.pow(2).sum(1)[:, None] + g_.pow(2).sum(1)[None] + 1.0)
    gd = gt.double()
    ref = 5.0 * (dice(s.amax(3)[:, 0], gd.amax(2)) + dice(s.amax(2)[:, 0], gd.amax(1)))
    assert torch.allclose(fused.cpu().double(), ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(two_step, fused, rtol=1e-5, atol=1e-6)


def test_mask_hungarian_assigner_front_half():
    """MaskHungarianAssigner.assign (mask_hungarian_assigner.py:60-132) with the Box2Mask costs: the GPU cost matrix equals the
    reference recipe (softmax classification cost + BoxMatchingCost on the UPSAMPLED prediction, box2mask_head.py:157-169) and
    the matching equals scipy's on that matrix; the low-resolution shortcut gives the same assignment."""
    from scipy.optimize import linear_sum_assignment
    from boxinstseg_b200.core import MaskHungarianAssigner
    gen = torch.Generator().manual_seed(3)
    Q, G, h, w, H, W, ncls = 100, 6, 64, 64, 256, 256, 81
    cls_pred = torch.randn(Q, ncls, generator=gen)
    mask_pred = torch.randn(Q, h, w, generator=gen) * 3
    gt_labels = torch.randint(0, 80, (G,), generator=gen)
    gt = torch.zeros(G, H, W)
    for i in range(G):
        gt[i, 20 + 30 * i: 60 + 30 * i, 10 + 25 * i: 90 + 25 * i] = 1
    asg = MaskHungarianAssigner(cls_cost=dict(type='ClassificationCost', weight=2.0),
                                dice_cost=dict(type='BoxMatchingCost', weight=5.0, pred_act=True, eps=1.0))
    up = F.interpolate(mask_pred.to(DEV).unsqueeze(1), (H, W), mode='bilinear', align_corners=False)
    res = asg.assign(cls_pred.to(DEV), up, gt_labels.to(DEV), gt.to(DEV).unsqueeze(1), None)
    res_low = asg.assign(cls_pred.to(DEV), mask_pred.to(DEV), gt_labels.to(DEV), gt.to(DEV), None, lowres=True)
    # reference recipe in float64 on the host
    s = up.cpu().double().sigmoid()[:, 0]
    gd = gt.double()
    def dice(p, g_):
        return 1 - (2 * p @ g_.t() + 1.0) / (p.pow(2).sum(1)[:, None] + g_.pow(2).sum(1)[None] + 1.0)
    ref = -2.0 * cls_pred.double().softmax(-1)[:, gt_labels] + 5.0 * (dice(s.amax(2), gd.amax(2)) + dice(s.amax(1), gd.amax(1)))
    cm = asg.cost_matrix(cls_pred.to(DEV), up, gt_labels.to(DEV), gt.to(DEV).unsqueeze(1))
    assert torch.allclose(cm.cpu().double(), ref, rtol=1e-5, atol=1e-6)
    rows, cols = linear_sum_assignment(ref.numpy())
    want = torch.zeros(Q, dtype=torch.long)
    want[torch.from_numpy(rows)] = torch.from_numpy(cols) + 1
    assert torch.equal(res.gt_inds.cpu(), want) and torch.equal(res_low.gt_inds.cpu(), want)          # index work: exact
    assert torch.equal(res.labels.cpu()[torch.from_numpy(rows)], gt_labels[torch.from_numpy(cols)])
    assert res.num_gts == G and (res.gt_inds > 0).sum().item() == G

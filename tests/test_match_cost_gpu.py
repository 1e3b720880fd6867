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

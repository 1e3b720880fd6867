"""GPU parity of the CondInst dynamic mask head (a1) against the golden vectors minted from the
reference's own CondInstMaskHead.forward and against the float64 oracle."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def _run(dev, feat, params, coors, levels, img_inds, g_out, in_stride=8, out_stride=4, rel=True):
    from boxinstseg_b200.ops.condinst import dynamic_mask_head
    f = feat.to(dev).requires_grad_(True)
    p = params.to(dev).requires_grad_(True)
    soi = torch.tensor([64, 128, 256, 512, 1024], device=dev)
    out = dynamic_mask_head(f, p, coors.to(dev), levels.to(dev), img_inds.to(dev), soi, in_stride, out_stride,
                            rel_coors=rel)
    gf, gp = torch.autograd.grad((out * g_out.to(dev)).sum(), [f, p])
    return out.cpu(), gf.cpu(), gp.cpu()


def test_head_golden(golden):
    dev = torch.device('cuda:0')
    g = golden('condinst_head')
    out, gf, gp = _run(dev, T(g['feat']), T(g['params']), T(g['coors']), T(g['level_inds']), T(g['img_inds']),
                       T(g['g_out']))
    assert torch.allclose(out, T(g['out']), rtol=1e-3, atol=1e-5)
    assert rel_err(gf, T(g['g_feat'])) < 1e-4 and rel_err(gp, T(g['g_params'])) < 1e-4
    assert torch.allclose(gf, T(g['g_feat']), rtol=1e-3, atol=1e-4)
    assert torch.allclose(gp, T(g['g_params']), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize('cfg', [dict(B=2, C=16, h=25, w=32, N=9, in_stride=8, out_stride=4, rel=True),
                                 dict(B=3, C=8, h=13, w=50, N=7, in_stride=8, out_stride=8, rel=True),    # factor 1
                                 dict(B=2, C=16, h=12, w=20, N=5, in_stride=8, out_stride=2, rel=True),   # factor 4
                                 dict(B=2, C=4, h=40, w=33, N=6, in_stride=8, out_stride=4, rel=False),
                                 dict(B=4, C=16, h=9, w=9, N=3, in_stride=8, out_stride=4, rel=True)])    # empty images
def test_head_vs_oracle(cfg):
    from oracle.boxinst import condinst_mask_head
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(17)
    B, C, h, w, N = cfg['B'], cfg['C'], cfg['h'], cfg['w'], cfg['N']
    cin = C + (2 if cfg['rel'] else 0)
    P = cin * 8 + 64 + 8 + 8 + 8 + 1
    feat = torch.randn(B, C, h, w, generator=gen)
    params = torch.randn(N, P, generator=gen) * 0.3
    coors = torch.rand(N, 2, generator=gen) * torch.tensor([w * 8.0, h * 8.0])
    levels = torch.randint(0, 5, (N,), generator=gen)
    img_inds = torch.sort(torch.randint(0, min(B, 2), (N,), generator=gen))[0]
    if N > 2:
        img_inds[-1] = B - 1                      # unsorted-by-construction tail is still valid input
    f = cfg['in_stride'] // cfg['out_stride']
    g_out = torch.randn(N, 1, f * h, f * w, generator=gen)
    f64 = feat.double().requires_grad_(True)
    p64 = params.double().requires_grad_(True)
    ref = condinst_mask_head(f64, p64, coors.double(), levels, img_inds, cfg['in_stride'], cfg['out_stride'],
                             rel_coors=cfg['rel'])
    rf, rp = torch.autograd.grad((ref * g_out.double()).sum(), [f64, p64])
    out, gf, gp = _run(dev, feat, params, coors, levels, img_inds, g_out, cfg['in_stride'], cfg['out_stride'],
                       cfg['rel'])
    assert rel_err(out, ref.detach()) < 1e-5
    assert rel_err(gf, rf) < 1e-4 and rel_err(gp, rp) < 1e-4
    out2, gf2, gp2 = _run(dev, feat, params, coors, levels, img_inds, g_out, cfg['in_stride'], cfg['out_stride'],
                          cfg['rel'])
    assert torch.equal(gf, gf2) and torch.equal(gp, gp2)          # deterministic


def test_head_config_a_shapes():
    """config A: feat [2,16,100,128], N=128 -> [128,1,200,256]; cross-check against a torch composition."""
    dev = torch.device('cuda:0')
    from boxinstseg_b200.models import build_head
    gen = torch.Generator().manual_seed(5)
    head = build_head(dict(type='CondInstMaskHead', in_channels=16, in_stride=8, out_stride=4, topk_per_img=64,
                           max_proposals=-1, boxinst_enabled=True)).to(dev)
    feat = torch.randn(2, 16, 100, 128, generator=gen).to(dev)
    params = (torch.randn(128, 233, generator=gen) * 0.3).to(dev)
    coors = (torch.rand(128, 2, generator=gen) * torch.tensor([1024.0, 800.0])).to(dev)
    levels = torch.randint(0, 5, (128,), generator=gen).to(dev)
    img_inds = torch.arange(128, device=dev) // 64
    out = head(feat, params, coors, levels, img_inds)
    assert out.shape == (128, 1, 200, 256)
    from oracle.boxinst import condinst_mask_head
    ref = condinst_mask_head(feat[:, :, :20].cpu().double(), params[:4].cpu().double(), coors[:4].cpu().double(),
                             levels[:4].cpu(), img_inds[:4].cpu())
    part = head(feat[:, :, :20].contiguous(), params[:4], coors[:4], levels[:4], img_inds[:4])
    assert rel_err(part.cpu(), ref) < 1e-5

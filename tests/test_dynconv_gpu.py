"""GPU parity of the tcgen05 dynamic 1x1 convolution (a2/a3/a4) against a float64 reference.
Tolerance: TF32 operands (10-bit mantissa) with K=256 -> ~5e-4 relative; the reference's own
cuDNN conv2d runs in TF32 by default on Ampere and later GPUs."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _ref(feat, kern):
    return torch.einsum('bic,bchw->bihw', kern.double(), feat.double())


@pytest.mark.parametrize('B,C,h,w,I', [(1, 256, 16, 32, 16), (2, 256, 50, 64, 100), (1, 256, 200, 256, 37),
                                       (2, 128, 25, 36, 130), (1, 256, 96, 96, 300), (1, 32, 8, 4, 5)])
def test_dynconv_forward(B, C, h, w, I):
    from boxinstseg_b200.ops.dynconv import dynconv1x1
    gen = torch.Generator().manual_seed(I)
    feat = torch.randn(B, C, h, w, generator=gen)
    kern = torch.randn(B, I, C, generator=gen) * 0.05
    out = dynconv1x1(feat.to(DEV), kern.to(DEV))
    torch.cuda.synchronize()
    ref = _ref(feat, kern)
    assert out.shape == ref.shape
    assert rel_err(out.cpu(), ref) < 2e-3
    # exactness probe: inputs representable in TF32 -> only fp32 accumulation error remains
    f2 = (torch.randint(-8, 9, (B, C, h, w), generator=gen).float() / 8)
    k2 = (torch.randint(-8, 9, (B, I, C), generator=gen).float() / 16)
    out2 = dynconv1x1(f2.to(DEV), k2.to(DEV))
    assert torch.allclose(out2.cpu().double(), _ref(f2, k2), rtol=1e-5, atol=1e-4)


def test_dynconv_backward_and_call_sites():
    from boxinstseg_b200.ops.dynconv import box2mask_mask_pred, dynconv1x1, solo_dynamic_conv
    gen = torch.Generator().manual_seed(0)
    feat = torch.randn(2, 256, 20, 24, generator=gen).to(DEV).requires_grad_(True)
    kern = (torch.randn(2, 48, 256, generator=gen) * 0.05).to(DEV).requires_grad_(True)
    gout = torch.randn(2, 48, 20, 24, generator=gen).to(DEV)
    out = dynconv1x1(feat, kern)
    gf, gk = torch.autograd.grad((out * gout).sum(), [feat, kern])
    f64, k64 = feat.detach().double().requires_grad_(True), kern.detach().double().requires_grad_(True)
    rf, rk = torch.autograd.grad((torch.einsum('bic,bchw->bihw', k64, f64) * gout.double()).sum(), [f64, k64])
    assert rel_err(gf, rf) < 2e-3 and rel_err(gk, rk) < 2e-3
    # SOLO order: cell s = gy*S + gx
    kp = torch.randn(2, 256, 4, 4, generator=gen).to(DEV) * 0.05
    solo = solo_dynamic_conv(feat.detach(), kp)
    ref = torch.nn.functional.conv2d(feat.detach().reshape(1, 512, 20, 24).double(),
                                     kp.permute(0, 2, 3, 1).reshape(-1, 256, 1, 1).double(), groups=2).view(2, 16, 20, 24)
    assert rel_err(solo, ref) < 2e-3
    assert rel_err(box2mask_mask_pred(kern.detach(), feat.detach()), out.detach().double()) < 1e-6

"""f2 on the GPU: the SOLO grid targets computed on CUDA tensors equal the CPU evaluation of the same code (which the CPU tests
pin on the oracle and on the reference's golden vectors), bit for bit, and ``solo_target_single`` adds the a18 resizes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
CFG = dict(scale_ranges=((1, 24), (12, 48), (24, 96), (48, 192), (96, 512)), strides=(8, 8, 16, 32, 32),
           seg_num_grids=[40, 36, 24, 16, 12], sigma=0.2, num_classes=80)


@pytest.mark.parametrize('seed', [11, 12])
def test_solo_targets_cuda_equals_cpu(seed):
    from boxinstseg_b200.models.dense_heads.solo_targets import solo_grid_targets, solo_target_single
    from oracle.make_golden_solo import case
    boxes, labels, masks, img, lst, fs = case(seed)
    masks = torch.from_numpy(masks)
    want = solo_grid_targets(boxes, labels, masks, fs, **CFG)
    got = solo_target_single(boxes.to(DEV), labels.to(DEV), masks.to(DEV), img.to(DEV), lst.to(DEV), fs, **CFG)
    for lvl in range(5):
        for k in range(3):
            assert torch.equal(got[k][lvl].cpu(), want[k][lvl])
        assert torch.allclose(got[3][lvl].cpu(), F.interpolate(img[None], size=fs[lvl], mode='bilinear'), atol=1e-5)
        assert torch.allclose(got[4][lvl].cpu(), F.interpolate(lst[None], size=fs[lvl], mode='bilinear'), atol=1e-5)

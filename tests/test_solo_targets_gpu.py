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


@pytest.mark.parametrize('seed', [0, 1, 2, 6])
@pytest.mark.parametrize('best', [False, True])
def test_disco_targets_cuda_equal_oracle_and_golden_cases(seed, best):
    """DiscoBox's two builders (discobox_head.py:1362-1529) on CUDA tensors against the oracle's loop form (whose golden
    vectors come from the reference's own methods; seeds 0-2 are the golden cases)."""
    import numpy as np
    from boxinstseg_b200.models.dense_heads.disco_targets import disco_target_single
    from oracle import solo_targets as ost
    from oracle.make_golden_disco import CFG as DCFG, case
    boxes, labels, masks, fsize = case(seed) if seed < 5 else case(seed, H=96, W=256, G=14)
    if seed == 6:
        gen = torch.Generator().manual_seed(9)
        masks = masks * (torch.rand(masks.shape, generator=gen) < 0.6).numpy().astype(np.uint8)
    want = ost.disco_target_single(boxes, labels, masks, fsize, best=best, **DCFG)
    got = disco_target_single(boxes.to(DEV), labels.to(DEV), torch.from_numpy(masks).to(DEV), fsize, best=best, **DCFG)
    for k in range(3):
        for w, g in zip(want[k], got[k]):
            assert g.is_cuda and w.shape == g.shape and torch.equal(w, g.cpu())
    assert [list(x) for x in want[3]] == [x.cpu().tolist() for x in got[3]]

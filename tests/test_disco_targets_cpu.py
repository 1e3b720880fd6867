"""f2: DiscoBox's SOLO target builders (discobox_head.py:1362-1529).  (1) the oracle restatement reproduces the golden vectors
minted from the reference's own two methods (oracle/make_golden_disco.py); (2) the device-tensor builder
(models/dense_heads/disco_targets.py, plain torch: runs on any device) equals the oracle bit for bit -- instance labels, category
grid, positive flags and the covered-cell order -- on the golden cases, on random masks and on the empty cases.  The GPU twin
runs the same comparison with CUDA tensors (tests/test_solo_targets_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle import solo_targets as ost
from oracle.make_golden_disco import CFG, case


def _same(want, got):
    for w_list, g_list in zip(want[:3], got[:3]):
        for w, g in zip(w_list, g_list):
            assert w.shape == g.shape and torch.equal(w, g.cpu())
    assert [list(x) for x in want[3]] == [x.cpu().tolist() for x in got[3]]


@pytest.mark.parametrize('seed', [0, 1, 2])
@pytest.mark.parametrize('best', [False, True])
def test_oracle_reproduces_reference_golden(golden, seed, best):
    g = golden('disco_targets')
    name = 'best' if best else 'gen'
    boxes, labels, masks, fsize = case(seed)
    ins, cate, ind, order = ost.disco_target_single(boxes, labels, masks, fsize, best=best, **CFG)
    for lvl in range(5):
        assert np.array_equal(cate[lvl].numpy(), g[f's{seed}_{name}_cate{lvl}'].astype(np.int64))
        assert np.array_equal(ind[lvl].numpy(), g[f's{seed}_{name}_ind{lvl}'])
        assert np.array_equal(np.asarray(order[lvl], dtype=np.int32), g[f's{seed}_{name}_order{lvl}'])
        shape = tuple(g[f's{seed}_{name}_insshape{lvl}'])
        assert tuple(ins[lvl].shape) == shape
        assert np.array_equal(np.unpackbits(g[f's{seed}_{name}_ins{lvl}'], axis=-1)[..., :shape[-1]], ins[lvl].numpy())


@pytest.mark.parametrize('seed', [0, 1, 2, 5, 6])
@pytest.mark.parametrize('best', [False, True])
def test_device_builder_equals_oracle(seed, best):
    from boxinstseg_b200.models.dense_heads.disco_targets import disco_target_single
    boxes, labels, masks, fsize = case(seed) if seed < 5 else case(seed, H=96, W=256, G=14)
    if seed == 6:                                                       # ragged random masks instead of rectangles
        gen = torch.Generator().manual_seed(9)
        masks = masks * (torch.rand(masks.shape, generator=gen) < 0.6).numpy().astype(np.uint8)
    want = ost.disco_target_single(boxes, labels, masks, fsize, best=best, **CFG)
    got = disco_target_single(boxes, labels, torch.from_numpy(masks), fsize, best=best, **CFG)
    _same(want, got)
    assert sum(len(x) for x in want[3]) > 0


@pytest.mark.parametrize('best', [False, True])
def test_no_ground_truth_and_all_masks_empty(best):
    from boxinstseg_b200.models.dense_heads.disco_targets import disco_target_single
    boxes, labels, masks, fsize = case(3)
    for b, l, m in ((boxes[:0], labels[:0], masks[:0]), (boxes, labels, masks * 0)):
        got = disco_target_single(b, l, torch.from_numpy(m), fsize, best=best, **CFG)
        if len(b):
            _same(ost.disco_target_single(b, l, m, fsize, best=best, **CFG), got)
        assert all(x.shape == (0, fsize[0], fsize[1]) for x in got[0]) and all(x.numel() == 0 for x in got[3])
        assert all((c == CFG['num_classes']).all() for c in got[1]) and not any(i.any() for i in got[2])

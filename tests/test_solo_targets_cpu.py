"""f2: the device-side SOLO grid targets (plain torch + integer resize; runs on CPU tensors here, on CUDA tensors in
tests/test_solo_targets_gpu.py) against the oracle's restatement of box_solov2_head.py:390-472, against the golden vectors
minted from the reference's own method (oracle/make_golden_solo.py), and the uint8 bilinear resize against OpenCV itself."""
import os

import numpy as np
import pytest
import torch

from boxinstseg_b200.models.dense_heads.solo_targets import cv2_resize_linear_u8, solo_grid_targets

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'solo_targets.npz')
CFG = dict(scale_ranges=((1, 24), (12, 48), (24, 96), (48, 192), (96, 512)), strides=(8, 8, 16, 32, 32),
           seg_num_grids=[40, 36, 24, 16, 12], sigma=0.2, num_classes=80)


def test_uint8_bilinear_resize_equals_opencv():
    cv2 = pytest.importorskip('cv2')
    gen = np.random.default_rng(0)
    for (h, w, nh, nw) in [(160, 192, 40, 48), (160, 192, 20, 24), (160, 192, 10, 12), (97, 131, 24, 33), (50, 70, 13, 18),
                           (64, 64, 64, 64), (33, 20, 8, 5), (200, 300, 50, 75), (101, 203, 25, 50)]:
        for kind in range(3):
            if kind == 0:                                    # rectangles (box supervision)
                m = np.zeros((h, w), np.uint8)
                y0, x0 = gen.integers(0, h // 2), gen.integers(0, w // 2)
                m[y0:y0 + gen.integers(1, h - y0), x0:x0 + gen.integers(1, w - x0)] = 1
            elif kind == 1:
                m = (gen.random((h, w)) < 0.5).astype(np.uint8)
            else:
                m = gen.integers(0, 256, (h, w)).astype(np.uint8)
            want = cv2.resize(m, (nw, nh), interpolation=cv2.INTER_LINEAR)
            got = cv2_resize_linear_u8(torch.from_numpy(m), nh, nw).numpy()
            assert np.array_equal(got, want), (h, w, nh, nw, kind)


def _case(seed, H=160, W=192, G=9):
    from oracle.make_golden_solo import case
    return case(seed, H, W, G)


@pytest.mark.parametrize('seed', [3, 4, 5, 6])
def test_solo_targets_equal_oracle(seed):
    from oracle.solo_targets import solo_target_single as oracle_targets
    boxes, labels, masks, img, lst, fs = _case(seed)
    want = oracle_targets(boxes, labels, masks, img, lst, fs, **CFG)
    got = solo_grid_targets(boxes, labels, torch.from_numpy(masks), fs, **CFG)
    for lvl in range(5):
        assert torch.equal(got[0][lvl], want[0][lvl]) and torch.equal(got[1][lvl], want[1][lvl])
        assert torch.equal(got[2][lvl], want[2][lvl])
    # the compact form describes the same canvas
    comp = solo_grid_targets(boxes, labels, torch.from_numpy(masks), fs, dense=False, **CFG)
    for lvl in range(5):
        winner, small = comp[0][lvl]
        pos = (winner >= 0).nonzero().flatten()
        assert torch.equal(small[winner[pos]], want[0][lvl][pos][:, :small.shape[1], :small.shape[2]])


def test_solo_targets_equal_reference_golden():
    g = np.load(GOLD)
    for seed in (0, 1, 2):
        boxes, labels = torch.from_numpy(g[f's{seed}_boxes']), torch.from_numpy(g[f's{seed}_labels'])
        shape = tuple(g[f's{seed}_shape'])
        masks = torch.from_numpy(np.unpackbits(g[f's{seed}_masks'], axis=-1)[..., :shape[-1]].reshape(shape))
        H, W = shape[-2:]
        fs = [(H // 4, W // 4), (H // 4, W // 4), (H // 8, W // 8), (H // 16, W // 16), (H // 16, W // 16)]
        got = solo_grid_targets(boxes, labels, masks, fs, **CFG)
        for lvl in range(5):
            assert torch.equal(got[1][lvl], torch.from_numpy(g[f's{seed}_cate{lvl}'].astype(np.int64)))
            ind = torch.from_numpy(g[f's{seed}_ind{lvl}'])
            assert torch.equal(got[2][lvl], ind)
            ishape = tuple(g[f's{seed}_insshape{lvl}'])
            want = torch.from_numpy(np.unpackbits(g[f's{seed}_ins{lvl}'], axis=-1)[..., :ishape[-1]].reshape(ishape))
            assert torch.equal(got[0][lvl][ind], want)
            assert int(got[0][lvl].sum()) == int(want.sum())                 # nothing outside the positive cells


def test_no_ground_truth_and_tiny_masks():
    fs = [(20, 24), (20, 24), (10, 12), (5, 6), (5, 6)]
    out = solo_grid_targets(torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64), torch.zeros(0, 80, 96, dtype=torch.uint8), fs,
                            **CFG)
    assert all(int(c.min()) == 80 for c in out[1]) and not any(bool(i.any()) for i in out[2])
    boxes = torch.tensor([[2.0, 2.0, 30.0, 30.0]])
    masks = torch.zeros(1, 80, 96, dtype=torch.uint8)
    masks[0, 2:5, 2:5] = 1                                                   # 9 pixels: skipped (:441-442)
    out = solo_grid_targets(boxes, torch.tensor([7]), masks, fs, **CFG)
    assert not any(bool(i.any()) for i in out[2])

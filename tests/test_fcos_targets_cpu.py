"""f2: CondInstBoxHead.get_targets (condinst_head.py:477-633), CPU side: (1) the oracle restatement reproduces the golden
vectors minted from the reference's own methods (oracle/make_golden_fcos.py); (2) the HOST build of the kernel's own
source (csrc/assign_core.cuh through tests/host_harness/assign_host.cpp) equals the oracle bit for bit: labels, indices and
fp32 regression targets, on the golden cases, on random boxes, on an image without ground truth and with more ground truths
than the kernel stages in shared memory.  The GPU twin of this file is tests/test_fcos_targets_gpu.py."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import fcos_targets as oft
from oracle.make_golden_fcos import CASES, CFG, case
from tests.helpers import host_twin


def _oracle(points, boxes, labels, flags):
    return oft.get_targets(points, boxes, labels, CFG['regress_ranges'], CFG['strides'], CFG['num_classes'],
                           flags['center_sampling'], 1.5, flags['norm_on_bbox'])


def _host(points, boxes, labels, flags):
    lib = host_twin('assign_host')
    B, L = len(boxes), len(points)
    pts = torch.cat(points).contiguous()
    counts = [p.shape[0] for p in points]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    P = int(off[-1])
    gb = torch.cat([b.reshape(-1, 4) for b in boxes]).contiguous().float()
    gl = torch.cat(labels).contiguous().long()
    goff = torch.tensor(np.concatenate([[0], np.cumsum([b.shape[0] for b in boxes])]), dtype=torch.int64)
    f32 = lambda v: torch.tensor(v, dtype=torch.float64).float()
    lo, hi = f32([float(r[0]) for r in CFG['regress_ranges']]), f32([float(r[1]) for r in CFG['regress_ranges']])
    sr, st = f32([s * 1.5 for s in CFG['strides']]), f32([float(s) for s in CFG['strides']])
    lab = torch.empty(B * P, dtype=torch.int64)
    tgt = torch.empty(B * P, 4)
    ind = torch.empty(B * P, dtype=torch.int64)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.host_fcos_targets(p(pts), p(gb), p(gl), p(goff), p(lab), p(tgt), p(ind), ctypes.c_int64(B), ctypes.c_int64(L),
                               off.ctypes.data_as(ctypes.c_void_p), p(lo), p(hi), p(sr), p(st),
                               int(flags['center_sampling']), int(flags['norm_on_bbox']), ctypes.c_int64(CFG['num_classes']))
    assert rc == 0
    sizes = [B * c for c in counts]
    return lab.split(sizes), tgt.split(sizes), ind.split(sizes)


@pytest.mark.parametrize('seed,kw', CASES)
def test_oracle_reproduces_reference_golden(golden, seed, kw):
    g = golden('fcos_targets')
    points, boxes, labels, flags = case(seed, **kw)
    lab, tgt, ind = _oracle(points, boxes, labels, flags)
    assert np.array_equal(torch.cat(lab).numpy(), g[f's{seed}_labels'].astype(np.int64))
    assert np.array_equal(torch.cat(ind).numpy(), g[f's{seed}_inds'].astype(np.int64))
    assert np.array_equal(torch.cat(tgt).numpy(), g[f's{seed}_targets'])                # fp32, bit for bit


@pytest.mark.parametrize('seed,kw', CASES + [(7, dict(B=4, H=320, W=416, G=(3, 40, 9, 17))), (8, dict(B=2, G=(1, 2)))])
def test_kernel_source_on_host_equals_oracle(seed, kw):
    points, boxes, labels, flags = case(seed, **kw)
    want, got = _oracle(points, boxes, labels, flags), _host(points, boxes, labels, flags)
    for w_list, g_list in zip(want, got):
        for w, g in zip(w_list, g_list):
            assert torch.equal(w, g)
    assert int((torch.cat(want[2]) >= 0).sum()) > 0


def test_image_without_ground_truth_and_many_ground_truths():
    points, boxes, labels, flags = case(11, B=3, H=128, W=160, G=(5, 5, 5))
    boxes[1], labels[1] = boxes[1][:0], labels[1][:0]                      # an empty image between two normal ones
    gen = torch.Generator().manual_seed(5)
    xy = torch.rand(1500, 2, generator=gen) * 120                          # more than the 1024 the kernel stages in smem
    boxes[2] = torch.cat([xy, xy + 4 + torch.rand(1500, 2, generator=gen) * 30], 1)
    labels[2] = torch.randint(0, 80, (1500,), generator=gen)
    want, got = _oracle(points, boxes, labels, flags), _host(points, boxes, labels, flags)
    for w_list, g_list in zip(want, got):
        for w, g in zip(w_list, g_list):
            assert torch.equal(w, g)
    n0 = points[0].shape[0]
    assert (got[0][0][n0:2 * n0] == 80).all() and (got[2][0][n0:2 * n0] == -1).all() and (got[1][0][n0:2 * n0] == 0).all()
    assert int(torch.cat(got[2]).max()) >= 5                               # indices are offsets into the concatenated list


def test_product_path_has_no_cpu_fallback():
    from boxinstseg_b200.models import build_head
    points, boxes, labels, flags = case(0)
    head = build_head(dict(type='CondInstBoxHead', num_classes=80, in_channels=256))
    with pytest.raises(RuntimeError):
        head.get_targets(points, boxes, labels)


@pytest.mark.parametrize('seed', range(6))
def test_boundary_heavy_boxes_on_the_host_build(seed):
    """Integer-valued boxes whose edges, centre-sampling windows and regress distances land exactly on locations and on the
    regress-range bounds (64, 128, 256, 512), duplicated boxes (area ties: the first one wins), degenerate (zero-area) boxes:
    every `> 0` / `>=` / `<=` / `<` of the reference is exercised at equality."""
    gen = torch.Generator().manual_seed(300 + seed)
    H, W = 256, 320
    sizes = [((H + s - 1) // s, (W + s - 1) // s) for s in CFG['strides']]
    points = oft.grid_points(sizes, CFG['strides'])
    boxes, labels = [], []
    for b in range(2):
        g = 14
        x1 = torch.randint(0, W // 8, (g,), generator=gen) * 4.0
        y1 = torch.randint(0, H // 8, (g,), generator=gen) * 4.0
        span = torch.tensor([8., 12., 64., 128., 136., 256.])
        bw = span[torch.randint(0, len(span), (g,), generator=gen)]
        bh = span[torch.randint(0, len(span), (g,), generator=gen)]
        bx = torch.stack([x1, y1, x1 + bw, y1 + bh], 1)
        bx[3] = bx[2]                                    # duplicate: equal area, equal geometry
        bx[5, 2:] = bx[5, :2]                            # zero-area box
        bx[7] = torch.tensor([4., 4., 132., 132.])       # max regress distance of the location (4, 4)... = 128 exactly
        boxes.append(bx)
        labels.append(torch.randint(0, 80, (g,), generator=gen))
    for flags in (dict(center_sampling=True, norm_on_bbox=True), dict(center_sampling=False, norm_on_bbox=False)):
        want, got = _oracle(points, boxes, labels, flags), _host(points, boxes, labels, flags)
        for w_list, g_list in zip(want, got):
            for w, g in zip(w_list, g_list):
                assert torch.equal(w, g)
        assert int((torch.cat(want[2]) >= 0).sum()) > 50

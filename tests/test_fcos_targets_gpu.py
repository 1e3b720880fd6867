"""f2 on the GPU: ``CondInstBoxHead.get_targets`` through the C ABI (``bxs_fcos_targets``, one launch per batch) against the
oracle restatement and the golden vectors of the reference's own methods -- labels, ground-truth indices and fp32 regression
targets bit for bit -- at the golden sizes, at config A's full size (2 x 800 x 1024: 17 064 locations per image) and on the
edge cases (image without ground truth, more ground truths than fit the shared-memory stage)."""
import numpy as np
import pytest
import torch

from oracle import fcos_targets as oft
from oracle.make_golden_fcos import CASES, CFG, case

pytestmark = pytest.mark.gpu


def _head(flags):
    from boxinstseg_b200.models import build_head
    return build_head(dict(type='CondInstBoxHead', num_classes=CFG['num_classes'], in_channels=256,
                           regress_ranges=CFG['regress_ranges'], strides=CFG['strides'], center_sample_radius=1.5, **flags))


def _run(points, boxes, labels, flags):
    dev = torch.device('cuda:0')
    head = _head(flags)
    out = head.get_targets([p.to(dev) for p in points], [b.to(dev) for b in boxes], [l.to(dev) for l in labels])
    torch.cuda.synchronize()
    return [[t.cpu() for t in lst] for lst in out]


def _oracle(points, boxes, labels, flags):
    return oft.get_targets(points, boxes, labels, CFG['regress_ranges'], CFG['strides'], CFG['num_classes'],
                           flags['center_sampling'], 1.5, flags['norm_on_bbox'])


@pytest.mark.parametrize('seed,kw', CASES)
def test_matches_reference_golden_and_oracle(golden, seed, kw):
    g = golden('fcos_targets')
    points, boxes, labels, flags = case(seed, **kw)
    lab, tgt, ind = _run(points, boxes, labels, flags)
    assert np.array_equal(torch.cat(lab).numpy(), g[f's{seed}_labels'].astype(np.int64))
    assert np.array_equal(torch.cat(ind).numpy(), g[f's{seed}_inds'].astype(np.int64))
    assert np.array_equal(torch.cat(tgt).numpy(), g[f's{seed}_targets'])
    for w_list, g_list in zip(_oracle(points, boxes, labels, flags), (lab, tgt, ind)):
        assert [tuple(w.shape) for w in w_list] == [tuple(x.shape) for x in g_list]


def test_config_a_full_size():
    points, boxes, labels, flags = case(21, B=2, H=800, W=1024, G=(8, 8))
    assert sum(p.shape[0] for p in points) == 17064
    want, got = _oracle(points, boxes, labels, flags), _run(points, boxes, labels, flags)
    for w_list, g_list in zip(want, got):
        for w, g in zip(w_list, g_list):
            assert torch.equal(w, g)
    assert int((torch.cat(got[2]) >= 0).sum()) > 50


def test_empty_image_and_many_ground_truths():
    points, boxes, labels, flags = case(11, B=3, H=128, W=160, G=(5, 5, 5))
    boxes[1], labels[1] = boxes[1][:0], labels[1][:0]
    gen = torch.Generator().manual_seed(5)
    xy = torch.rand(1500, 2, generator=gen) * 120
    boxes[2] = torch.cat([xy, xy + 4 + torch.rand(1500, 2, generator=gen) * 30], 1)
    labels[2] = torch.randint(0, 80, (1500,), generator=gen)
    want, got = _oracle(points, boxes, labels, flags), _run(points, boxes, labels, flags)
    for w_list, g_list in zip(want, got):
        for w, g in zip(w_list, g_list):
            assert torch.equal(w, g)


def test_batch_without_any_ground_truth_and_graph_capture():
    points, boxes, labels, flags = case(3, B=2, H=96, W=128, G=(4, 6))
    empty = _run(points, [b[:0] for b in boxes], [l[:0] for l in labels], flags)
    assert (torch.cat(empty[0]) == CFG['num_classes']).all() and (torch.cat(empty[2]) == -1).all()
    assert (torch.cat(empty[1]) == 0).all()
    # the launch takes its level table by value: capturable, and a replay reproduces the eager result
    dev = torch.device('cuda:0')
    head = _head(flags)
    pts, bx, lb = [p.to(dev) for p in points], [b.to(dev) for b in boxes], [l.to(dev) for l in labels]
    eager = head.get_targets(pts, bx, lb)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        head.get_targets(pts, bx, lb)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            cap = head.get_targets(pts, bx, lb)
    graph.replay()
    torch.cuda.synchronize()
    for e_list, c_list in zip(eager, cap):
        for e, c in zip(e_list, c_list):
            assert torch.equal(e, c)

"""f4, CPU side: DiscoBox's semantic-correspondence path (discobox_head.py:132-226, 347-411, 851-865, 1080-1096).
(1) the oracle restatement (oracle/corr.py) reproduces the golden vectors minted from the reference's own classes
(oracle/make_golden_corr.py); (2) the HOST build of the kernels' own source (csrc/corr_core.cuh through
tests/host_harness/corr_host.cpp) agrees with the oracle: ``solve`` to 5e-6 of the table's scale (row sums are taken in a
different order than ATen's reductions), the transfer to 1e-5; (3) the object bank of the product (plain torch) equals the
oracle's and the reference's retrievals.  The GPU twin is tests/test_corr_gpu.py."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import corr as oc
from oracle.make_golden_corr import BANK, FEAT, MASK, SOLVER, bank_case, case
from tests.helpers import host_twin


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _host_solve(Cu, h, w, dk, iters, smooth):
    lib = host_twin('corr_host')
    Cu = Cu.contiguous()
    T = torch.empty_like(Cu)
    assert lib.host_corr_solve(_p(Cu), _p(T), ctypes.c_int64(Cu.shape[0]), ctypes.c_int64(h), ctypes.c_int64(w), dk, iters, smooth) == 0
    return T


def _host_transfer(T, Cu, m0, m1, h, w):
    lib = host_twin('corr_host')
    K, Hm, Wm = m1.shape
    fg, bg = torch.empty(Hm, Wm), torch.empty(Hm, Wm)
    m0c, m1c = m0.reshape(-1).contiguous(), m1.contiguous()
    assert lib.host_corr_transfer(_p(T.contiguous()), _p(Cu.contiguous()), _p(m0c), _p(m1c), _p(fg), _p(bg), ctypes.c_int64(K),
                                  ctypes.c_int64(h), ctypes.c_int64(w), ctypes.c_int64(Hm), ctypes.c_int64(Wm)) == 0
    return fg, bg


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_oracle_reproduces_reference_golden(golden, seed):
    g = golden('corr')
    f0, f1, m0, m1 = case(seed)
    Cu = oc.cosine_table(f0, f1)
    assert np.array_equal(Cu.numpy(), g[f's{seed}_Cu'])
    T = oc.solve_votes(Cu, FEAT, FEAT, SOLVER['dist_kernel'], SOLVER['num_iter'], SOLVER['num_smooth_iter'])
    assert np.abs(T.numpy() - g[f's{seed}_T']).max() <= 2e-6 * np.abs(g[f's{seed}_T']).max()
    asg, fg, bg = oc.transfer(torch.from_numpy(g[f's{seed}_T']), Cu, m0, m1, FEAT, FEAT)
    assert np.abs(fg.numpy() - g[f's{seed}_fg']).max() <= 1e-6 and np.abs(bg.numpy() - g[f's{seed}_bg']).max() <= 1e-6
    assert abs(float(oc.nce_loss(Cu, asg)) - float(g[f's{seed}_nce'])) < 1e-6


@pytest.mark.parametrize('seed,shape,dk,iters,smooth', [(0, (7, 7), 9, 10, 1), (1, (7, 7), 9, 10, 1), (2, (7, 7), 9, 10, 1),
                                                        (3, (5, 8), 3, 4, 2), (4, (6, 4), 5, 0, 1), (5, (3, 3), 1, 3, 0)])
def test_solve_kernel_source_on_host(golden, seed, shape, dk, iters, smooth):
    h, w = shape
    f0, f1, _, _ = case(seed, K=3 if seed > 2 else 5, h=h, w=w)
    Cu = oc.cosine_table(f0, f1)
    want = oc.solve_votes(Cu, h, w, dk, iters, smooth)
    got = _host_solve(Cu, h, w, dk, iters, smooth)
    assert (got - want).abs().max() <= 5e-6 * want.abs().max()
    if seed <= 2:
        ref = torch.from_numpy(golden('corr')[f's{seed}_T'])
        assert (got - ref).abs().max() <= 5e-6 * ref.abs().max()


def test_no_round_is_the_windowed_table_and_one_round_matches():
    gen = torch.Generator().manual_seed(3)
    h, w = 5, 6
    T = torch.rand(2, h * w, h * w, generator=gen)
    want = oc.pass_message(T, h, w)
    out = _host_solve(T, h, w, 2 * max(h, w) + 1, 0, 0)
    assert torch.equal(out, T)                                              # no iteration: C = Cu * window (all ones)
    one = _host_solve(T, h, w, 2 * max(h, w) + 1, 1, 1)                     # C = (T + pm(T)/(rowsum+1e-4)) / (rowsum'+1e-4)
    votes = want / (want.sum(2, keepdim=True) + 1e-4)
    C = T + votes
    C = C / (C.sum(2, keepdim=True) + 1e-4)
    assert (one - C).abs().max() <= 2e-7 * C.abs().max()


@pytest.mark.parametrize('seed,shape,Hm', [(0, (7, 7), 28), (1, (7, 7), 28), (2, (7, 7), 28), (3, (5, 8), 20), (4, (4, 4), 9)])
def test_transfer_kernel_source_on_host(golden, seed, shape, Hm):
    h, w = shape
    f0, f1, m0, m1 = case(seed, K=5 if seed <= 2 else 3, h=h, w=w, Hm=Hm)
    Cu = oc.cosine_table(f0, f1)
    T = oc.solve_votes(Cu, h, w, 9, 10, 1)
    _, fg, bg = oc.transfer(T, Cu, m0, m1, h, w)
    got_fg, got_bg = _host_transfer(T, Cu, m0, m1, h, w)
    assert (got_fg - fg).abs().max() <= 1e-5 * max(1.0, float(fg.abs().max()))
    assert (got_bg - bg).abs().max() <= 1e-5 * max(1.0, float(bg.abs().max()))
    assert float(fg.max()) > 0.02 and float(bg.max()) > 0.02
    if seed <= 2:
        g = golden('corr')
        assert np.abs(got_fg.numpy() - g[f's{seed}_fg']).max() <= 2e-5 and np.abs(got_bg.numpy() - g[f's{seed}_bg']).max() <= 2e-5


@pytest.mark.parametrize('seed', [0, 1])
def test_object_bank_equals_oracle_and_reference_retrieval(golden, seed):
    """The product's bank (plain torch, any device) against the oracle's, and the retrieved slots against the golden indices of
    the reference's ObjectQueues.get_similar_obj: ring-buffer wrap-around, the four retrieval tests, the cap."""
    from boxinstseg_b200.models.dense_heads.disco_corr import ObjectQueues, create_one
    feats, masks, boxes = bank_case(seed)
    mine, orc = ObjectQueues(num_class=3, **BANK), oc.Queues(num_class=3, **BANK)
    for i in range(feats.shape[0]):
        assert mine.append(1, i, feats, masks, boxes) == orc.append(1, i, feats, masks, boxes)
    assert torch.equal(mine.queues[1].feature, orc.banks[1].feature) and torch.equal(mine.queues[1].mask, orc.banks[1].mask)
    assert torch.equal(mine.queues[1].box, orc.banks[1].box) and mine.queues[1].ptr == orc.banks[1].ptr == feats.shape[0] % BANK['len_queue']
    gen = torch.Generator().manual_seed(7 + seed)
    qf = oc.relu_and_l2_norm_feat(feats[:1] + 0.1 * torch.randn(1, feats.shape[1], FEAT, FEAT, generator=gen))
    q = create_one(masks[7:8].clone(), qf.clone(), boxes[7:8].clone(), 1)
    idx = mine.similar_indices(q)
    assert np.array_equal(idx.numpy(), golden('corr')[f'b{seed}_idx'])
    got = mine.get_similar_obj(q)
    assert torch.equal(got['mask'], orc.banks[1].mask[idx]) and got['category'] == 1
    assert mine.get_similar_obj(create_one(masks[7:8].clone(), qf.clone(), boxes[7:8].clone(), 2)) is None


def test_kernels_have_no_cpu_fallback():
    from boxinstseg_b200.models.dense_heads.disco_corr import SemanticCorrSolver
    s = SemanticCorrSolver(**SOLVER)
    with pytest.raises(RuntimeError):
        s.votes(torch.rand(2, 49, 49), 7, 7)
    with pytest.raises(RuntimeError):
        s.transfer(torch.rand(2, 49, 49), torch.rand(2, 49, 49), torch.rand(1, 28, 28), torch.rand(2, 28, 28), 7, 7)

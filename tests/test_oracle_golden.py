"""CPU: the oracle restatement against the golden vectors minted from the reference itself
(oracle/make_golden.py) and against known answers for the third-party pieces."""
import numpy as np
import pytest
import torch

from oracle import boxinst as ob
from oracle import levelset as ol
from oracle import tree as ot

T = torch.from_numpy


def _close(a, b, rtol=1e-5, atol=1e-6):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    assert a.shape == b.shape
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a - b).abs().max().item()


@pytest.mark.parametrize('tag', ['k3d2', 'k5d1'])
def test_pairwise(golden, tag):
    g = golden('pairwise_' + tag)
    k, d = int(g['size']), int(g['dilation'])
    x = T(g['logits']).requires_grad_(True)
    out = ob.pairwise_nlog(x, k, d)
    _close(out, g['out'])
    (gx,) = torch.autograd.grad((out * T(g['g_out'])).sum(), x)
    _close(gx, g['g_logits'], rtol=1e-4)
    # float64 truth
    x64 = T(g['logits']).double().requires_grad_(True)
    out64 = ob.pairwise_nlog(x64, k, d)
    _close(out64, g['out64'], rtol=1e-10, atol=1e-12)
    # out-of-image neighbours contribute exactly zero
    assert out64[0, 0, 0, 0].abs().item() < 1e-12


def test_projection(golden):
    g = golden('projection')
    s, t, soft = T(g['scores']), T(g['targets']), T(g['soft'])
    _close(ob.projection_losses(s, t).mean(), g['boxinst_mean'])
    _close(3.0 * ob.projection_losses(s, t), g['loss_w3'])
    _close(3.0 * ob.projection_losses(s, soft), g['loss_w3_soft'])
    _close(ol.disco_mil_loss(s[:, 0], t[:, 0]), g['disco_mil'])


def _boxinst_inputs(g):
    metas = [dict(img_shape=tuple(int(v) for v in g['img_shapes'][i]) + (3,),
                  ori_shape=tuple(int(v) for v in g['ori_shapes'][i]) + (3,),
                  img_norm_cfg=dict(mean=g['mean'], std=g['std'], to_rgb=True)) for i in range(2)]
    return T(g['img']), metas, [T(g['boxes0']), T(g['boxes1'])]


def test_boxinst_targets_and_loss(golden):
    g = golden('boxinst_loss')
    img, metas, boxes = _boxinst_inputs(g)
    sim, bm = ob.boxinst_targets(img, metas, boxes)
    for i in range(2):
        assert torch.equal(bm[i], T(g[f'bitmask{i}']))                       # index work: bit-exact
        _close(sim[i], g[f'sim{i}'], rtol=1e-5, atol=1e-7)
        assert torch.equal(sim[i] >= 0.3, T(g[f'sim{i}']) >= 0.3)           # thresholded weights
    gt_inds, img_inds = T(g['gt_inds']), T(g['img_inds'])
    logits = T(g['logits']).requires_grad_(True)
    prj, pair = ob.boxinst_mask_loss(logits, sim[img_inds], torch.cat(bm)[gt_inds][:, None],
                                     warmup_factor=float(g['warmup']))
    _close(prj, g['loss_prj'])
    _close(pair, g['loss_pairwise'])
    (gl,) = torch.autograd.grad(prj * float(g['g_prj']) + pair * float(g['g_pair']), logits)
    _close(gl, g['g_logits'], rtol=1e-4, atol=1e-7)


def test_condinst_head(golden):
    g = golden('condinst_head')
    feat = T(g['feat']).requires_grad_(True)
    params = T(g['params']).requires_grad_(True)
    out = ob.condinst_mask_head(feat, params, T(g['coors']), T(g['level_inds']), T(g['img_inds']))
    _close(out, g['out'], rtol=1e-4, atol=1e-5)
    gf, gp = torch.autograd.grad((out * T(g['g_out'])).sum(), [feat, params])
    _close(gf, g['g_feat'], rtol=1e-3, atol=1e-5)
    _close(gp, g['g_params'], rtol=1e-3, atol=1e-4)


def test_levelset(golden):
    g = golden('levelset')
    sc = T(g['scores']).requires_grad_(True)
    tg = T(g['target']).requires_grad_(True)
    m = T(g['mask'])
    phi = torch.cat([sc, 1 - sc], 1) * m
    loss = ol.levelset_loss(phi, tg * m, T(g['pixel_num']))
    _close(loss, g['loss'])
    gs, gt = torch.autograd.grad((loss * T(g['g_loss'])).sum(), [sc, tg])
    _close(gs, g['g_scores'], rtol=1e-4)
    _close(gt, g['g_target'], rtol=1e-4)
    _close(ol.length_regularization(sc.detach()), g['length'])


def test_lcm(golden):
    g = golden('lcm')
    phis = T(g['phis']).requires_grad_(True)
    loss = ol.lcm_loss(T(g['imgs']), phis, T(g['box']))
    _close(loss, g['loss'], rtol=1e-4)
    (gp,) = torch.autograd.grad(loss, phis)
    _close(gp, g['g_phis'], rtol=1e-3, atol=1e-7)


def test_meanfield(golden):
    g = golden('meanfield')
    k = ol.meanfield_kernel(T(g['feature']), 3, 0.5, 30.0, 2.0)
    _close(k, g['kernel'], rtol=1e-5, atol=1e-9)
    ps, va = ol.meanfield_forward(k, T(g['x']), T(g['targets']), 3, 10, 0.1)
    assert torch.equal(ps, T(g['pseudo'])) and torch.equal(va, T(g['valid']))


def test_tree(golden):
    g = golden('tree')
    gm = T(g['guide'])
    assert torch.equal(ot.grid_edge_weights(gm), T(g['edge_weight']))
    ei = ot.grid_edges(9, 13)
    for b in range(3):
        ids = ot.mst_edge_ids(ei, g['edge_weight'][b], 9 * 13)
        assert np.array_equal(ids, g['mst_edge_ids'][b])          # reference Boruvka edge set
        if ot.have_reference_boruvka():                          # live cross-check where _ref exists
            ref = ot.edges_to_ids(ot.mst_reference_boruvka(ei, g['edge_weight'][b], 9 * 13), 9, 13)
            assert np.array_equal(ids, ref)
    tree = ot.mst(gm)
    idx, par, chd = ot.bfs(tree)
    # BFS validity: permutation, root first, parents precede children, child lists consistent
    for b in range(3):
        i, p, c = idx[b].numpy(), par[b].numpy(), chd[b].numpy()
        assert sorted(i.tolist()) == list(range(9 * 13)) and i[0] == 0 and p[0] == 0
        assert np.all(p[1:] < np.arange(1, 9 * 13))
        for pos in range(9 * 13):
            kids = [k for k in c[pos] if k > 0]
            assert all(p[k] == pos for k in kids)
            assert len(kids) == int(np.sum(p[1:] == pos))
    _close(ot.build_edge_weight(gm, T(g['sorted_index']), T(g['sorted_parent']), True), g['w_low'])
    _close(ot.build_edge_weight(T(g['embed']), T(g['sorted_index']), T(g['sorted_parent']), False), g['w_high'])
    out = ot.tree_filter(T(g['feature']), gm, tree, low_tree=False)
    _close(out, g['filtered_high'], rtol=1e-4, atol=1e-6)


def test_tree_filter_gradients_match_closed_form():
    torch.manual_seed(3)
    gm = torch.randn(2, 2, 4, 5)
    tree = ot.mst(gm)
    x = torch.rand(2, 1, 4, 5, requires_grad=True)
    emb = (torch.randn(2, 3, 4, 5) * 0.5).requires_grad_(True)
    gout = torch.randn(2, 1, 4, 5)
    y = ot.tree_filter(x, emb, tree, low_tree=False, sigma=0.02)
    gx, ge = torch.autograd.grad((y * gout).sum(), [x, emb])
    # finite-difference truth on the float64 closed form
    def f(x_, e_):
        return (ot.tree_filter_dense(x_, e_, tree, low_tree=False) * gout.double()).sum().item()
    eps = 1e-4
    for (b, c, i, j) in [(0, 0, 1, 2), (1, 0, 3, 4)]:
        xp, xm = x.detach().clone(), x.detach().clone()
        xp[b, c, i, j] += eps
        xm[b, c, i, j] -= eps
        fd = (f(xp, emb.detach()) - f(xm, emb.detach())) / (2 * eps)
        assert abs(fd - gx[b, c, i, j].item()) < 2e-3 * max(1.0, abs(fd))
    for (b, c, i, j) in [(0, 1, 2, 2), (1, 2, 0, 0)]:
        ep, em = emb.detach().clone(), emb.detach().clone()
        ep[b, c, i, j] += eps
        em[b, c, i, j] -= eps
        fd = (f(x.detach(), ep) - f(x.detach(), em)) / (2 * eps)
        assert abs(fd - ge[b, c, i, j].item()) < 5e-3 * max(1.0, abs(fd))


def test_rgb2lab_known_answers_and_opencv():
    """third-party restatement pin: CIE known answers + OpenCV's independent implementation."""
    kat = {(255, 255, 255): (100.0, 0.0, 0.0), (0, 0, 0): (0.0, 0.0, 0.0),
           (255, 0, 0): (53.24, 80.09, 67.20), (0, 255, 0): (87.73, -86.18, 83.18),
           (0, 0, 255): (32.30, 79.19, -107.86)}
    for rgb, lab in kat.items():
        got = ob.rgb2lab_u8(np.array([[rgb]], dtype=np.uint8))[0, 0]
        assert np.allclose(got, lab, atol=0.02), (rgb, got)
    cv2 = pytest.importorskip('cv2')
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(32, 32, 3)).astype(np.uint8)
    ref = cv2.cvtColor((img / 255.0).astype(np.float32), cv2.COLOR_RGB2LAB)
    # OpenCV's float path uses an interpolated LUT (~0.3 LAB units); a coarse independent sanity check
    assert np.abs(ob.rgb2lab_u8(img) - ref).max() < 0.5


@pytest.mark.parametrize('seed,gamma', [(0, 0.01), (1, 0.5), (2, 2.0)])
def test_meanfield_with_inter_image_term_reproduces_reference_golden(golden, seed, gamma):
    """MeanField.forward(x, targets, inter_img_mask) (discobox_head.py:616-651) as corr_loss calls it."""
    import numpy as np
    from oracle import levelset as ol
    from oracle.make_golden_meanfield_inter import case
    g = golden('meanfield_inter')
    fm, x, t, iiu = case(seed)
    k = ol.meanfield_kernel(fm, 3, 0.5, 30.0, 2.0)
    ps, va = ol.meanfield_forward(k, x, t, 3, 10, 0.1, inter=iiu, gamma=gamma)
    want = np.unpackbits(g[f's{seed}_pseudo'], axis=-1)[..., :ps.shape[-1]]
    assert np.array_equal(ps.numpy().astype(np.uint8), want) and np.array_equal(va.numpy(), g[f's{seed}_valid'])

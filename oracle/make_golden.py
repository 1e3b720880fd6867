"""Mint golden vectors from the REFERENCE ITSELF (authoring container only).

Runs nowhere but where ``/root/reference`` exists.  The reference cannot be imported
(``mmcv``/``skimage`` are absent) so its own pure-PyTorch functions/classes are loaded by
AST-extracting their ``def``/``class`` nodes from the files where they lie and ``exec``-ing
them with trivial stubs for the mmcv decorators.  Nothing from the
reference is written into this repository except numeric input/output vectors
(``tests/golden/*.npz``).

    python -m oracle.make_golden            # regenerate + cross-check the oracle

Third-party stand-ins (absent from /root/reference, see oracle/__init__.py):
  * ``tensor2imgs``  -> OpenCV ``cv2.multiply``/``cv2.add`` exactly as mmcv.imdenormalize calls them
  * ``color.rgb2lab``-> oracle.boxinst.rgb2lab_u8 (restated scikit-image algorithm)
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get('BXS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def extract(relpath, names, extra_globals=None):
    src = open(os.path.join(REF, relpath)).read()
    tree = ast.parse(src)
    picked = [n for n in tree.body
              if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    assert not missing, f'{relpath}: missing {missing}'
    mod = ast.Module(body=picked, type_ignores=[])
    g = dict(torch=torch, nn=nn, F=F, np=np)
    g.update(extra_globals or {})
    exec(compile(mod, relpath, 'exec'), g)
    return types.SimpleNamespace(**{n: g[n] for n in names})


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def _force_fp32(*a, **k):
    return lambda fn: fn


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()


def cv2_tensor2imgs(tensor, mean, std, to_rgb=True):
    """mmcv.image.tensor2imgs restated with the same OpenCV calls mmcv.imdenormalize makes."""
    import cv2
    mean = np.array(mean, dtype=np.float32)
    std = np.array(std, dtype=np.float32)
    imgs = []
    for i in range(tensor.size(0)):
        img = tensor[i].cpu().numpy().transpose(1, 2, 0)
        img = np.ascontiguousarray(img, dtype=np.float32)
        m = mean.reshape(1, -1).astype(np.float64)
        s = std.reshape(1, -1).astype(np.float64)
        img = cv2.multiply(img, s)
        cv2.add(img, m, img)
        if to_rgb:
            cv2.cvtColor(img, cv2.COLOR_RGB2BGR, img)
        imgs.append(np.ascontiguousarray(img.astype(np.uint8)))
    return imgs


def load_reference():
    from oracle import boxinst as ob
    color = types.SimpleNamespace(rgb2lab=ob.rgb2lab_u8)
    stubs = dict(HEADS=_Registry(), LOSSES=_Registry(), MATCH_COST=_Registry(),
                 force_fp32=_force_fp32, BaseModule=_BaseModule, INF=1e8,
                 tensor2imgs=cv2_tensor2imgs, color=color)
    ch = extract('mmdet/models/dense_heads/condinst_head.py',
                 ['compute_pairwise_term', 'dice_coefficient', 'compute_project_term',
                  'aligned_bilinear', 'get_original_image', 'unfold_wo_center',
                  'get_image_color_similarity', 'CondInstMaskHead'], stubs)
    # the class body resolves module-level helpers through its own globals
    g = ch.CondInstMaskHead.loss.__globals__
    g['pairwise_nlog'] = ch.compute_pairwise_term
    ls = extract('mmdet/models/losses/levelset_loss.py',
                 ['LevelsetLoss', 'region_levelset', 'length_regularization', 'LCM',
                  'LocalConsistencyModule'], stubs)
    bp = extract('mmdet/models/losses/box_projection_loss.py', ['BoxProjectionLoss'], stubs)
    db = extract('mmdet/models/dense_heads/discobox_head.py',
                 ['MeanField', 'dice_loss', 'mil_loss'], stubs)
    tf = extract('mmdet/ops/tree_filter/modules/tree_filter.py',
                 ['MinimumSpanningTree', 'TreeFilter2D'],
                 dict(mst=None, bfs=None, refine=None, dist=None))
    mc = extract('mmdet/core/bbox/match_costs/match_cost.py', ['BoxMatchingCost'], stubs)
    return types.SimpleNamespace(ch=ch, ls=ls, bp=bp, db=db, tf=tf, mc=mc)


def _np(t):
    return t.detach().cpu().numpy()


def synth_image(gen, h, w, lowres=8):
    """Low-frequency field + noise, uint8-valued RGB [3,h,w] float32."""
    coarse = torch.rand(1, 3, max(h // lowres, 2), max(w // lowres, 2), generator=gen) * 255
    img = F.interpolate(coarse, size=(h, w), mode='bilinear', align_corners=False)[0]
    img = (img + torch.randn(3, h, w, generator=gen) * 8).clamp(0, 255).floor()
    return img


MEAN = [123.675, 116.28, 103.53]
STD = [58.395, 57.12, 57.375]


def main():
    from oracle import boxinst as ob
    from oracle import levelset as ol
    from oracle import tree as ot
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    gen = torch.Generator().manual_seed(20260923)
    report = {}

    def close(name, a, b, rtol=1e-5, atol=1e-6):
        a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
        err = ((a - b).abs() / (atol / rtol + b.abs())).max().item() if a.numel() else 0.0
        report[name] = err
        assert err <= rtol, f'{name}: oracle deviates from the reference, scaled err {err:.3e}'

    # ---- a7 pairwise (fwd + autograd bwd), two neighbourhoods, incl. extreme logits --------
    for tag, (k, d, shape) in dict(k3d2=(3, 2, (2, 1, 9, 11)), k5d1=(5, 1, (1, 1, 7, 6))).items():
        x = (torch.randn(shape, generator=gen) * 3).requires_grad_(True)
        with torch.no_grad():
            x[0, 0, 0, 0], x[0, 0, 0, 1], x[0, 0, 1, 0] = 30.0, -30.0, 60.0
        out = ref.ch.compute_pairwise_term(x, k, d)
        g = torch.rand(out.shape, generator=gen)
        (gx,) = torch.autograd.grad((out * g).sum(), x)
        x64 = x.detach().double().requires_grad_(True)
        out64 = ref.ch.compute_pairwise_term(x64, k, d)
        (gx64,) = torch.autograd.grad((out64 * g.double()).sum(), x64)
        mine = ob.pairwise_nlog(x.detach(), k, d)
        close(f'pairwise_{tag}', mine, out, rtol=1e-5, atol=1e-6)
        np.savez(os.path.join(OUT, f'pairwise_{tag}.npz'), size=k, dilation=d, logits=_np(x),
                 out=_np(out), g_out=_np(g), g_logits=_np(gx), out64=_np(out64), g_logits64=_np(gx64))

    # ---- a6 projection: BoxInst mean form, BoxProjectionLoss, BoxMatchingCost, disco MIL -----
    s = torch.rand(4, 1, 7, 9, generator=gen)
    t = torch.zeros(4, 1, 7, 9)
    t[0, 0, 1:5, 2:8] = 1
    t[1, 0, 0:7, 0:9] = 1
    t[2, 0, 3:4, 4:5] = 1            # instance 3 has an empty target
    soft = torch.rand(4, 1, 7, 9, generator=gen) * t
    prj_mean = ref.ch.compute_project_term(s, t)
    bpl = ref.bp.BoxProjectionLoss(loss_weight=3.0)
    prj_vec = bpl(s, t)
    prj_soft = bpl(s, soft)
    cost = ref.mc.BoxMatchingCost(weight=2.0, pred_act=False, eps=1.0)(s, t)
    mil = ref.db.mil_loss(ref.db.dice_loss, s[:, 0], None, t[:, 0])
    close('projection_mean', ob.projection_losses(s, t).mean(), prj_mean)
    close('projection_vec', 3.0 * ob.projection_losses(s, t), prj_vec)
    close('projection_soft', 3.0 * ob.projection_losses(s, soft), prj_soft)
    close('disco_mil', ol.disco_mil_loss(s[:, 0], t[:, 0]), mil)
    np.savez(os.path.join(OUT, 'projection.npz'), scores=_np(s), targets=_np(t), soft=_np(soft),
             boxinst_mean=_np(prj_mean), loss_w3=_np(prj_vec), loss_w3_soft=_np(prj_soft),
             match_cost_w2_eps1=_np(cost), disco_mil=_np(mil))

    # ---- a5 targets + a8 full BoxInst loss through the reference CondInstMaskHead ----------
    head = ref.ch.CondInstMaskHead(in_channels=8, in_stride=8, out_stride=4, topk_per_img=64,
                                   max_proposals=-1, boxinst_enabled=True)
    head._iter += 4999          # loss() adds one -> warm-up factor 0.5
    hp, wp = 48, 64
    metas, imgs = [], []
    for (ih, iw) in [(48, 64), (40, 56)]:
        raw = synth_image(gen, ih, iw)
        norm = (raw - torch.tensor(MEAN).view(3, 1, 1)) / torch.tensor(STD).view(3, 1, 1)
        imgs.append(F.pad(norm, (0, wp - iw, 0, hp - ih)))
        metas.append(dict(img_shape=(ih, iw, 3), ori_shape=(ih * 2, iw * 2, 3),
                          img_norm_cfg=dict(mean=np.array(MEAN, dtype=np.float32),
                                            std=np.array(STD, dtype=np.float32), to_rgb=True)))
    img = torch.stack(imgs)
    gt_bboxes = [torch.tensor([[3.2, 4.7, 40.9, 30.1], [20.0, 10.0, 63.9, 47.5], [5.5, 5.5, 7.9, 7.9]]),
                 torch.tensor([[0.0, 0.0, 55.0, 39.0], [30.3, 2.2, 50.8, 38.6]])]
    gt_inds = torch.tensor([0, 0, 1, 2, 3, 3, 4, 4, 4])
    img_inds = torch.tensor([0, 0, 0, 0, 1, 1, 1, 1, 1])
    logits = (torch.randn(9, 1, hp // 4, wp // 4, generator=gen) * 2).requires_grad_(True)
    sims, bitmasks, _ = head.get_targets(gt_bboxes, None, img, metas)
    losses = head.loss(img, metas, logits, gt_inds, gt_bboxes, None, None)
    (g_logits,) = torch.autograd.grad(losses['loss_prj'] * 0.7 + losses['loss_pairwise'] * 1.3, logits)
    o_sim, o_bm = ob.boxinst_targets(img, metas, gt_bboxes)
    for i in range(2):
        assert torch.equal(o_bm[i], bitmasks[i]), 'bitmask (index work) must be bit-exact'
        close(f'similarity_{i}', o_sim[i], sims[i][0], rtol=1e-5, atol=1e-7)
        assert torch.equal(o_sim[i] >= 0.3, sims[i][0] >= 0.3), 'thresholded weights must be bit-exact'
    bm_cat = torch.cat(o_bm)[gt_inds][:, None]
    o_prj, o_pair = ob.boxinst_mask_loss(logits.detach(), o_sim[img_inds], bm_cat, warmup_factor=0.5)
    close('boxinst_loss_prj', o_prj, losses['loss_prj'])
    close('boxinst_loss_pairwise', o_pair, losses['loss_pairwise'])
    np.savez(os.path.join(OUT, 'boxinst_loss.npz'), img=_np(img),
             img_shapes=np.array([m['img_shape'][:2] for m in metas]),
             ori_shapes=np.array([m['ori_shape'][:2] for m in metas]), mean=np.array(MEAN, np.float32),
             std=np.array(STD, np.float32), boxes0=_np(gt_bboxes[0]), boxes1=_np(gt_bboxes[1]),
             gt_inds=_np(gt_inds), img_inds=_np(img_inds), logits=_np(logits),
             sim0=_np(sims[0][0]), sim1=_np(sims[1][0]), bitmask0=_np(bitmasks[0]), bitmask1=_np(bitmasks[1]),
             loss_prj=_np(losses['loss_prj']), loss_pairwise=_np(losses['loss_pairwise']),
             warmup=0.5, g_prj=0.7, g_pair=1.3, g_logits=_np(g_logits))

    # ---- a1 dynamic mask head --------------------------------------------------------------
    feat = torch.randn(2, 8, 6, 8, generator=gen).requires_grad_(True)
    params = (torch.randn(9, head.num_gen_params, generator=gen) * 0.3).requires_grad_(True)
    coors = torch.rand(9, 2, generator=gen) * torch.tensor([64.0, 48.0])
    levels = torch.tensor([0, 1, 2, 3, 4, 0, 1, 2, 3])
    out = head(feat, params, coors, levels, img_inds)
    gq = torch.randn(out.shape, generator=gen)
    g_feat, g_params = torch.autograd.grad((out * gq).sum(), [feat, params])
    close('condinst_head', ob.condinst_mask_head(feat.detach(), params.detach(), coors, levels, img_inds), out,
          rtol=1e-4, atol=1e-5)
    np.savez(os.path.join(OUT, 'condinst_head.npz'), feat=_np(feat), params=_np(params), coors=_np(coors),
             level_inds=_np(levels), img_inds=_np(img_inds), out=_np(out), g_out=_np(gq),
             g_feat=_np(g_feat), g_params=_np(g_params))
    tt = torch.randn(2, 3, 5, 7, generator=gen)
    for f in (2, 4):
        close(f'aligned_bilinear_{f}', ob.aligned_bilinear(tt, f), ref.ch.aligned_bilinear(tt, f))

    # ---- a9 / a10 level set ----------------------------------------------------------------
    m = torch.zeros(3, 1, 6, 7)
    m[0, 0, 1:5, 1:6] = 1
    m[1, 0, :, :] = 1
    sc = torch.rand(3, 1, 6, 7, generator=gen).requires_grad_(True)
    tg = torch.randn(3, 3, 6, 7, generator=gen).requires_grad_(True)
    phi = torch.cat([sc, 1 - sc], 1) * m
    pix = m.sum((1, 2, 3)).clamp(min=1)
    lsl = ref.ls.LevelsetLoss(loss_weight=1.0)(phi, tg * m, pix)
    gv = torch.rand(3, generator=gen)
    g_sc, g_tg = torch.autograd.grad((lsl * gv).sum(), [sc, tg])
    close('levelset', ol.levelset_loss(phi.detach(), (tg * m).detach(), pix), lsl)
    close('length', ol.length_regularization(sc.detach()), ref.ls.length_regularization()(sc.detach()))
    np.savez(os.path.join(OUT, 'levelset.npz'), scores=_np(sc), target=_np(tg), mask=_np(m), pixel_num=_np(pix),
             loss=_np(lsl), g_loss=_np(gv), g_scores=_np(g_sc), g_target=_np(g_tg),
             length=_np(ref.ls.length_regularization()(sc.detach())))

    # ---- a11 LCM ---------------------------------------------------------------------------
    li = torch.rand(2, 3, 12, 12, generator=gen)
    lp = torch.rand(2, 1, 12, 12, generator=gen).requires_grad_(True)
    lb = torch.zeros(2, 1, 12, 12)
    lb[0, 0, 2:9, 3:11] = 1
    lb[1, 0, 0:6, 0:5] = 0.5
    lv = ref.ls.LCM(li, lp, lb)
    (g_lp,) = torch.autograd.grad(lv, lp)
    close('lcm', ol.lcm_loss(li, lp.detach(), lb), lv, rtol=1e-4)
    np.savez(os.path.join(OUT, 'lcm.npz'), imgs=_np(li), phis=_np(lp), box=_np(lb), loss=_np(lv), g_phis=_np(g_lp))

    # ---- a16 mean field (DiscoBox config: k3, iter 10, alpha0 2, theta0 .5, theta1 30, base .1)
    fm = torch.randn(1, 3, 10, 12, generator=gen)
    mf = ref.db.MeanField(fm, kernel_size=3, theta0=0.5, theta1=30, theta2=10, alpha0=2, iter=10, base=0.1)
    mx = torch.rand(4, 1, 10, 12, generator=gen)
    mt = torch.zeros(4, 1, 10, 12)
    mt[0, 0, 1:8, 2:10] = 1
    mt[1, 0, :, :] = 1
    mt[2, 0, 4:6, 4:6] = 1
    pseudo, valid = mf(mx, mt)
    ok = ol.meanfield_kernel(fm, 3, 0.5, 30.0, 2.0)
    close('meanfield_kernel', ok, mf.kernel[:, 0], rtol=1e-5, atol=1e-9)
    o_ps, o_va = ol.meanfield_forward(ok, mx, mt, 3, 10, 0.1)
    assert torch.equal(o_ps, pseudo) and torch.equal(o_va, valid), 'mean-field pseudo labels must be bit-exact'
    np.savez(os.path.join(OUT, 'meanfield.npz'), feature=_np(fm), kernel=_np(mf.kernel[:, 0]), x=_np(mx),
             targets=_np(mt), pseudo=_np(pseudo), valid=_np(valid))

    # ---- a12 MST (reference boruvka.cpp compiled from where it lies) + a14 edge weights -------
    ot.build()
    assert ot.have_reference_boruvka(), 'run `make -C oracle` with /root/reference present'
    gm = torch.randn(3, 3, 9, 13, generator=gen)
    gm[2] = torch.round(gm[2])           # heavy ties -> exercises the (weight, index) order
    MST = ref.tf.MinimumSpanningTree(ref.tf.TreeFilter2D.norm2_distance)
    ei = MST._build_matrix_index(gm)
    ew = MST._build_feature_weight(gm)
    assert np.array_equal(_np(ei[0]), ot.grid_edges(9, 13))
    assert torch.equal(ew, ot.grid_edge_weights(gm))
    ids = []
    for b in range(3):
        ref_edges = ot.mst_reference_boruvka(_np(ei[b]), _np(ew[b]), 9 * 13)
        ref_ids = ot.edges_to_ids(ref_edges, 9, 13)
        mine_ids = ot.mst_edge_ids(_np(ei[b]), _np(ew[b]), 9 * 13)
        assert np.array_equal(ref_ids, mine_ids), 'MST edge set must equal the reference Boruvka set'
        ids.append(ref_ids)
    tree = ot.mst(gm)
    idx, par, chd = ot.bfs(tree)
    emb = torch.randn(3, 5, 9, 13, generator=gen).requires_grad_(True)
    tf2d = ref.tf.TreeFilter2D(sigma=0.02)
    w_low = tf2d.build_edge_weight(gm, idx, par, True)
    w_high = tf2d.build_edge_weight(emb, idx, par, False)
    close('edge_weight_low', ot.build_edge_weight(gm, idx, par, True), w_low)
    close('edge_weight_high', ot.build_edge_weight(emb.detach(), idx, par, False), w_high)
    # tree filter (refine) itself is CUDA-only in the reference -> pinned on the closed form
    feat_in = torch.rand(3, 1, 9, 13, generator=gen)
    dense = ot.tree_filter_dense(feat_in, gm, tree, low_tree=False)
    fast = ot.tree_filter(feat_in, gm, tree, low_tree=False)
    close('tree_filter_closed_form', fast, dense, rtol=1e-4, atol=1e-6)
    np.savez(os.path.join(OUT, 'tree.npz'), guide=_np(gm), mst_edge_ids=np.stack(ids), edge_weight=_np(ew),
             embed=_np(emb), w_low=_np(w_low), w_high=_np(w_high), sorted_index=_np(idx), sorted_parent=_np(par),
             sorted_child=_np(chd), feature=_np(feat_in), filtered_high=_np(dense))

    for k, v in sorted(report.items()):
        print(f'{k:28s} scaled err vs reference {v:.2e}')
    print('golden vectors written to', os.path.normpath(OUT))


if __name__ == '__main__':
    sys.exit(main())

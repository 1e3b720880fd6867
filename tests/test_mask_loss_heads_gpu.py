"""GPU: the per-head mask-loss assemblies (a17) against compositions of the oracle that follow the
reference's own lines (box_solov2_head.py:334-367, box2mask_head.py:269-335, discobox_head.py:1266-1300), at toy sizes
AND at the sizes of BASELINE.json's configs (200x256 level maps; 256x256 predictions filtered at 96x96).
The tree filter's fp32 recursion depends on the (valid, but in the reference racy) BFS order, so the oracle's refine is
driven with the order the CUDA path used: then every loss and gradient agrees to <= 1e-3 (observed ~1e-5)."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _boxes(n, h, w):
    m = torch.zeros(n, h, w)
    for i in range(n):
        m[i, h // 6 + i: h // 6 + i + h // 2, w // 5: w // 5 + w // 2 + i] = 1
    return m


def test_boxlevelset_mask_loss_vs_oracle():
    from boxinstseg_b200.models import build_head
    from oracle import levelset as ol, tree as ot
    from oracle.boxinst import projection_losses
    gen = torch.Generator().manual_seed(0)
    h, w, n = 24, 32, 4
    ins_pred = torch.randn(n, h, w, generator=gen)
    box = _boxes(n, h, w)
    img2 = torch.randn(2, 3, h, w, generator=gen)
    lst2 = torch.randn(2, 5, h, w, generator=gen) * 0.3
    img_t = img2.repeat_interleave(2, 0)
    lst_t = lst2.repeat_interleave(2, 0)

    # ---- oracle composition (fp32, CPU)
    p = ins_pred.clone().requires_grad_(True)
    l = lst_t.clone().requires_grad_(True)
    s = torch.sigmoid(p.unsqueeze(1))
    b = box.unsqueeze(1)
    lp = 3.0 * projection_losses(s, b)
    phi = torch.cat((s, 1 - s), 1) * b
    pix = b.sum((1, 2, 3)).clamp(min=1)
    l_img = ol.levelset_loss(phi, img_t * b, pix) * 0.05
    f1 = _oracle_tf_same_order(s, img_t, img_t, True)
    f2 = _oracle_tf_same_order(f1, l, l, False)
    l_feat = ol.levelset_loss(phi, torch.cat((f1, f2), 1) * b, pix) * 5.0
    ref_prj, ref_ls = lp.mean(), (l_img + l_feat).mean()
    gp_ref, gl_ref = torch.autograd.grad(ref_prj + ref_ls, [p, l])

    head = build_head(dict(type='BoxSOLOv2Head', num_classes=80, in_channels=256,
                           loss_boxpro=dict(type='BoxProjectionLoss', loss_weight=3.0),
                           loss_levelset=dict(type='LevelsetLoss', loss_weight=1.0)))
    pg = ins_pred.to(DEV).requires_grad_(True)
    lg = lst_t.to(DEV).requires_grad_(True)
    out = head.mask_loss([pg], [box.to(DEV)], [img_t.to(DEV)], [lg])
    gp, gl = torch.autograd.grad(out['loss_boxpro'] + out['loss_levelset'], [pg, lg])
    assert abs(out['loss_boxpro'].item() - ref_prj.item()) < 1e-4 * abs(ref_prj.item())
    assert abs(out['loss_levelset'].item() - ref_ls.item()) < 1e-3 * abs(ref_ls.item())
    assert rel_err(gp.cpu(), gp_ref) < 1e-3 and rel_err(gl.cpu(), gl_ref) < 1e-3
    # de-duplicated trees == per-instance trees
    out2 = head.mask_loss([pg], [box.to(DEV)], [img_t.to(DEV)], [lg], shared_trees=False)
    assert torch.allclose(out2['loss_levelset'], out['loss_levelset'], rtol=1e-6)


def test_box2mask_mask_loss_single_vs_oracle():
    from boxinstseg_b200.models import build_head
    from oracle import levelset as ol, tree as ot
    from oracle.boxinst import projection_losses
    gen = torch.Generator().manual_seed(1)
    h, w, n = 40, 48, 3
    preds = torch.randn(n, h, w, generator=gen)
    targets = _boxes(n, 4 * h, 4 * w)
    norm_img = torch.randn(2, 3, 4 * h, 4 * w, generator=gen)
    lst_feat = torch.randn(2, 1, h, w, generator=gen)
    num = [2, 1]
    rs = lambda t: F.interpolate(t, (h, w), mode='bilinear', align_corners=False)     # noqa: E731
    s96 = lambda t: F.interpolate(t, (96, 96), mode='bilinear', align_corners=False)  # noqa: E731

    p = preds.clone().requires_grad_(True)
    img = rs(norm_img)
    lst = rs(lst_feat)
    rep = torch.tensor(num)
    img_t, lst_t = img.repeat_interleave(rep, 0), lst.repeat_interleave(rep, 0)
    box = rs(targets.unsqueeze(1))
    s = torch.sigmoid(p.unsqueeze(1))
    l_prj = (5.0 * projection_losses(s, box)).mean()
    phi = torch.cat((s, 1 - s), 1) * box
    pix = box.sum((1, 2, 3)).clamp(min=1)
    l_img = ol.levelset_loss(phi, img_t * box, pix).mean() * 0.05
    f1 = _oracle_tf_same_order(s96(s), s96(img_t), s96(img_t), True)
    f2 = _oracle_tf_same_order(f1, s96(lst_t), s96(lst_t), False)
    deep = torch.cat((rs(f1), rs(f2)), 1) * box
    l_feat = ol.levelset_loss(phi, deep, pix).mean() * 5.0
    l_lcm = 0.2 * ol.lcm_loss(s96(img_t), s96(s), s96(box))
    ref_ls = l_img + l_feat + l_lcm
    (g_ref,) = torch.autograd.grad(l_prj + ref_ls, p)

    head = build_head(dict(type='Box2MaskHead', num_queries=100))
    pg = preds.to(DEV).requires_grad_(True)
    prj, ls = head.mask_loss_single(pg, targets.to(DEV), num, norm_img.to(DEV), lst_feat.to(DEV))
    (g,) = torch.autograd.grad(prj + ls, pg)
    assert abs(prj.item() - l_prj.item()) < 1e-4 * abs(l_prj.item())
    assert abs(ls.item() - ref_ls.item()) < 1e-3 * abs(ref_ls.item())
    assert rel_err(g.cpu(), g_ref) < 1e-3


def test_discobox_mask_loss_vs_oracle():
    from boxinstseg_b200.models import build_head
    from oracle import levelset as ol
    gen = torch.Generator().manual_seed(2)
    h, w = 64, 80
    color = F.interpolate(torch.randn(2, 3, 9, 11, generator=gen), size=(h, w), mode='bilinear', align_corners=True)
    s_pred = torch.randn(5, h, w, generator=gen)
    target = _boxes(5, h, w)
    target[4] = 0                                                  # all-zero target is dropped (:1283-1287)
    img_inds = torch.tensor([0., 0., 1., 1., 1.])
    cfg = dict(type='DiscoBoxSOLOv2Head', num_classes=80, in_channels=256, loss_ins=dict(loss_weight=1.0),
               loss_ts=dict(loss_weight=1.0, alpha0=2.0, theta0=0.5, theta1=30.0, theta2=20.0, kernel=3, base=0.10, max_iter=10))
    p = s_pred.clone().requires_grad_(True)
    keep = target.flatten(1).sum(1) > 0
    s = torch.sigmoid(p)[keep]
    tg, ii = target[keep], img_inds[keep]
    l_ins = ol.disco_mil_loss(s, tg).mean()
    enl = F.max_pool2d(tg.unsqueeze(1), 3, 1, 1).squeeze(1)
    ts = []
    for b in range(2):
        sel = ii == b
        k = ol.meanfield_kernel(color[b:b + 1], 3, 0.5, 30.0, 2.0)
        pseudo, _ = ol.meanfield_forward(k, s[sel].detach().unsqueeze(1), tg[sel].unsqueeze(1), 3, 10, 0.1)
        ts.append(ol.disco_dice_loss(s[sel] * enl[sel], pseudo))
    l_ts = torch.cat(ts).mean()
    (g_ref,) = torch.autograd.grad(l_ins + l_ts, p)

    head = build_head(cfg)
    pg = s_pred.to(DEV).requires_grad_(True)
    out = head.mask_loss([pg], [target.to(DEV)], [img_inds.to(DEV)], color.to(DEV))
    (g,) = torch.autograd.grad(out['loss_ins'] + out['loss_ts'], pg)
    assert abs(out['loss_ins'].item() - l_ins.item()) < 1e-4 * abs(l_ins.item())
    assert abs(out['loss_ts'].item() - l_ts.item()) < 1e-3 * abs(l_ts.item())
    assert rel_err(g.cpu(), g_ref) < 1e-3


# ------------------------------------------------------------------ full-size assemblies, oracle on the same BFS order
def _oracle_tf_same_order(feature, embed, guide_for_tree, low):
    """oracle tree filter (C restatement of refine.cu) on OUR tree and BFS order for `guide_for_tree` (CPU tensors)."""
    from boxinstseg_b200.ops.tree_filter import MinimumSpanningTree, TreeFilter2D, bfs
    from oracle import tree as ot
    tree = MinimumSpanningTree(TreeFilter2D.norm2_distance)(guide_for_tree.detach().to(DEV))
    idx, par, chd = (t.cpu() for t in bfs(tree, 4))
    w = ot.build_edge_weight(embed, idx, par, low)
    shape = feature.shape
    return ot.refine(feature.reshape(shape[0], shape[1], -1), w, idx, par, chd, low).reshape(shape)


def test_boxlevelset_mask_loss_full_size_grouped_and_oracle():
    """Config D level 0/1 size (200x256), 2 images x 2 instances: the grouped path (trees once per image, no host sync)
    equals the per-instance path, and both equal the oracle composition of box_solov2_head.py:334-367."""
    from boxinstseg_b200.models import build_head
    from oracle import levelset as ol
    from oracle.boxinst import projection_losses
    gen = torch.Generator().manual_seed(5)
    h, w, n = 200, 256, 4
    ins_pred = torch.randn(n, h, w, generator=gen) * 2
    box = _boxes(n, h, w)
    img2 = F.interpolate(torch.randn(2, 3, 25, 32, generator=gen), size=(h, w), mode='bilinear') + 0.05 * torch.randn(2, 3, h, w, generator=gen)
    lst2 = torch.randn(2, 5, h, w, generator=gen) * 0.3
    inst = torch.tensor([0, 0, 1, 1])
    # oracle
    p = ins_pred.clone().requires_grad_(True)
    l = lst2.clone().requires_grad_(True)
    img_t, lst_t = img2[inst], l[inst]
    s = torch.sigmoid(p.unsqueeze(1))
    b = box.unsqueeze(1)
    phi = torch.cat((s, 1 - s), 1) * b
    pix = b.sum((1, 2, 3)).clamp(min=1)
    f1 = _oracle_tf_same_order(s, img_t, img_t, True)
    f2 = _oracle_tf_same_order(f1, lst_t, lst_t, False)
    ref_prj = (3.0 * projection_losses(s, b)).mean()
    ref_ls = (ol.levelset_loss(phi, img_t * b, pix) * 0.05 + ol.levelset_loss(phi, torch.cat((f1, f2), 1) * b, pix) * 5.0).mean()
    gp_ref, gl_ref = torch.autograd.grad(ref_prj + ref_ls, [p, l])
    head = build_head(dict(type='BoxSOLOv2Head', num_classes=80, in_channels=256))
    res = []
    for grouped in (True, False):
        pg = ins_pred.to(DEV).requires_grad_(True)
        lg = lst2.to(DEV).requires_grad_(True)
        if grouped:
            out = head.mask_loss([pg], [box.to(DEV)], [img2.to(DEV)], [lg], inst_imgs=[inst.to(DEV, torch.int32)])
        else:
            out = head.mask_loss([pg], [box.to(DEV)], [img2.to(DEV)[inst.to(DEV)]], [lg[inst.to(DEV)]])
        gp, gl = torch.autograd.grad(out['loss_boxpro'] + out['loss_levelset'], [pg, lg])
        res.append((out, gp, gl))
        assert abs(out['loss_boxpro'].item() - ref_prj.item()) <= 1e-4 * abs(ref_prj.item())
        assert abs(out['loss_levelset'].item() - ref_ls.item()) <= 1e-3 * abs(ref_ls.item())
        assert rel_err(gp.cpu(), gp_ref) <= 1e-3 and rel_err(gl.cpu(), gl_ref) <= 1e-3
    assert rel_err(res[0][1], res[1][1]) <= 1e-5 and rel_err(res[0][2], res[1][2]) <= 1e-5


def test_box2mask_mask_loss_single_full_size_vs_oracle():
    """Config E sizes: 256x256 predictions, 1024x1024 image and box masks, tree filter / LCM at 96x96 (box2mask_head.py:229-335)."""
    from boxinstseg_b200.models import build_head
    from oracle import levelset as ol
    from oracle.boxinst import projection_losses
    gen = torch.Generator().manual_seed(6)
    h = w = 256
    num = [2, 1]
    n = sum(num)
    preds = torch.randn(n, h, w, generator=gen) * 2
    targets = _boxes(n, 1024, 1024)
    norm_img = F.interpolate(torch.randn(2, 3, 32, 32, generator=gen), size=(1024, 1024), mode='bilinear') + 0.05 * torch.randn(2, 3, 1024, 1024, generator=gen)
    lst_feat = torch.randn(2, 1, h, w, generator=gen)
    rs = lambda t: F.interpolate(t, (h, w), mode='bilinear', align_corners=False)     # noqa: E731
    s96 = lambda t: F.interpolate(t, (96, 96), mode='bilinear', align_corners=False)  # noqa: E731
    p = preds.clone().requires_grad_(True)
    lf = lst_feat.clone().requires_grad_(True)
    rep = torch.tensor(num)
    img, lst = rs(norm_img), rs(lf)
    img_t, lst_t = img.repeat_interleave(rep, 0), lst.repeat_interleave(rep, 0)
    box = rs(targets.unsqueeze(1))
    s = torch.sigmoid(p.unsqueeze(1))
    l_prj = (5.0 * projection_losses(s, box)).mean()
    phi = torch.cat((s, 1 - s), 1) * box
    pix = box.sum((1, 2, 3)).clamp(min=1)
    l_img = ol.levelset_loss(phi, img_t * box, pix).mean() * 0.05
    f1 = _oracle_tf_same_order(s96(s), s96(img_t), s96(img_t), True)
    f2 = _oracle_tf_same_order(f1, s96(lst_t), s96(lst_t), False)
    l_feat = ol.levelset_loss(phi, torch.cat((rs(f1), rs(f2)), 1) * box, pix).mean() * 5.0
    ref_ls = l_img + l_feat + 0.2 * ol.lcm_loss(s96(img_t), s96(s), s96(box))
    g_ref, gl_ref = torch.autograd.grad(l_prj + ref_ls, [p, lf])
    head = build_head(dict(type='Box2MaskHead', num_queries=100))
    pg = preds.to(DEV).requires_grad_(True)
    lg = lst_feat.to(DEV).requires_grad_(True)
    prj, ls = head.mask_loss_single(pg, targets.to(DEV), num, norm_img.to(DEV), lg)
    g, gl = torch.autograd.grad(prj + ls, [pg, lg])
    assert abs(prj.item() - l_prj.item()) <= 1e-4 * abs(l_prj.item())
    assert abs(ls.item() - ref_ls.item()) <= 1e-3 * abs(ref_ls.item())
    assert rel_err(g.cpu(), g_ref) <= 1e-3 and rel_err(gl.cpu(), gl_ref) <= 1e-3
    # the image trees do not depend on the decoder layer: passing them in changes nothing
    prj2, ls2 = head.mask_loss_single(pg, targets.to(DEV), num, norm_img.to(DEV), lg, trees=head.image_trees(norm_img.to(DEV), (h, w)))
    assert torch.equal(prj, prj2) and torch.equal(ls, ls2)


def test_discobox_mask_loss_full_size_vs_oracle():
    """Config C size (200x256), 2 images, an all-zero target in the batch (kept with weight 0 instead of removed)."""
    from boxinstseg_b200.models import build_head
    from oracle import levelset as ol
    gen = torch.Generator().manual_seed(7)
    h, w = 200, 256
    color = F.interpolate(torch.randn(2, 3, 25, 32, generator=gen), size=(h, w), mode='bilinear', align_corners=True)
    s_pred = torch.randn(6, h, w, generator=gen)
    target = _boxes(6, h, w)
    target[3] = 0
    img_inds = torch.tensor([0, 0, 0, 1, 1, 1])
    p = s_pred.clone().requires_grad_(True)
    keep = target.flatten(1).sum(1) > 0
    s = torch.sigmoid(p)[keep]
    tg, ii = target[keep], img_inds[keep]
    l_ins = ol.disco_mil_loss(s, tg).mean()
    enl = F.max_pool2d(tg.unsqueeze(1), 3, 1, 1).squeeze(1)
    ts = []
    for b in range(2):
        sel = ii == b
        k = ol.meanfield_kernel(color[b:b + 1], 3, 0.5, 30.0, 2.0)
        pseudo, _ = ol.meanfield_forward(k, s[sel].detach().unsqueeze(1), tg[sel].unsqueeze(1), 3, 10, 0.1)
        ts.append(ol.disco_dice_loss(s[sel] * enl[sel], pseudo))
    l_ts = torch.cat(ts).mean()
    (g_ref,) = torch.autograd.grad(l_ins + l_ts, p)
    head = build_head(dict(type='DiscoBoxSOLOv2Head', num_classes=80, in_channels=256))
    pg = s_pred.to(DEV).requires_grad_(True)
    out = head.mask_loss([pg], [target.to(DEV)], [img_inds.to(DEV, torch.int32)], color.to(DEV))
    (g,) = torch.autograd.grad(out['loss_ins'] + out['loss_ts'], pg)
    assert abs(out['loss_ins'].item() - l_ins.item()) <= 1e-4 * abs(l_ins.item())
    assert abs(out['loss_ts'].item() - l_ts.item()) <= 1e-3 * abs(l_ts.item())
    assert rel_err(g.cpu(), g_ref) <= 1e-3
    assert float(g[3].abs().max()) == 0.0                       # the dropped instance receives no gradient


def test_multi_level_lanes_equal_single_level_calls_and_graph_replay():
    """The FPN levels of BoxSOLOv2Head.mask_loss run on lane streams (ops/_streams.py): the multi-level call must equal the
    levels evaluated one call at a time (losses and every gradient, bit for bit -- same kernels, only the streams differ),
    eagerly AND as a CUDA-graph replay of forward + backward captured with the forks and joins inside."""
    from boxinstseg_b200.models import build_head
    gen = torch.Generator().manual_seed(7)
    sizes = [(40, 48), (40, 48), (24, 32), (12, 16)]
    n_img, per_img = 2, 3
    head = build_head(dict(type='BoxSOLOv2Head', num_classes=80, in_channels=256,
                           loss_boxpro=dict(type='BoxProjectionLoss', loss_weight=3.0),
                           loss_levelset=dict(type='LevelsetLoss', loss_weight=1.0)))
    inst = torch.arange(n_img, dtype=torch.int32).repeat_interleave(per_img).to(DEV)
    preds = [torch.randn(n_img * per_img, h, w, generator=gen).to(DEV).requires_grad_(True) for h, w in sizes]
    boxes = [_boxes(n_img * per_img, h, w).to(DEV) for h, w in sizes]
    img_t = [torch.randn(n_img, 3, h, w, generator=gen).to(DEV) for h, w in sizes]
    lst_t = [(torch.randn(n_img, 5, h, w, generator=gen) * 0.3).to(DEV).requires_grad_(True) for h, w in sizes]
    leaves = preds + lst_t

    def step():
        out = head.mask_loss(preds, boxes, img_t, lst_t, inst_imgs=[inst] * len(sizes))
        return [out['loss_boxpro'], out['loss_levelset']] + list(torch.autograd.grad(out['loss_boxpro'] + out['loss_levelset'],
                                                                                     leaves))

    got = [t.clone() for t in step()]
    # one level per call: no lanes involved (a single level runs on the calling stream)
    prj, ls, grads = [], [], [None] * len(leaves)
    for k in range(len(sizes)):
        o = head.mask_loss([preds[k]], [boxes[k]], [img_t[k]], [lst_t[k]], inst_imgs=[inst])
        prj.append(o['loss_boxpro']); ls.append(o['loss_levelset'])
    # the multi-level losses are means over the concatenated per-level vectors: equal level sizes -> mean of the means
    want_prj, want_ls = torch.stack(prj).mean(), torch.stack(ls).mean()
    g = torch.autograd.grad(want_prj + want_ls, leaves)
    assert torch.allclose(got[0], want_prj, rtol=1e-6, atol=0) and torch.allclose(got[1], want_ls, rtol=1e-6, atol=0)
    for a, b in zip(got[2:], g):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-9)
    # CUDA-graph capture of the whole step, replayed twice
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(2):
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip(out, got):
            assert torch.equal(a, b)

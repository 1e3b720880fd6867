"""GPU: the per-head mask-loss assemblies (a17) against compositions of the oracle that follow the
reference's own lines (box_solov2_head.py:334-367, box2mask_head.py:269-335, discobox_head.py:1266-1300)."""
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
    f1 = ot.tree_filter(s, img_t, ot.mst(img_t))
    f2 = ot.tree_filter(f1, l, ot.mst(l.detach()), low_tree=False)
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
    assert rel_err(gp.cpu(), gp_ref) < 2e-3 and rel_err(gl.cpu(), gl_ref) < 5e-3
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
    t_img, t_lst = ot.mst(s96(img)), ot.mst(s96(lst))
    f1 = ot.tree_filter(s96(s), s96(img_t), t_img.repeat_interleave(rep, 0))
    f2 = ot.tree_filter(f1, s96(lst_t), t_lst.repeat_interleave(rep, 0), low_tree=False)
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
    assert rel_err(g.cpu(), g_ref) < 5e-3


def test_discobox_mask_loss_vs_oracle():
    from boxinstseg_b200.models import build_head
    from oracle import levelset as ol
    gen = torch.Generator().manual_seed(2)
    h, w = 32, 40
    color = F.interpolate(torch.randn(2, 3, 5, 6, generator=gen), size=(h, w), mode='bilinear', align_corners=True)
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
    assert abs(out['loss_ts'].item() - l_ts.item()) < 2e-3 * abs(l_ts.item())      # a flipped tie pixel in the pseudo label
    assert rel_err(g.cpu(), g_ref) < 5e-3

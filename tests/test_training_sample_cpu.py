"""f2: CondInstMaskHead.training_sample -- the vectorised sampling against the restated loops of
condinst_head.py:1166-1232 (oracle.boxinst.training_sample_topk).  Pure index work: bit-exact.  Runs on CPU tensors (the
method is plain torch; no kernel is involved)."""
import pytest
import torch


@pytest.mark.parametrize('seed,imgs,gts,topk', [(0, 2, 8, 64), (1, 3, 5, 16), (2, 1, 1, 64), (3, 4, 30, 64), (4, 2, 3, 2)])
def test_training_sample_matches_reference_loops(seed, imgs, gts, topk):
    from boxinstseg_b200.models import build_head
    from oracle.boxinst import training_sample_topk
    gen = torch.Generator().manual_seed(seed)
    head = build_head(dict(type='CondInstMaskHead', in_channels=16, boxinst_enabled=True, max_proposals=-1, topk_per_img=topk))
    sizes = [(10, 12), (5, 6), (3, 3)]
    ncls = 7
    cls_scores = [torch.randn(imgs, ncls, h, w, generator=gen) for h, w in sizes]
    ctrs = [torch.randn(imgs, 1, h, w, generator=gen) for h, w in sizes]
    params = [torch.randn(imgs, head.num_gen_params, h, w, generator=gen) for h, w in sizes]
    total = sum(imgs * h * w for h, w in sizes)
    # flattening order of the reference: per level, (image, y, x)
    img_inds = torch.cat([torch.arange(imgs).repeat_interleave(h * w) for h, w in sizes])
    level_inds = torch.cat([torch.full((imgs * h * w,), i) for i, (h, w) in enumerate(sizes)])
    coors = torch.rand(total, 2, generator=gen) * 100
    gt_inds = torch.randint(-1, gts, (total,), generator=gen)
    gt_inds[torch.rand(total, generator=gen) < 0.5] = -1
    gt_inds = torch.where(gt_inds >= 0, gt_inds + img_inds * gts, gt_inds)        # indices into the concatenated GT list
    out = head.training_sample(cls_scores, ctrs, params, coors, level_inds, img_inds, gt_inds)
    pos = gt_inds != -1
    flat_cls = torch.cat([c.permute(0, 2, 3, 1).flatten(end_dim=2) for c in cls_scores])[pos]
    flat_ctr = torch.cat([c.permute(0, 2, 3, 1).reshape(-1) for c in ctrs])[pos]
    flat_par = torch.cat([p.permute(0, 2, 3, 1).flatten(end_dim=2) for p in params])[pos]
    sel = training_sample_topk(flat_cls, flat_ctr, img_inds[pos], gt_inds[pos], topk)
    assert torch.equal(out[0], flat_par[sel]) and torch.equal(out[1], coors[pos][sel])
    assert torch.equal(out[2], level_inds[pos][sel]) and torch.equal(out[3], img_inds[pos][sel])
    assert torch.equal(out[4], gt_inds[pos][sel])

"""f3 on the GPU: the tail of CondInst.forward_train (detectors/condinst.py:66-75) through ``mask_branch_step`` equals the
three calls made by hand; the part from ``mask_head(...)`` on (head forward + fused loss + backward + parse_losses) is captured
in a CUDA graph and replayed -- no device->host read inside -- with bit-identical losses and gradients."""
import pytest
import torch

from tests.helpers import boxinst_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _inputs(head, case, imgs, gts_per_img, seed):
    gen = torch.Generator().manual_seed(seed)
    sizes = [(25, 32), (13, 16), (7, 8)]
    cls_scores = [torch.randn(imgs, 5, h, w, generator=gen).to(DEV) for h, w in sizes]
    ctrs = [torch.randn(imgs, 1, h, w, generator=gen).to(DEV) for h, w in sizes]
    params = [(torch.randn(imgs, head.num_gen_params, h, w, generator=gen) * 0.3).to(DEV) for h, w in sizes]
    total = sum(imgs * h * w for h, w in sizes)
    img_inds = torch.cat([torch.arange(imgs).repeat_interleave(h * w) for h, w in sizes])
    level_inds = torch.cat([torch.full((imgs * h * w,), i) for i, (h, w) in enumerate(sizes)])
    coors = torch.rand(total, 2, generator=gen) * torch.tensor([256.0, 200.0])
    gt_inds = torch.randint(0, gts_per_img, (total,), generator=gen) + img_inds * gts_per_img
    gt_inds[torch.rand(total, generator=gen) < 0.7] = -1
    return cls_scores, ctrs, params, coors.to(DEV), level_inds.to(DEV), img_inds.to(DEV), gt_inds.to(DEV)


def test_mask_branch_step_and_graph_capture():
    from boxinstseg_b200.models import build_head, mask_branch_step, parse_losses
    imgs, gts = 2, 3
    case = boxinst_case(21, B=imgs, hp=200, wp=256, gts_per_img=gts, inst_per_gt=1)
    head = build_head(dict(type='CondInstMaskHead', in_channels=16, in_stride=8, out_stride=4, topk_per_img=12, max_proposals=-1,
                           boxinst_enabled=True)).to(DEV)
    cls_scores, ctrs, params, coors, level_inds, img_inds, gt_inds = _inputs(head, case, imgs, gts, 0)
    mask_feat = torch.randn(imgs, 16, 25, 32, device=DEV).requires_grad_(True)
    for p in params:
        p.requires_grad_(True)
    img = case['img'].to(DEV)
    boxes = [b.to(DEV) for b in case['gt_bboxes']]

    head._iter.fill_(5000)
    losses = mask_branch_step(head, mask_feat, cls_scores, ctrs, params, coors, level_inds, img_inds, gt_inds, img, case['metas'],
                              boxes)
    # by hand, condinst.py:69-75
    head._iter.fill_(5000)
    sampled = head.training_sample(cls_scores, ctrs, params, coors, level_inds, img_inds, gt_inds)
    pred = head(mask_feat, *sampled[:4])
    want = head.loss(img, case['metas'], pred, sampled[4], boxes, None, None)
    assert set(losses) == {'loss_prj', 'loss_pairwise'}
    for k in want:
        assert torch.equal(losses[k], want[k])
    loss, log_vars = parse_losses(losses)
    assert torch.equal(loss, losses['loss_prj'] + losses['loss_pairwise'])
    assert log_vars["loss"] == pytest.approx(float(loss.detach())) and list(log_vars.keys()) == ['loss_prj', 'loss_pairwise', 'loss']

    # ---- the capturable part: head forward -> loss -> parse_losses -> backward, static sampled instances
    samp = [t.detach().clone() for t in sampled]
    par_s = samp[0].requires_grad_(True)
    feat_s = mask_feat.detach().clone().requires_grad_(True)

    def step():
        pr = head(feat_s, par_s, samp[1], samp[2], samp[3])
        ls = head.loss(img, case['metas'], pr, samp[4], boxes, None, None)
        total, lv = parse_losses(ls)
        g = torch.autograd.grad(total, [feat_s, par_s])
        return total.detach(), lv.values_device, g[0], g[1]

    head._iter.fill_(5000)
    ref = [t.clone() for t in step()]
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        head._iter.fill_(5000)
        step()
    torch.cuda.current_stream().wait_stream(side)
    head._iter.fill_(5000)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(2):
        head._iter.fill_(5000)
        graph.replay()
        torch.cuda.synchronize()
        for a, b in zip(out, ref):
            assert torch.equal(a, b)

"""TEST INFRASTRUCTURE (oracle): CPU restatement of ``CondInstBoxHead.get_targets`` / ``_get_target_single``
(mmdet/models/dense_heads/condinst_head.py:477-548, 550-633), the FCOS point assignment, image by image in dense
[points, gts] torch tensors exactly as the reference evaluates it (same fp32 operations in the same order; CPU float32
division is IEEE).  Pinned on the reference's own two methods (AST-extracted, ``oracle/make_golden_fcos.py``) through
``tests/golden/fcos_targets.npz``.  One deliberate difference: an image WITHOUT ground truth returns (background, zero
targets, index -1) -- the reference returns only two values there (:554-556) and would fail to unpack in ``multi_apply``.
"""
import torch

INF = 1e8


def target_single(gt_bboxes, gt_labels, points, regress_ranges, num_points_per_lvl, strides, num_classes, center_sampling,
                  radius):
    """:550-633 for one image.  points [P,2], regress_ranges [P,2] (expanded per point)."""
    P, G = points.size(0), gt_labels.size(0)
    if G == 0:
        return (gt_labels.new_full((P,), num_classes), gt_bboxes.new_zeros((P, 4)), gt_labels.new_full((P,), -1))
    wh = gt_bboxes[:, 2:] - gt_bboxes[:, :2]
    areas = (wh[:, 0] * wh[:, 1])[None].repeat(P, 1)                                   # :558-562
    box = gt_bboxes[None].expand(P, G, 4)
    xs = points[:, 0, None].expand(P, G)
    ys = points[:, 1, None].expand(P, G)
    ltrb = torch.stack((xs - box[..., 0], ys - box[..., 1], box[..., 2] - xs, box[..., 3] - ys), -1)     # :571-575
    if center_sampling:                                                                # :577-615
        cx = (box[..., 0] + box[..., 2]) / 2
        cy = (box[..., 1] + box[..., 3]) / 2
        reach = cx.new_zeros(cx.shape)
        start = 0
        for lvl, n in enumerate(num_points_per_lvl):
            reach[start:start + n] = strides[lvl] * radius
            start += n
        x0 = torch.maximum(cx - reach, box[..., 0])          # where(x_mins > x1, x_mins, x1)
        y0 = torch.maximum(cy - reach, box[..., 1])
        x1 = torch.minimum(cx + reach, box[..., 2])          # where(x_maxs > x2, x2, x_maxs)
        y1 = torch.minimum(cy + reach, box[..., 3])
        inside = torch.stack((xs - x0, ys - y0, x1 - xs, y1 - ys), -1).min(-1)[0] > 0
    else:
        inside = ltrb.min(-1)[0] > 0                                                   # :618
    far = ltrb.max(-1)[0]                                                              # :621
    rr = regress_ranges[:, None, :].expand(P, G, 2)
    in_range = (far >= rr[..., 0]) & (far <= rr[..., 1])
    areas[~inside] = INF
    areas[~in_range] = INF
    min_area, arg = areas.min(dim=1)                                                   # first minimal entry on CPU
    labels = gt_labels[arg].clone()
    labels[min_area == INF] = num_classes
    targets = ltrb[torch.arange(P), arg]
    arg = arg.clone()
    arg[min_area == INF] = -1
    return labels, targets, arg


def get_targets(points, gt_bboxes_list, gt_labels_list, regress_ranges, strides, num_classes, center_sampling=True, radius=1.5,
                norm_on_bbox=True):
    """:477-548.  points: list of [P_l,2]; returns the three per-level lists."""
    L = len(points)
    rr = torch.cat([points[i].new_tensor(regress_ranges[i])[None].expand_as(points[i]) for i in range(L)], 0)
    pts = torch.cat(points, 0)
    counts = [p.size(0) for p in points]
    per_img = [target_single(b, l, pts, rr, counts, strides, num_classes, center_sampling, radius)
               for b, l in zip(gt_bboxes_list, gt_labels_list)]
    cum = 0
    for (_, _, inds), b in zip(per_img, gt_bboxes_list):                               # :519-522
        inds[inds != -1] += cum
        cum += b.size(0)
    labels, targets, inds = [], [], []
    for i in range(L):
        lo, hi = sum(counts[:i]), sum(counts[:i + 1])
        labels.append(torch.cat([r[0][lo:hi] for r in per_img]))
        t = torch.cat([r[1][lo:hi] for r in per_img])
        targets.append(t / strides[i] if norm_on_bbox else t)
        inds.append(torch.cat([r[2][lo:hi] for r in per_img]))
    return labels, targets, inds


def grid_points(featmap_sizes, strides, dtype=torch.float32):
    """mmdet MlvlPointGenerator.grid_priors with offset 0.5 (what AnchorFreeHead's prior_generator yields for FCOS heads:
    ((x + 0.5) * stride, (y + 0.5) * stride), x fastest)."""
    out = []
    for (h, w), s in zip(featmap_sizes, strides):
        xs = (torch.arange(w, dtype=dtype) + 0.5) * s
        ys = (torch.arange(h, dtype=dtype) + 0.5) * s
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
        out.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    return out

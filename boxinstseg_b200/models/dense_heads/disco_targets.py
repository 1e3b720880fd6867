"""f2: the two SOLO target builders of DiscoBox -- ``DiscoBoxSOLOv2Head.solov2_target_single``
(mmdet/models/dense_heads/discobox_head.py:1442-1529) and ``best_target_single`` (:1362-1440) -- on device tensors, all
ground truths and all levels of an image at once.

The reference runs them on the host side of the step: a Python loop over levels, ground truths and grid cells with one
``int(tensor)`` device read per cell term, a numpy round trip of the masks and one ``mmcv.imrescale`` (``cv2.resize``) per
ground truth.  Here: the mass centres are two integer reductions rounded to float32 once (the reference's float32
``center_of_mass`` (:522-532) accumulates in float32: identical whenever sum(x * mask) < 2^24, i.e. below ~4 000 mask pixels
per image column sum -- above that its own CPU and CUDA runs differ by summation order, and this is the correctly rounded
value of what both approximate); the cell windows are float32 tensor arithmetic with the reference's operations (IEEE
division and Python floor division on every device, ``solo_targets._true_div/_floor_div``); "the last ground truth written
wins" is a max over ground-truth indices; the covered cells of ALL levels come out of one ``nonzero`` (the only host
synchronisation: the output sizes are data dependent) in the reference's loop order (level, ground truth, row, column); the
1/4 down-scaling is OpenCV's fixed-point bilinear restated in integer ops (``cv2_resize_linear_u8``, bit-exact).
"""
import torch

from .solo_targets import _floor_div, _true_div, cv2_resize_linear_u8, rescale_size


def scale_mids(scale_ranges, device):
    """:701-702: (lower * upper) ** 0.5 of an int64 tensor -> float32."""
    r = torch.tensor([[int(a), int(b)] for a, b in scale_ranges], dtype=torch.int64, device=device)
    return torch.sqrt((r[:, 0] * r[:, 1]).to(torch.float32))


def disco_target_single(gt_bboxes, gt_labels, gt_masks, mask_feat_size, scale_ranges, strides, seg_num_grids, sigma,
                        num_classes, best=False):
    """gt_bboxes [G,4] float32, gt_labels [G] int64, gt_masks [G,H,W] uint8, one device.  Returns the four per-level lists of
    the reference: ``ins_label`` [n_l, fh, fw] uint8 (one entry per covered cell, loop order), ``cate_label`` [grid, grid]
    int64, ``ins_ind_label`` [grid^2] bool, ``grid_order`` [n_l] int64 (a device tensor where the reference has a Python list:
    its consumers (:926-935) only index with it).  ``best=True`` is ``best_target_single``."""
    dev = gt_bboxes.device
    G = gt_bboxes.shape[0]
    fh, fw = int(mask_feat_size[0]), int(mask_feat_size[1])
    L = len(seg_num_grids)
    if G == 0:
        return ([torch.zeros((0, fh, fw), dtype=torch.uint8, device=dev) for _ in range(L)],
                [torch.full((g, g), num_classes, dtype=torch.int64, device=dev) for g in seg_num_grids],
                [torch.zeros(g * g, dtype=torch.bool, device=dev) for g in seg_num_grids],
                [torch.zeros(0, dtype=torch.int64, device=dev) for _ in range(L)])
    areas = torch.sqrt((gt_bboxes[:, 2] - gt_bboxes[:, 0]) * (gt_bboxes[:, 3] - gt_bboxes[:, 1]))
    up_h, up_w = fh * 4, fw * 4
    H, W = gt_masks.shape[-2:]
    m64 = gt_masks.to(torch.int64)
    row_sum, col_sum = m64.sum(2), m64.sum(1)                                       # [G,H], [G,W]
    total = row_sum.sum(1)
    m00 = total.clamp(min=1).to(torch.float32)
    ch = (row_sum * torch.arange(H, device=dev)).sum(1).to(torch.float32) / m00     # center_of_mass (:522-532), float32
    cw = (col_sum * torch.arange(W, device=dev)).sum(1).to(torch.float32) / m00
    half_w = 0.5 * (gt_bboxes[:, 2] - gt_bboxes[:, 0]) * sigma
    half_h = 0.5 * (gt_bboxes[:, 3] - gt_bboxes[:, 1]) * sigma
    nonempty = total > 0
    if best:                                                                        # :1375-1378
        diffs = scale_mids(scale_ranges, dev)[None] / (areas[:, None] + 1e-6)
        diffs = torch.where(diffs < 1, 1 / (diffs + 1e-6), diffs)
        level_of = diffs.argmin(1)
    order = torch.arange(G, device=dev)
    off = torch.arange(-1, 2, device=dev)
    cate_list, ind_list, cand_list, cell_list = [], [], [], []
    for lvl, ((lower, upper), grid) in enumerate(zip(scale_ranges, seg_num_grids)):
        hit = (level_of == lvl) if best else ((areas >= lower) & (areas <= upper))
        valid = hit & nonempty
        cell = 1. / grid

        def cells(x, size):
            return _floor_div(_true_div(x, float(size)), cell).to(torch.int64)

        coord_h, coord_w = cells(ch, up_h), cells(cw, up_w)
        ii = coord_h[:, None, None] + off[None, :, None]                            # [G,3,1] candidate rows
        jj = coord_w[:, None, None] + off[None, None, :]                            # [G,1,3] candidate columns
        if best:
            cand = valid[:, None, None] & (ii == coord_h[:, None, None]) & (jj == coord_w[:, None, None])
        else:
            top_box = cells(ch - half_h, up_h).clamp(min=0)
            down_box = cells(ch + half_h, up_h).clamp(max=grid - 1)
            left_box = cells(cw - half_w, up_w).clamp(min=0)
            right_box = cells(cw + half_w, up_w).clamp(max=grid - 1)
            top, down = torch.maximum(top_box, coord_h - 1), torch.minimum(down_box, coord_h + 1)
            left, right = torch.maximum(coord_w - 1, left_box), torch.minimum(right_box, coord_w + 1)
            cand = (valid[:, None, None] & (ii >= top[:, None, None]) & (ii <= down[:, None, None]) &
                    (jj >= left[:, None, None]) & (jj <= right[:, None, None]))     # [G,3,3]
        cell_id = ii * grid + jj                                                    # [G,3,3]
        # cate_label / ins_ind_label: scatter the covering ground-truth index with max ("the last one written wins")
        winner = torch.full((grid * grid,), -1, dtype=torch.int64, device=dev)
        safe = torch.where(cand, cell_id, torch.zeros_like(cell_id)).flatten()
        vals = torch.where(cand, order[:, None, None].expand_as(cand), order.new_full((), -1)).flatten()
        winner = winner.scatter_reduce(0, safe, vals, reduce='amax', include_self=True)
        has = winner >= 0
        cate_list.append(torch.where(has, gt_labels[winner.clamp(min=0)], gt_labels.new_full((), num_classes)).view(grid, grid))
        ind_list.append(has)
        cand_list.append(cand)
        cell_list.append(cell_id)
    nh, nw = rescale_size(H, W, 1. / 4)
    small = cv2_resize_linear_u8(gt_masks, nh, nw)                                  # mmcv.imrescale(seg_mask, 1 / 4), :1423,1508
    canvas = torch.zeros((G, fh, fw), dtype=torch.uint8, device=dev)
    canvas[:, :nh, :nw] = small
    cand_all = torch.stack(cand_list)                                               # [L,G,3,3]
    nz = cand_all.nonzero()                                                         # rows sorted by (level, gt, row, column)
    counts = torch.bincount(nz[:, 0], minlength=L).tolist()
    cells_all = torch.stack(cell_list)[nz[:, 0], nz[:, 1], nz[:, 2], nz[:, 3]]
    ins_list = list(canvas[nz[:, 1]].split(counts))
    order_list = list(cells_all.split(counts))
    return ins_list, cate_list, ind_list, order_list

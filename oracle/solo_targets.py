"""TEST INFRASTRUCTURE (oracle): CPU restatement of ``BoxSOLOv2Head.solo_target_single``
(mmdet/models/dense_heads/box_solov2_head.py:390-472): the per-image SOLO grid targets of every FPN level.

Host-side code in the reference as well (numpy masks, ``scipy.ndimage.center_of_mass``, ``mmcv.imrescale`` = ``cv2.resize``
bilinear, Python loops over levels / ground truths / grid cells).  Third-party pieces absent from /root/reference:
``mmcv.imrescale`` (mmcv-full 1.3.17-1.6.0: ``rescale_size`` = int(dim * scale + 0.5), then ``cv2.resize(..., INTER_LINEAR)``)
is restated with the very OpenCV call; scipy / cv2 are importable in this image.  Pinned against the reference's own method
(AST-extracted, ``oracle/make_golden_solo.py``) through ``tests/golden/solo_targets.npz``.
"""
import numpy as np
import torch
import torch.nn.functional as F


def imrescale(img, scale):
    """mmcv.imrescale(img, scale) for a 2-D uint8 array: new size int(w*scale+0.5) x int(h*scale+0.5), bilinear."""
    import cv2
    h, w = img.shape[:2]
    new_w, new_h = int(w * float(scale) + 0.5), int(h * float(scale) + 0.5)
    return cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)


def solo_target_single(gt_bboxes_raw, gt_labels_raw, gt_masks_raw, norm_img, lst_feats, featmap_sizes, scale_ranges, strides,
                       seg_num_grids, sigma, num_classes):
    """gt_bboxes_raw [G,4] float tensor, gt_labels_raw [G] long, gt_masks_raw [G,H,W] uint8 NUMPY, norm_img [3,H,W],
    lst_feats [C,h,w].  Returns the five per-level lists of :390-472."""
    from scipy import ndimage
    device = gt_labels_raw.device
    gt_areas = torch.sqrt((gt_bboxes_raw[:, 2] - gt_bboxes_raw[:, 0]) * (gt_bboxes_raw[:, 3] - gt_bboxes_raw[:, 1]))
    ins_labels, cate_labels, ins_inds, scale_imgs, scale_lsts = [], [], [], [], []
    up_h, up_w = featmap_sizes[0][0] * 4, featmap_sizes[0][1] * 4
    for (lower, upper), stride, fsize, grid in zip(scale_ranges, strides, featmap_sizes, seg_num_grids):
        scale_imgs.append(F.interpolate(norm_img.unsqueeze(0), size=fsize, mode='bilinear'))
        scale_lsts.append(F.interpolate(lst_feats.unsqueeze(0), size=fsize, mode='bilinear'))
        ins = torch.zeros([grid ** 2, fsize[0], fsize[1]], dtype=torch.uint8, device=device)
        cate = torch.zeros([grid, grid], dtype=torch.int64, device=device) + num_classes
        ind = torch.zeros([grid ** 2], dtype=torch.bool, device=device)
        hits = ((gt_areas >= lower) & (gt_areas <= upper)).nonzero().flatten()
        if len(hits):
            boxes, labels = gt_bboxes_raw[hits], gt_labels_raw[hits]
            masks = gt_masks_raw[hits.cpu().numpy(), ...]
            half_ws = 0.5 * (boxes[:, 2] - boxes[:, 0]) * sigma          # float32 tensors
            half_hs = 0.5 * (boxes[:, 3] - boxes[:, 1]) * sigma
            out_stride = stride / 2
            for m, label, half_h, half_w in zip(masks, labels, half_hs, half_ws):
                if m.sum() < 10:
                    continue
                ch, cw = ndimage.center_of_mass(m)                         # float64
                coord_w = int((cw / up_w) // (1. / grid))
                coord_h = int((ch / up_h) // (1. / grid))
                # the box terms mix a numpy float64 with a 0-d float32 tensor: torch computes them in float32
                top_box = max(0, int(((ch - half_h) / up_h) // (1. / grid)))
                down_box = min(grid - 1, int(((ch + half_h) / up_h) // (1. / grid)))
                left_box = max(0, int(((cw - half_w) / up_w) // (1. / grid)))
                right_box = min(grid - 1, int(((cw + half_w) / up_w) // (1. / grid)))
                top, down = max(top_box, coord_h - 1), min(down_box, coord_h + 1)
                left, right = max(coord_w - 1, left_box), min(right_box, coord_w + 1)
                cate[top:(down + 1), left:(right + 1)] = label
                small = torch.from_numpy(imrescale(m, 1. / out_stride)).to(device)
                for i in range(top, down + 1):
                    for j in range(left, right + 1):
                        cell = int(i * grid + j)
                        ins[cell, :small.shape[0], :small.shape[1]] = small
                        ind[cell] = True
        ins_labels.append(ins)
        cate_labels.append(cate)
        ins_inds.append(ind)
    return ins_labels, cate_labels, ins_inds, scale_imgs, scale_lsts


# ---------------------------------------------------------------------------------------------------------------------
# DiscoBox: DiscoBoxSOLOv2Head.solov2_target_single (mmdet/models/dense_heads/discobox_head.py:1442-1529) and
# best_target_single (:1362-1440).  Differences from the BoxSOLOv2 builder above: the mass centre is the float32 torch
# `center_of_mass` (:522-532), every cell term is float32, the mask is always rescaled by 1/4, every covered cell appends
# its own copy of the mask (duplicates included) and the covered cells are also returned as `grid_order`;
# best_target_single assigns each ground truth to the ONE level whose geometric-mean scale is closest and marks only the
# centre cell.  Pinned through tests/golden/disco_targets.npz (oracle/make_golden_disco.py).
# ---------------------------------------------------------------------------------------------------------------------
def center_of_mass_f32(masks):
    """:522-532.  masks [n,h,w] uint8 tensor -> (center_x, center_y) float32."""
    _, h, w = masks.shape
    ys = torch.arange(0, h, dtype=torch.float32)
    xs = torch.arange(0, w, dtype=torch.float32)
    m00 = masks.sum(-1).sum(-1).clamp(min=1e-6)
    m10 = (masks * xs).sum(-1).sum(-1)
    m01 = (masks * ys[:, None]).sum(-1).sum(-1)
    return m10 / m00, m01 / m00


def disco_target_single(gt_bboxes_raw, gt_labels_raw, gt_masks_raw, mask_feat_size, scale_ranges, strides, seg_num_grids,
                        sigma, num_classes, best=False):
    """gt_masks_raw [G,H,W] uint8 NUMPY.  Returns (ins_label_list, cate_label_list, ins_ind_label_list, grid_order_list)."""
    gt_areas = torch.sqrt((gt_bboxes_raw[:, 2] - gt_bboxes_raw[:, 0]) * (gt_bboxes_raw[:, 3] - gt_bboxes_raw[:, 1]))
    fh, fw = mask_feat_size
    up_h, up_w = fh * 4, fw * 4
    if best:                                                                            # :1375-1378
        mids = torch.tensor(np.array(scale_ranges))
        mids = ((mids[:, 0] * mids[:, 1]) ** 0.5)[None]
        diffs = mids / (gt_areas[:, None] + 1e-6)
        small_ones = diffs < 1
        diffs[small_ones] = 1 / (diffs[small_ones] + 1e-6)
        level_of = diffs.argmin(1)
    ins_list, cate_list, ind_list, order_list = [], [], [], []
    for lvl, ((lower, upper), grid) in enumerate(zip(scale_ranges, seg_num_grids)):
        if best:
            hits = (level_of == lvl).nonzero().flatten()
        else:
            hits = ((gt_areas >= lower) & (gt_areas <= upper)).nonzero().flatten()
        ins, order = [], []
        cate = torch.zeros([grid, grid], dtype=torch.int64) + num_classes
        ind = torch.zeros([grid ** 2], dtype=torch.bool)
        if len(hits):
            boxes, labels = gt_bboxes_raw[hits], gt_labels_raw[hits]
            masks = gt_masks_raw[hits.numpy(), ...]
            half_ws = 0.5 * (boxes[:, 2] - boxes[:, 0]) * sigma
            half_hs = 0.5 * (boxes[:, 3] - boxes[:, 1]) * sigma
            masks_pt = torch.from_numpy(masks)
            cws, chs = center_of_mass_f32(masks_pt)
            nonempty = masks_pt.sum(-1).sum(-1) > 0
            for m, label, half_h, half_w, ch, cw, ok in zip(masks, labels, half_hs, half_ws, chs, cws, nonempty):
                if not ok:
                    continue
                coord_w = int((cw / up_w) // (1. / grid))
                coord_h = int((ch / up_h) // (1. / grid))
                small = torch.from_numpy(imrescale(m, 1. / 4))
                canvas = torch.zeros([fh, fw], dtype=torch.uint8)
                canvas[:small.shape[0], :small.shape[1]] = small
                if best:                                                                # :1422-1431
                    cate[coord_h, coord_w] = label
                    cells = [coord_h * grid + coord_w]
                else:                                                                   # :1497-1521
                    top_box = max(0, int(((ch - half_h) / up_h) // (1. / grid)))
                    down_box = min(grid - 1, int(((ch + half_h) / up_h) // (1. / grid)))
                    left_box = max(0, int(((cw - half_w) / up_w) // (1. / grid)))
                    right_box = min(grid - 1, int(((cw + half_w) / up_w) // (1. / grid)))
                    top, down = max(top_box, coord_h - 1), min(down_box, coord_h + 1)
                    left, right = max(coord_w - 1, left_box), min(right_box, coord_w + 1)
                    cate[top:(down + 1), left:(right + 1)] = label
                    cells = [i * grid + j for i in range(top, down + 1) for j in range(left, right + 1)]
                for cell in cells:
                    ins.append(canvas.clone())
                    ind[cell] = True
                    order.append(int(cell))
        ins_list.append(torch.stack(ins, 0) if ins else torch.zeros([0, fh, fw], dtype=torch.uint8))
        cate_list.append(cate)
        ind_list.append(ind)
        order_list.append(order)
    return ins_list, cate_list, ind_list, order_list

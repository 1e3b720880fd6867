"""f2: the SOLO grid targets of ``BoxSOLOv2Head.solo_target_single`` (mmdet/models/dense_heads/box_solov2_head.py:390-472) on
the device, for all ground truths of an image at once.

The reference builds them on the host: numpy masks, ``scipy.ndimage.center_of_mass``, ``mmcv.imrescale`` (= ``cv2.resize``,
bilinear, uint8) per ground truth, Python loops over levels, ground truths and grid cells, one ``int(tensor)`` device read per
box term.  Here the masks stay on the device: the mass centres are two integer reductions, the cell windows are tensor
arithmetic in the SAME floating-point types the reference's mixed numpy / torch expressions end up in (float64 for the centre
cell, float32 for the box terms), "the last ground truth written wins" becomes a max over the covering ground-truth indices,
and the uint8 bilinear down-scaling is OpenCV's fixed-point algorithm restated with integer tensor ops
(``cv2_resize_linear_u8``: bit-exact against cv2 in the tests).  No host synchronisation unless the caller asks for the
compacted positives.
"""
import torch

from ...ops.resize import bilinear_resize

_COEF = 2048                      # OpenCV INTER_RESIZE_COEF_SCALE (11 bits)


def _taps(src, dst, device):
    """Source index pair and 11-bit integer weights of every destination coordinate, as cv::resize computes them for
    INTER_LINEAR on 8-bit images (float32 coordinate, round-half-even weights)."""
    scale = 1.0 / (float(dst) / float(src))
    d = torch.arange(dst, dtype=torch.float64, device=device)
    f = ((d + 0.5) * scale - 0.5).to(torch.float32)
    s = torch.floor(f)
    f = f - s
    s = s.to(torch.int64)
    low, high = s < 0, s >= src - 1
    f = torch.where(low | high, torch.zeros_like(f), f)
    s = torch.where(low, torch.zeros_like(s), torch.where(high, torch.full_like(s, src - 1), s))
    a0 = torch.round((1.0 - f) * _COEF).to(torch.int64)
    a1 = torch.round(f * _COEF).to(torch.int64)
    return s, torch.clamp(s + 1, max=src - 1), a0, a1


def cv2_resize_linear_u8(img, new_h, new_w):
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 [..., H, W] tensors (single channel), bit
    exact for DOWN-scaling (what mmcv.imrescale(mask, 1 / output_stride) does here; checked against OpenCV 4.13 on rectangles,
    random 0/1 masks and full-range images at integer and non-integer ratios)."""
    H, W = img.shape[-2:]
    x0, x1, ax0, ax1 = _taps(W, new_w, img.device)
    y0, y1, by0, by1 = _taps(H, new_h, img.device)
    src = img.to(torch.int64)
    rows = src.index_select(-1, x0) * ax0 + src.index_select(-1, x1) * ax1            # horizontal pass, [..., H, new_w]
    # vertical pass exactly as OpenCV's 8-bit specialisation does it (VResizeLinear<uchar, int, short, ...>: operands are
    # pre-shifted, each product is truncated, then rounded once) -- NOT the textbook (b0 S0 + b1 S1 + 2^21) >> 22, which
    # differs from cv2 in ~5 % of the pixels of a random 0/1 mask at non-integer ratios
    s0, s1 = rows.index_select(-2, y0), rows.index_select(-2, y1)
    out = (((by0[:, None] * (s0 >> 4)) >> 16) + ((by1[:, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return out.clamp_(0, 255).to(torch.uint8)


def _true_div(a, b):
    """a / b with IEEE division on every device.  (ATen's CUDA kernels divide by a host scalar by multiplying with its
    reciprocal, which can differ from the quotient by one ulp -- enough to move a grid-cell boundary.  The reference's own
    float32 box terms go through that path when its boxes are CUDA tensors, so its CUDA and CPU runs can disagree in rare
    boundary cases; this implementation reproduces its CPU result -- exact division, as numpy does for the float64 terms -- on
    both devices.)"""
    return a / torch.full((), b, dtype=a.dtype, device=a.device)


def _floor_div(a, b):
    """Python's float ``a // b`` (CPython float_floor_div, which numpy and ATen's CPU kernel follow), from IEEE primitives."""
    bt = torch.full((), b, dtype=a.dtype, device=a.device)
    mod = torch.fmod(a, bt)
    div = (a - mod) / bt
    div = torch.where((mod != 0) & ((mod < 0) != (bt < 0)), div - 1, div)
    fl = torch.floor(div)
    return torch.where(div - fl > 0.5, fl + 1, fl)


def rescale_size(h, w, scale):
    """mmcv.rescale_size: int(dim * scale + 0.5)."""
    return int(h * float(scale) + 0.5), int(w * float(scale) + 0.5)


def solo_target_single(gt_bboxes, gt_labels, gt_masks, norm_img, lst_feats, featmap_sizes, scale_ranges, strides, seg_num_grids,
                       sigma, num_classes, dense=True):
    """The reference's signature and five per-level lists (:390-472): the grid targets of ``solo_grid_targets`` plus the image
    and the level-set features resized to every level (``F.interpolate(..., mode='bilinear')`` there, the a18 kernel here)."""
    ins, cate, ind = solo_grid_targets(gt_bboxes, gt_labels, gt_masks, featmap_sizes, scale_ranges, strides, seg_num_grids, sigma,
                                       num_classes, dense)
    imgs = [bilinear_resize(norm_img.unsqueeze(0), tuple(f)) for f in featmap_sizes]
    lsts = [bilinear_resize(lst_feats.unsqueeze(0), tuple(f)) for f in featmap_sizes]
    return ins, cate, ind, imgs, lsts


def solo_grid_targets(gt_bboxes, gt_labels, gt_masks, featmap_sizes, scale_ranges, strides, seg_num_grids, sigma, num_classes,
                      dense=True):
    """gt_bboxes [G,4] float32, gt_labels [G] int64, gt_masks [G,H,W] uint8, all on one device (plain torch: any device).
    Returns three per-level lists: ``ins_label`` [grid^2, fh, fw] uint8, ``cate_label`` [grid, grid] int64, ``ins_ind_label``
    [grid^2] bool; with ``dense=False`` the first list holds ``(winner [grid^2] int64 (-1: none), small [G, h', w'] uint8)``
    instead of the 80 MB dense canvas: the label of cell c is ``small[winner[c]]`` placed at the top-left corner."""
    dev = gt_bboxes.device
    G = gt_bboxes.shape[0]
    if G == 0:                                                          # an image without ground truth: empty targets
        ins = [torch.zeros((g * g, f[0], f[1]), dtype=torch.uint8, device=dev) if dense else
               (torch.full((g * g,), -1, dtype=torch.int64, device=dev), gt_masks.new_zeros((0, 0, 0)))
               for g, f in zip(seg_num_grids, featmap_sizes)]
        cate = [torch.full((g, g), num_classes, dtype=torch.int64, device=dev) for g in seg_num_grids]
        return ins, cate, [torch.zeros(g * g, dtype=torch.bool, device=dev) for g in seg_num_grids]
    areas = torch.sqrt((gt_bboxes[:, 2] - gt_bboxes[:, 0]) * (gt_bboxes[:, 3] - gt_bboxes[:, 1]))
    up_h, up_w = featmap_sizes[0][0] * 4, featmap_sizes[0][1] * 4
    H, W = gt_masks.shape[-2:]
    m64 = gt_masks.to(torch.int64)
    row_sum, col_sum = m64.sum(2), m64.sum(1)                            # [G,H], [G,W]
    total = row_sum.sum(1)
    tot_f = total.clamp(min=1).to(torch.float64)
    ch = (row_sum * torch.arange(H, device=dev)).sum(1).to(torch.float64) / tot_f       # ndimage.center_of_mass, float64
    cw = (col_sum * torch.arange(W, device=dev)).sum(1).to(torch.float64) / tot_f
    ch32, cw32 = ch.to(torch.float32), cw.to(torch.float32)             # numpy float64 (op) float32 tensor -> float32
    half_w = 0.5 * (gt_bboxes[:, 2] - gt_bboxes[:, 0]) * sigma
    half_h = 0.5 * (gt_bboxes[:, 3] - gt_bboxes[:, 1]) * sigma
    order = torch.arange(G, device=dev)
    small_by_stride = {}
    ins_list, cate_list, ind_list = [], [], []
    for (lower, upper), stride, fsize, grid in zip(scale_ranges, strides, featmap_sizes, seg_num_grids):
        valid = (areas >= lower) & (areas <= upper) & (total >= 10)
        cell = 1. / grid

        def cells(x, size):
            return _floor_div(_true_div(x, float(size)), cell).to(torch.int64)

        coord_h, coord_w = cells(ch, up_h), cells(cw, up_w)                              # float64 path
        top_box = cells(ch32 - half_h, up_h).clamp(min=0)                                # float32 path
        down_box = cells(ch32 + half_h, up_h).clamp(max=grid - 1)
        left_box = cells(cw32 - half_w, up_w).clamp(min=0)
        right_box = cells(cw32 + half_w, up_w).clamp(max=grid - 1)
        top, down = torch.maximum(top_box, coord_h - 1), torch.minimum(down_box, coord_h + 1)
        left, right = torch.maximum(coord_w - 1, left_box), torch.minimum(right_box, coord_w + 1)
        ii = torch.arange(grid, device=dev)
        cover = (valid[:, None, None] & (ii[None, :, None] >= top[:, None, None]) & (ii[None, :, None] <= down[:, None, None]) &
                 (ii[None, None, :] >= left[:, None, None]) & (ii[None, None, :] <= right[:, None, None]))       # [G,grid,grid]
        winner = torch.where(cover, order[:, None, None], order.new_full((), -1)).max(0)[0]   # the last ground truth written wins
        has = winner >= 0
        cate_list.append(torch.where(has, gt_labels[winner.clamp(min=0)], gt_labels.new_full((), num_classes)))
        ind_list.append(has.flatten())
        out_stride = stride / 2
        if out_stride not in small_by_stride:
            nh, nw = rescale_size(H, W, 1. / out_stride)
            small_by_stride[out_stride] = cv2_resize_linear_u8(gt_masks, nh, nw)
        small = small_by_stride[out_stride]
        if dense:
            ins = torch.zeros((grid * grid, fsize[0], fsize[1]), dtype=torch.uint8, device=dev)
            picked = small[winner.flatten().clamp(min=0)] * has.flatten()[:, None, None].to(torch.uint8)
            ins[:, :small.shape[1], :small.shape[2]] = picked
            ins_list.append(ins)
        else:
            ins_list.append((winner.flatten(), small))
    return ins_list, cate_list, ind_list

"""f2: the target side of ``CondInstBoxHead`` (mmdet/models/dense_heads/condinst_head.py:247-548): ``get_targets`` +
``_get_target_single`` -- the FCOS point assignment that runs immediately before ``training_sample`` and the mask loss.

The reference loops over the images (``multi_apply``), builds ~25 ``[points, gts]`` tensors per image, then splits per
level, concatenates per image and divides by the stride.  Here the whole batch is ONE launch of ``bxs_fcos_targets``
(`csrc/assign_targets.cu`): one thread per (image, location), results written in the reference's output layout, bit-exact.
No host synchronisation and no host-to-device copy: the level table and the per-image ground-truth offsets (host-known
lengths) travel in the kernel parameters, so the call can be captured in a CUDA graph.

Registered as a *partial* head: with mmdet importable the registered class inherits everything else (layers, forward,
loss, get_bboxes) from the reference's own ``CondInstBoxHead``.
"""
import ctypes

import torch

from ... import _lib as L
from ..builder import HEADS, register

INF = 1e8      # condinst_head.py:16


def fcos_get_targets(points, gt_bboxes_list, gt_labels_list, regress_ranges, strides, num_classes, center_sampling=True,
                     center_sample_radius=1.5, norm_on_bbox=True):
    """points: list (one per level) of [P_l, 2] float32 CUDA tensors; gt_bboxes_list / gt_labels_list: per image [G_b, 4]
    float32 / [G_b] int64.  Returns the three per-level lists of condinst_head.py:546-548: labels [B*P_l] int64,
    bbox_targets [B*P_l, 4] float32, gt_inds [B*P_l] int64 (offset by the ground truths of the preceding images)."""
    num_levels = len(points)
    assert num_levels == len(regress_ranges) == len(strides)
    assert len(gt_bboxes_list) == len(gt_labels_list) and len(gt_bboxes_list) > 0
    B = len(gt_bboxes_list)
    pts = torch.cat([p.reshape(-1, 2) for p in points], 0).contiguous().float()
    L.require_cuda(pts)
    dev = pts.device
    counts = [int(p.shape[0]) for p in points]
    level_off = [0]
    for c in counts:
        level_off.append(level_off[-1] + c)
    P = level_off[-1]
    gts = [int(g.shape[0]) for g in gt_bboxes_list]
    off = [0]
    for g in gts:
        off.append(off[-1] + g)
    boxes = torch.cat([g.reshape(-1, 4) for g in gt_bboxes_list], 0).to(device=dev, dtype=torch.float32).contiguous()
    labs = torch.cat([g.reshape(-1) for g in gt_labels_list], 0).to(device=dev, dtype=torch.int64).contiguous()
    labels = torch.empty(B * P, dtype=torch.int64, device=dev)
    targets = torch.empty((B * P, 4), dtype=torch.float32, device=dev)
    inds = torch.empty(B * P, dtype=torch.int64, device=dev)
    if P:
        f32 = lambda v: torch.tensor(v, dtype=torch.float64).to(torch.float32)        # double -> fp32, like a tensor assignment
        lo = f32([float(r[0]) for r in regress_ranges])
        hi = f32([float(r[1]) for r in regress_ranges])
        sr = f32([float(s) * center_sample_radius for s in strides])                   # :595 (Python float product, then fp32)
        st = f32([float(s) for s in strides])
        lvl = (ctypes.c_int64 * (num_levels + 1))(*level_off)
        gt_off = (ctypes.c_int64 * (B + 1))(*off)
        with torch.cuda.device(dev):
            L.check(L.lib().bxs_fcos_targets(
                L.ptr(pts), L.ptr(boxes) if boxes.numel() else None, L.ptr(labs) if labs.numel() else None,
                ctypes.cast(gt_off, ctypes.c_void_p),
                L.ptr(labels), L.ptr(targets), L.ptr(inds), B, num_levels, ctypes.cast(lvl, ctypes.c_void_p),
                L.c_p(lo.data_ptr()), L.c_p(hi.data_ptr()), L.c_p(sr.data_ptr()), L.c_p(st.data_ptr()),
                int(bool(center_sampling)), int(bool(norm_on_bbox)), int(num_classes), L.stream()), 'fcos_targets')
    sizes = [B * c for c in counts]
    return list(labels.split(sizes)), list(targets.split(sizes)), list(inds.split(sizes))


@register(HEADS, partial=True)
class CondInstBoxHead:
    """The target builder of the reference's ``CondInstBoxHead``; everything else is inherited from the reference's class
    when mmdet is importable (``register(partial=True)``).  Stand-alone it carries the attributes ``get_targets`` reads."""

    def __init__(self, num_classes, in_channels=256, regress_ranges=((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF)),
                 strides=(8, 16, 32, 64, 128), center_sampling=True, center_sample_radius=1.5, norm_on_bbox=True, **kwargs):
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.regress_ranges = regress_ranges
        self.strides = strides
        self.center_sampling = center_sampling
        self.center_sample_radius = center_sample_radius
        self.norm_on_bbox = norm_on_bbox

    def get_targets(self, points, gt_bboxes_list, gt_labels_list):
        """condinst_head.py:477-548, same arguments and return value."""
        assert len(points) == len(self.regress_ranges)
        return fcos_get_targets(points, gt_bboxes_list, gt_labels_list, self.regress_ranges, self.strides, self.num_classes,
                                self.center_sampling, self.center_sample_radius, self.norm_on_bbox)

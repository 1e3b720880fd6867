"""Mask-loss assembly of the three SOLO/Mask2Former-style heads and their
dynamic-conv call sites (a2-a4), on the kernels of libboxseg_b200:

  ``BoxSOLOv2Head``       mmdet/models/dense_heads/box_solov2_head.py:204-217 (dynamic conv), :334-367 (loss)
  ``DiscoBoxSOLOv2Head``  mmdet/models/dense_heads/discobox_head.py:1206-1300 (per-image conv, MIL dice, mean-field teacher)
  ``Box2MaskHead``        mmdet/models/dense_heads/box2mask_head.py:338-359 (mask_pred einsum), :269-335 (loss_single)

Only the mask-loss hot path is implemented here.  The conv towers, SOLO / Hungarian target building,
category losses, inference and DiscoBox's cross-image correspondence loss are out of scope:
the classes keep the reference's registry names and accept (and store) the reference's constructor kwargs so
the configs build, and expose the hot-path pieces as methods with the tensors the reference passes between
its own lines.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ...ops.dynconv import box2mask_mask_pred, dynconv1x1, solo_dynamic_conv
from ...ops.tree_filter import MinimumSpanningTree, TreeFilter2D
from ..builder import HEADS, build_loss, register
from ..losses import LCM, mil_loss
from .meanfield import MeanField


def _scale_target(t, size=(96, 96)):
    """mmdet/models/utils/misc.py:75-86."""
    if t.dim() == 3:
        t = t.unsqueeze(1)
    return F.interpolate(t, size=size, mode='bilinear', align_corners=False)


def _phi_and_pixels(mask_pred, box_mask):
    """cat(s, 1-s) * box ; clamp(sum box, 1)  (box_solov2_head.py:341-349, box2mask_head.py:305-310)."""
    phi = torch.cat((mask_pred, 1.0 - mask_pred), dim=1) * box_mask
    pix = box_mask.sum((1, 2, 3)).clamp(min=1)
    return phi, pix


@register(HEADS, partial=True)
class BoxSOLOv2Head(nn.Module):
    """BoxLevelset head: mask-loss path only."""

    def __init__(self, num_classes=80, in_channels=256, loss_boxpro=None, loss_levelset=None, **cfg):
        super().__init__()
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.cfg = cfg
        self.loss_boxpro = build_loss(loss_boxpro or dict(type='BoxProjectionLoss', loss_weight=3.0))
        self.loss_levelset = build_loss(loss_levelset or dict(type='LevelsetLoss', loss_weight=1.0))
        self.mst = MinimumSpanningTree(TreeFilter2D.norm2_distance)
        self.tree_filter = TreeFilter2D()

    # a2 -- box_solov2_head.py:204-217
    def dynamic_conv(self, feature_pred, kernel_pred, out_size=None, cells=None):
        """feature_pred [B,C,h,w], kernel_pred [B,C,S,S] -> [B,S*S,h',w'].  ``cells`` (list of per-image index
        tensors) restricts the convolution to the positive grid cells the loss will index (:317-320), which is
        what the training loss needs; without it all S*S cells are computed as the reference does."""
        if cells is None:
            ins = solo_dynamic_conv(feature_pred, kernel_pred)
        else:
            B, C, S, _ = kernel_pred.shape
            flat = kernel_pred.permute(0, 2, 3, 1).reshape(B, S * S, C)
            ins = [dynconv1x1(feature_pred[b:b + 1], flat[b:b + 1, cells[b]])[0] for b in range(B)]
        if out_size is not None:
            up = lambda t: F.interpolate(t, size=out_size, mode='bilinear')   # noqa: E731  (:213)
            ins = up(ins) if cells is None else [up(t[None])[0] for t in ins]
        return ins

    # a17 -- box_solov2_head.py:334-367
    def mask_loss(self, ins_preds, ins_labels, img_targets, lst_targets, shared_trees=True):
        """Per level lists: ins_pred [n,h,w] logits, ins_label [n,h,w] box masks, img_target [n,3,h,w],
        lst_target [n,5,h,w].  ``shared_trees``: instances of one image share their MSTs -- the reference
        rebuilds identical trees per instance (:300-305,353); here identical rows are detected and deduplicated,
        which does not change the result."""
        loss_project, loss_levelset = [], []
        for ins_pred, box_mask, img_t, lst_t in zip(ins_preds, ins_labels, img_targets, lst_targets):
            if ins_pred.size(0) == 0:
                continue
            mask_pred = torch.sigmoid(ins_pred.unsqueeze(1))
            box = box_mask.unsqueeze(1).to(mask_pred.dtype)
            loss_project.append(self.loss_boxpro(mask_pred, box))
            phi, pix = _phi_and_pixels(mask_pred, box)
            loss_img = self.loss_levelset(phi, img_t * box, pix) * 0.05
            f_img = self.tree_filter(mask_pred, img_t, self._mst(img_t, shared_trees))
            f_lst = self.tree_filter(f_img, lst_t, self._mst(lst_t, shared_trees), low_tree=False)
            high = torch.cat((f_img, f_lst), dim=1) * box
            loss_feat = self.loss_levelset(phi, high, pix) * 5.0
            loss_levelset.append(loss_img + loss_feat)
        return dict(loss_boxpro=torch.cat(loss_project).mean(), loss_levelset=torch.cat(loss_levelset).mean())

    def _mst(self, guide, shared):
        if not shared or guide.size(0) == 1:
            return self.mst(guide)
        with torch.no_grad():
            flat = guide.flatten(1)
            new = torch.ones(guide.size(0), dtype=torch.bool, device=guide.device)
            new[1:] = (flat[1:] != flat[:-1]).any(1)           # instances arrive grouped by image (:296-305)
            group = torch.cumsum(new.long(), 0) - 1
            uniq = guide[new]
        return self.mst(uniq)[group]


@register(HEADS, partial=True)
class DiscoBoxSOLOv2Head(nn.Module):
    """DiscoBox head: mask-loss path only (corr_loss / object bank out of scope)."""

    def __init__(self, num_classes=80, in_channels=256, loss_ins=None, loss_ts=None, **cfg):
        super().__init__()
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.cfg = cfg
        loss_ins = loss_ins or dict(loss_weight=1.0)
        loss_ts = loss_ts or dict(loss_weight=1.0, alpha0=2.0, theta0=0.5, theta1=30.0, theta2=20.0, kernel=3, base=0.10,
                                  max_iter=10)
        self.ins_loss_weight = loss_ins['loss_weight']
        self.ts_loss_weight = loss_ts['loss_weight']
        self.alpha0, self.theta0, self.theta1 = loss_ts['alpha0'], loss_ts['theta0'], loss_ts['theta1']
        self.theta2 = loss_ts.get('theta2', 10)
        self.mkernel, self.crf_base, self.crf_max_iter = loss_ts['kernel'], loss_ts['base'], loss_ts['max_iter']

    # a3 -- discobox_head.py:1206-1246
    @staticmethod
    def dynamic_conv(mask_feat_img, kernels):
        """mask_feat_img [C,h,w] of ONE image, kernels [C,I] (as gathered by grid_order, :1180-1185) -> [I,h,w]."""
        return dynconv1x1(mask_feat_img[None], kernels.t()[None].contiguous())[0]

    # a17 -- discobox_head.py:1266-1300, 1302-1339 (without corr_loss)
    def mask_loss(self, s_ins_pred_list, ins_labels, img_ind_list, color_feats, t_ins_pred_list=None, use_loss_ts=True):
        """Per level: s_ins_pred [n,h,w] logits, ins_label [n,h,w] box masks, img_inds [n];
        color_feats [B,3,h,w] = image resized with align_corners=True (:1201)."""
        mean_fields = [MeanField(cf.unsqueeze(0), alpha0=self.alpha0, theta0=self.theta0, theta1=self.theta1,
                                 theta2=self.theta2, iter=self.crf_max_iter, kernel_size=self.mkernel,
                                 base=self.crf_base) for cf in color_feats] if use_loss_ts else []
        t_list = t_ins_pred_list if t_ins_pred_list is not None else s_ins_pred_list
        loss_ins, loss_ts = [], []
        for s_in, t_in, img_inds, target in zip(s_ins_pred_list, t_list, img_ind_list, ins_labels):
            if s_in is None:
                continue
            keep = target.flatten(1).sum(1) > 0                      # remove all-zero targets (:1283-1287)
            if not bool(keep.any()):
                continue
            s = torch.sigmoid(s_in)[keep]
            t = s if t_ins_pred_list is None else torch.sigmoid(t_in)[keep]
            img_inds, target = img_inds[keep], target[keep].float()
            loss_ins.append(mil_loss(None, s, s, target))
            if use_loss_ts:
                enlarged = F.max_pool2d(target.unsqueeze(1), kernel_size=3, stride=1, padding=1).squeeze(1)
                for img_idx, mf in enumerate(mean_fields):
                    sel = img_inds == img_idx
                    if not bool(sel.any()):
                        continue
                    pseudo, _ = mf(((t[sel] + s[sel]) / 2).unsqueeze(1), target[sel].unsqueeze(1))
                    loss_ts.append(_disco_dice(s[sel] * enlarged[sel], pseudo))
        zero = color_feats.new_zeros(())
        l_ins = torch.cat(loss_ins).mean() * self.ins_loss_weight if loss_ins else zero
        l_ts = torch.cat(loss_ts).mean() * self.ts_loss_weight if (use_loss_ts and loss_ts) else zero
        return dict(loss_ins=l_ins, loss_ts=l_ts)


def _disco_dice(x, t):
    """dice_loss of discobox_head.py:542-550 on full maps (plain reductions; the profiles are the MIL part)."""
    x, t = x.flatten(1).float(), t.flatten(1).float()
    return 1 - 2 * (x * t).sum(1) / ((x * x).sum(1) + 0.001 + (t * t).sum(1) + 0.001)


@register(HEADS, partial=True)
class Box2MaskHead(nn.Module):
    """Box2Mask head: mask-loss path only."""

    def __init__(self, in_channels=None, feat_channels=256, out_channels=256, num_things_classes=80, num_stuff_classes=0,
                 num_queries=100, loss_box=None, loss_mask=None, **cfg):
        super().__init__()
        self.num_queries = num_queries
        self.cfg = cfg
        self.loss_box = build_loss(loss_box or dict(type='BoxProjectionLoss', loss_weight=5.0))
        self.loss_mask = build_loss(loss_mask or dict(type='LevelsetLoss', loss_weight=1.0))
        self.mst = MinimumSpanningTree(TreeFilter2D.norm2_distance)
        self.tree_filter = TreeFilter2D()

    # a4 -- box2mask_head.py:343-345
    @staticmethod
    def mask_pred(mask_embed, mask_feature):
        return box2mask_mask_pred(mask_embed, mask_feature)

    # a17 -- box2mask_head.py:229-233, 269-335
    def mask_loss_single(self, mask_preds, mask_targets, num_per_img, norm_img, lst_feat):
        """mask_preds [n,h,w] matched logits (images concatenated), mask_targets [n,H,W] box masks,
        num_per_img: list of matched queries per image; norm_img [B,3,*,*], lst_feat [B,1,*,*]."""
        pred_shape = mask_preds.shape[-2:]
        norm_img = F.interpolate(norm_img, pred_shape, mode='bilinear', align_corners=False)
        lst_feat = F.interpolate(lst_feat, pred_shape, mode='bilinear', align_corners=False)
        if mask_preds.shape[0] == 0:                                  # zero match (:264-268)
            return mask_preds.sum(), mask_preds.sum()
        img_tree = self.mst(_scale_target(norm_img))                  # one tree per IMAGE (:269-272)
        lst_tree = self.mst(_scale_target(lst_feat))
        rep = torch.as_tensor(num_per_img, device=mask_preds.device)
        img_targets = norm_img.repeat_interleave(rep, 0)
        lst_targets = lst_feat.repeat_interleave(rep, 0)
        box = F.interpolate(mask_targets.unsqueeze(1).to(mask_preds.dtype), pred_shape, mode='bilinear', align_corners=False)
        s = torch.sigmoid(mask_preds.unsqueeze(1))
        loss_project = self.loss_box(s, box).mean()
        phi, pix = _phi_and_pixels(s, box)
        loss_img = self.loss_mask(phi, img_targets * box, pix).mean() * 0.05
        img96, lst96, s96 = _scale_target(img_targets), _scale_target(lst_targets), _scale_target(s)
        f_img = self.tree_filter(s96, img96, img_tree.repeat_interleave(rep, 0))
        f_lst = self.tree_filter(f_img, lst96, lst_tree.repeat_interleave(rep, 0), low_tree=False)
        up = lambda t: F.interpolate(t, pred_shape, mode='bilinear', align_corners=False)   # noqa: E731
        deep = torch.cat((up(f_img), up(f_lst)), dim=1) * box
        loss_feat = self.loss_mask(phi, deep, pix).mean() * 5.0
        loss_lcm = 0.2 * LCM(img96, s96, _scale_target(box))
        return loss_project, loss_img + loss_feat + loss_lcm

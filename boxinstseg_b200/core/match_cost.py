"""``BoxMatchingCost`` -- drop-in for mmdet/core/bbox/match_costs/match_cost.py:364-425.

Same constructor (``weight, pred_act, eps``) and ``__call__(mask_preds [Q,1,H,W], gt_box_masks [G,1,H,W]) -> [Q,G]``.
Additionally ``cost_from_lowres(mask_pred [Q,h,w], gt_box_masks [G,H,W])`` fuses the bilinear upsampling that
Box2MaskHead._get_target_single performs first (box2mask_head.py:157-161): the [Q,1,H,W] tensor is never built.
"""
import torch

from .. import _lib as L
from ..models.builder import MATCH_COST, register


@torch.no_grad()
def projection_profiles(x, out_size=None, sigmoid=False):
    """x [n,h,w] -> (row profile [n,H], column profile [n,W]) of x bilinearly resized to out_size."""
    x = x.contiguous().float()
    L.require_cuda(x)
    n, h, w = x.shape
    H, W = (h, w) if out_size is None else (int(out_size[0]), int(out_size[1]))
    row = torch.empty((n, H), dtype=torch.float32, device=x.device)
    col = torch.empty((n, W), dtype=torch.float32, device=x.device)
    if n:
        ws = torch.empty(n * (H + W), dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.lib().bxs_upsampled_rowcol_max(L.ptr(x), L.ptr(row), L.ptr(col), L.ptr(ws), n, h, w, H, W,
                                                     int(sigmoid), L.stream()), 'upsampled_rowcol_max')
    return row, col


@register(MATCH_COST)
class BoxMatchingCost:
    def __init__(self, weight=1., pred_act=False, eps=1e-3):
        self.weight = weight
        self.pred_act = pred_act
        self.eps = eps

    def bin_dice_loss(self, pred_prof, gt_prof):
        """[Q,L] x [G,L] -> [Q,G]:  1 - (2<p,g> + eps)/(|p|^2 + |g|^2 + eps)   (match_cost.py:386-398)."""
        num = 2 * pred_prof @ gt_prof.t()
        den = pred_prof.pow(2).sum(1)[:, None] + gt_prof.pow(2).sum(1)[None, :]
        return 1 - (num + self.eps) / (den + self.eps)

    def _cost(self, pred_row, pred_col, gt_row, gt_col):
        return (self.bin_dice_loss(pred_row, gt_row) + self.bin_dice_loss(pred_col, gt_col)) * self.weight

    @torch.no_grad()
    def __call__(self, mask_preds, gt_box_masks):
        q = mask_preds.reshape(mask_preds.shape[0], *mask_preds.shape[-2:])
        g = gt_box_masks.reshape(gt_box_masks.shape[0], *gt_box_masks.shape[-2:]).float()
        return self._cost(*projection_profiles(q, None, self.pred_act), *projection_profiles(g))

    @torch.no_grad()
    def cost_from_lowres(self, mask_pred, gt_box_masks):
        g = gt_box_masks.reshape(gt_box_masks.shape[0], *gt_box_masks.shape[-2:]).float()
        return self._cost(*projection_profiles(mask_pred, g.shape[-2:], self.pred_act), *projection_profiles(g))

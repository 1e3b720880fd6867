"""``BoxProjectionLoss`` -- registry drop-in for mmdet/models/losses/box_projection_loss.py:5-42
(same name, ``loss_weight`` kwarg, ``forward(mask_scores, box_bitmask) -> [n]``), computed by the
projection kernels of libboxseg_b200 (row/column maxima + 1-D dice, forward and backward)."""
import torch
import torch.nn as nn

from ... import _lib as L
from ..builder import LOSSES, register


class _ProjectionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, targets, eps, loss_weight):
        n, h, w = scores.shape[0], scores.shape[-2], scores.shape[-1]
        s = scores.reshape(n, h, w).contiguous().float()
        t = targets.reshape(n, h, w).contiguous().float()
        L.require_cuda(s, t)
        loss = torch.empty(n, dtype=torch.float32, device=s.device)
        lib = L.lib()
        ws = torch.empty(max(lib.bxs_projection_workspace_bytes(n, h, w), 1), dtype=torch.uint8, device=s.device)
        if n:
            with torch.cuda.device(s.device):
                L.check(lib.bxs_projection_loss_forward(L.ptr(s), L.ptr(t), L.ptr(loss), L.ptr(ws), n, h, w,
                                                        float(eps), float(loss_weight), L.stream()),
                        'projection_loss_forward')
        ctx.save_for_backward(ws)
        ctx.shape = scores.shape
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        (ws,) = ctx.saved_tensors
        shape = ctx.shape
        n, h, w = shape[0], shape[-2], shape[-1]
        g = torch.empty((n, h, w), dtype=torch.float32, device=ws.device)
        if n:
            with torch.cuda.device(ws.device):
                L.check(L.lib().bxs_projection_loss_backward(L.ptr(ws), L.ptr(g_loss.contiguous().float()), L.ptr(g),
                                                             n, h, w, L.stream()), 'projection_loss_backward')
        return g.reshape(shape), None, None, None


def projection_losses(mask_scores, box_bitmask, eps=1e-5, loss_weight=1.0):
    """[n] = loss_weight * (dice over the row profiles + dice over the column profiles).
    mask_scores / box_bitmask: [n,1,h,w] or [n,h,w]."""
    return _ProjectionLoss.apply(mask_scores, box_bitmask, eps, loss_weight)


@register(LOSSES)
class BoxProjectionLoss(nn.Module):
    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, mask_scores, box_bitmask):
        return projection_losses(mask_scores, box_bitmask, 1e-5, self.loss_weight)

    # same helper names as the reference class, for callers that use them directly
    def compute_project_term(self, mask_scores, gt_bitmasks):
        return projection_losses(mask_scores, gt_bitmasks, 1e-5, 1.0)


def mil_loss(loss_func, input, _, target):
    """DiscoBox multiple-instance projection loss (discobox_head.py:552-562) with its dice_loss
    (:542-550): eps 1e-3 on each squared norm.  ``loss_func`` is accepted for signature parity."""
    return projection_losses(input, target, 2e-3, 1.0)

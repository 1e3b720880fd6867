"""``pairwise_nlog(logits, pairwise_size, pairwise_dilation)`` -- drop-in for
mmdet/ops/pairwise/pairwise.py:6-26 (same name, signature, differentiability)."""
from torch.autograd import Function

from .pairwise_ext import pairwise_nlog_backward, pairwise_nlog_forward


class _pairwise_nlog(Function):
    @staticmethod
    def forward(ctx, logits, pairwise_size, pairwise_dilation):
        logits = logits.contiguous()
        ctx.args = (pairwise_size, pairwise_dilation)
        ctx.save_for_backward(logits)
        return pairwise_nlog_forward(pairwise_size, pairwise_dilation, logits)

    @staticmethod
    def backward(ctx, g_pairwise):
        (logits,) = ctx.saved_tensors
        size, dilation = ctx.args
        return pairwise_nlog_backward(size, dilation, logits, None, g_pairwise.contiguous()), None, None


pairwise_nlog = _pairwise_nlog.apply

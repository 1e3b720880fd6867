"""Stand-in for the reference's pybind module ``pairwise_ext``
(mmdet/ops/pairwise/csrc/pairwise/bind.cpp:31-36): same two functions, same argument order,
implemented on the C ABI of libboxseg_b200."""
import torch

from ... import _lib as L

_DTYPE = {torch.float32: 0, torch.float64: 1}


def _prep(logits):
    L.require_cuda(logits)
    if logits.dim() != 4 or logits.size(1) != 1:
        raise RuntimeError('logits must be [B,1,H,W]')
    if logits.dtype not in _DTYPE:
        raise RuntimeError('pairwise_nlog supports float32 / float64 (as AT_DISPATCH_FLOATING_TYPES)')
    return logits.shape[0], logits.shape[2], logits.shape[3], _DTYPE[logits.dtype]


def pairwise_nlog_forward(pairwise_size, pairwise_dilation, logits):
    B, H, W, dt = _prep(logits)
    K = pairwise_size * pairwise_size - 1
    out = torch.empty((B, K, H, W), dtype=logits.dtype, device=logits.device)
    if out.numel():
        with torch.cuda.device(logits.device):
            L.check(L.lib().bxs_pairwise_nlog_forward(L.ptr(logits), L.ptr(out), B, H, W, pairwise_size,
                                                      pairwise_dilation, dt, L.stream()), 'pairwise_nlog_forward')
    return out


def pairwise_nlog_backward(pairwise_size, pairwise_dilation, logits, pairwise, g_pairwise):
    """``pairwise`` (the saved forward output) is accepted for signature parity and not needed."""
    B, H, W, dt = _prep(logits)
    L.require_cuda(g_pairwise)
    g_logits = torch.empty_like(logits)
    if g_logits.numel():
        with torch.cuda.device(logits.device):
            L.check(L.lib().bxs_pairwise_nlog_backward(L.ptr(logits), L.ptr(g_pairwise), L.ptr(g_logits), B, H, W,
                                                       pairwise_size, pairwise_dilation, dt, L.stream()),
                    'pairwise_nlog_backward')
    return g_logits

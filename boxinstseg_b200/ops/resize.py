"""a18: bilinear resize on the C ABI -- what the heads call ``F.interpolate(..., mode='bilinear')`` for
(mmdet/models/utils/misc.py:75-86 ``_scale_target``; box2mask_head.py:232-233,300,315-317,323-324,329;
box_solov2_head.py:213,412-415; discobox_head.py:1201).  Same arithmetic as ATen's upsample_bilinear2d; the
backward is a deterministic gather (ATen's is an atomicAdd scatter)."""
import torch

from .. import _lib as L


class _Bilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W, align):
        xc = x.contiguous().float()
        L.require_cuda(xc)
        N, C, h, w = xc.shape
        out = torch.empty((N, C, H, W), dtype=torch.float32, device=xc.device)
        if out.numel():
            with torch.cuda.device(xc.device):
                L.check(L.lib().bxs_bilinear_resize_forward(L.ptr(xc), L.ptr(out), N * C, h, w, H, W, int(align), L.stream()),
                        'bilinear_resize_forward')
        ctx.shape, ctx.align = (N, C, h, w), align
        return out

    @staticmethod
    def backward(ctx, g):
        N, C, h, w = ctx.shape
        gc = g.contiguous().float()
        gi = torch.empty((N, C, h, w), dtype=torch.float32, device=gc.device)
        if gi.numel():
            with torch.cuda.device(gc.device):
                L.check(L.lib().bxs_bilinear_resize_backward(L.ptr(gc), L.ptr(gi), N * C, h, w, gc.shape[2], gc.shape[3],
                                                             int(ctx.align), L.stream()), 'bilinear_resize_backward')
        return gi, None, None, None


def bilinear_resize(x, size, align_corners=False):
    """x [N,C,h,w] float32 CUDA -> [N,C,size[0],size[1]]; F.interpolate(x, size, mode='bilinear', align_corners=...)."""
    H, W = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    if x.dim() == 3:
        return _Bilinear.apply(x.unsqueeze(1), H, W, bool(align_corners)).squeeze(1)
    return _Bilinear.apply(x, H, W, bool(align_corners))


def scale_target(t, size=(96, 96)):
    """``_scale_target`` of mmdet/models/utils/misc.py:75-86: [n,h,w] or [n,c,h,w] -> [n,c,*size]."""
    if t.dim() == 3:
        t = t.unsqueeze(1)
    return bilinear_resize(t, size, align_corners=False)

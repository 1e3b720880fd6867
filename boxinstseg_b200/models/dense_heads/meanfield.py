"""``MeanField`` -- drop-in for mmdet/models/dense_heads/discobox_head.py:585-651 (same constructor
arguments, ``forward(x, targets, inter_img_mask=None) -> (pseudo_label, valid)``, no_grad), running on
the bit-map mean-field kernels of libboxseg_b200."""
import numpy as np
import torch
import torch.nn as nn

from ... import _lib as L


class MeanField(nn.Module):
    def __init__(self, feature_map, kernel_size=3, require_grad=False, theta0=0.5, theta1=30, theta2=10, alpha0=3,
                 iter=20, base=0.45, gamma=0.01):
        super().__init__()
        self.require_grad = require_grad
        self.kernel_size = kernel_size
        self.theta0, self.theta1, self.theta2 = theta0, theta1, theta2
        self.alpha0, self.gamma, self.base, self.iter = alpha0, gamma, base, iter
        with torch.no_grad():
            fm = feature_map.contiguous().float()
            L.require_cuda(fm)
            B, C, h, w = fm.shape
            self.kernel = torch.empty((B, kernel_size * kernel_size, h, w), dtype=torch.float32, device=fm.device)
            with torch.cuda.device(fm.device):
                L.check(L.lib().bxs_meanfield_kernel(L.ptr(fm), L.ptr(self.kernel), B, C, h, w, kernel_size,
                                                     float(np.float32(2 * theta0 ** 2)), float(np.float32(2 * theta1 ** 2)),
                                                     float(alpha0), L.stream()), 'meanfield_kernel')
        # the four -log(U) constants with the reference's own float32 arithmetic (discobox_head.py:619-620,638)
        b = torch.tensor(float(base), dtype=torch.float32)
        q = torch.tensor([0.0, 1.0]) * (1 - b * 2) + b                     # x for bit 0 / bit 1
        self._neglog = torch.cat([-torch.log(q), -torch.log(1 - q)]).numpy().astype(np.float32).copy()

    @torch.no_grad()
    def forward(self, x, targets, inter_img_mask=None, obj_img=None):
        xs = x.contiguous().float()
        tg = targets.contiguous().float()
        L.require_cuda(xs, tg)
        n, _, h, w = xs.shape
        ret = torch.empty_like(xs)
        valid = torch.empty(n, dtype=torch.float32, device=xs.device)
        if n == 0:
            return ret, valid
        lib = L.lib()
        ws = torch.empty(max(lib.bxs_meanfield_workspace_bytes(n, h, w), 1), dtype=torch.uint8, device=xs.device)
        if self.kernel.shape[0] > 1:
            assert obj_img is not None, 'a multi-image kernel needs obj_img'
            obj_img = obj_img.to(device=xs.device, dtype=torch.int32).contiguous()
        iiu = None
        if inter_img_mask is not None:                                    # corr_loss: [n,2,h,w] (background, foreground), :616,643-644
            iiu = inter_img_mask.contiguous().float()
            L.require_cuda(iiu)
            assert tuple(iiu.shape) == (n, 2, h, w), iiu.shape
        with torch.cuda.device(xs.device):
            L.check(lib.bxs_meanfield_forward_inter(L.ptr(self.kernel), L.ptr(obj_img), L.ptr(xs), L.ptr(tg), L.ptr(iiu),
                                                    float(np.float32(self.gamma)), self._neglog.ctypes.data, L.ptr(ret),
                                                    L.ptr(valid), L.ptr(ws), n, h, w, self.kernel_size, self.iter,
                                                    L.stream()), 'meanfield_forward')
        return ret, valid

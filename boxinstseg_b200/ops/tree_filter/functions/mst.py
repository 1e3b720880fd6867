"""``mst(edge_index, edge_weight, vertex_count)`` -- mmdet/ops/tree_filter/functions/mst.py:9-20."""
from torch.autograd import Function

from .. import tree_filter_cuda as _C


class _MST(Function):
    @staticmethod
    def forward(ctx, edge_index, edge_weight, vertex_index):
        out = _C.mst_forward(edge_index, edge_weight, vertex_index)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        return None, None, None


mst = _MST.apply

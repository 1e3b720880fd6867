"""``refine(feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, low_tree)`` --
mmdet/ops/tree_filter/functions/refine.py:9-41: differentiable wrt feature_in always and wrt
edge_weight when ``low_tree`` is False."""
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import tree_filter_cuda as _C


class _Refine(Function):
    @staticmethod
    def forward(ctx, feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, low_tree):
        levels = _C.levels_of(sorted_index, sorted_parent)
        out, aggr, aggr_up, wsum, wsum_up = _C.refine_forward(feature_in, edge_weight, sorted_index, sorted_parent,
                                                              sorted_child, levels)
        ctx.save_for_backward(edge_weight, sorted_index, sorted_parent, sorted_child, out, aggr, aggr_up, wsum, wsum_up,
                              *levels)
        ctx.low_tree = low_tree
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (edge_weight, sorted_index, sorted_parent, sorted_child, out, aggr, aggr_up, wsum, wsum_up, lvl,
         nlv) = ctx.saved_tensors
        args = (None, edge_weight, sorted_index, sorted_parent, sorted_child, out, aggr, aggr_up, wsum, wsum_up,
                grad_output.contiguous())
        grad_feature = _C.refine_backward_feature(*args, levels=(lvl, nlv))
        grad_weight = None if ctx.low_tree else _C.refine_backward_weight(*args, levels=(lvl, nlv))
        return grad_feature, grad_weight, None, None, None, None


refine = _Refine.apply

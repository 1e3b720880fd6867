"""Stand-in for the reference's pybind module ``tree_filter_cuda``
(mmdet/ops/tree_filter/src/tree_filter.cpp:7-13): the same five functions with the same argument
order and return values, implemented on the C ABI of libboxseg_b200."""
import torch

from ... import _lib as L

_LEVELS = '_bxs_levels'        # (level_start, num_levels) riding on the sorted_index tensor object


def _i32(t):
    return t if t.dtype == torch.int32 and t.is_contiguous() else t.to(torch.int32).contiguous()


def _f32(t):
    return t if t.dtype == torch.float32 and t.is_contiguous() else t.float().contiguous()


def mst_forward(edge_index, edge_weight, vertex_count):
    ei, ew = _i32(edge_index), _f32(edge_weight)
    L.require_cuda(ei, ew)
    B, E = ei.shape[0], ei.shape[1]
    V = int(vertex_count)
    out = torch.empty((B, V - 1, 2), dtype=torch.int32, device=ei.device)
    lib = L.lib()
    ws = torch.empty(lib.bxs_mst_workspace_bytes(B, E, V), dtype=torch.uint8, device=ei.device)
    with torch.cuda.device(ei.device):
        L.check(lib.bxs_mst_forward(L.ptr(ei), L.ptr(ew), L.ptr(out), L.ptr(ws), B, E, V, L.stream()), 'mst_forward')
    return out


def bfs_forward(edge_index, max_adj_per_node):
    te = _i32(edge_index)
    L.require_cuda(te)
    B, V = te.shape[0], te.shape[1] + 1
    dev = te.device
    idx = torch.empty((B, V), dtype=torch.int32, device=dev)
    par = torch.empty((B, V), dtype=torch.int32, device=dev)
    chd = torch.empty((B, V, max_adj_per_node), dtype=torch.int32, device=dev)
    lvl = torch.empty((B, V + 1), dtype=torch.int32, device=dev)
    nlv = torch.empty(B, dtype=torch.int32, device=dev)
    lib = L.lib()
    ws = torch.empty(lib.bxs_bfs_workspace_bytes(B, V), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.bxs_bfs_forward(L.ptr(te), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv), L.ptr(ws), B, V,
                                    int(max_adj_per_node), L.stream()), 'bfs_forward')
    setattr(idx, _LEVELS, (lvl, nlv))
    return idx, par, chd


def levels_of(sorted_index, sorted_parent):
    """level boundaries of a level-contiguous order: cached by bfs_forward, else recomputed on the GPU."""
    cached = getattr(sorted_index, _LEVELS, None)
    if cached is not None:
        return cached
    par = _i32(sorted_parent)
    B, V = par.shape
    lvl = torch.empty((B, V + 1), dtype=torch.int32, device=par.device)
    nlv = torch.empty(B, dtype=torch.int32, device=par.device)
    scratch = torch.empty(B * V * 4, dtype=torch.int32, device=par.device)
    with torch.cuda.device(par.device):
        L.check(L.lib().bxs_tree_levels(L.ptr(par), L.ptr(lvl), L.ptr(nlv), L.ptr(scratch), B, V, L.stream()), 'tree_levels')
    return lvl, nlv


def _scratch(B, C, V, dev):
    return torch.empty(max(L.lib().bxs_refine_scratch_bytes(B, C, V), 4), dtype=torch.uint8, device=dev)


def refine_forward(feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, levels=None):
    f, w = _f32(feature_in), _f32(edge_weight)
    idx, par, chd = _i32(sorted_index), _i32(sorted_parent), _i32(sorted_child)
    L.require_cuda(f, w, idx, par, chd)
    lvl, nlv = levels if levels is not None else levels_of(sorted_index, sorted_parent)
    B, C, V = f.shape
    out, aggr, aggr_up = torch.empty_like(f), torch.empty_like(f), torch.empty_like(f)
    wsum = torch.empty((B, V), dtype=torch.float32, device=f.device)
    wsum_up = torch.empty_like(wsum)
    with torch.cuda.device(f.device):
        L.check(L.lib().bxs_refine_forward(L.ptr(f), L.ptr(w), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv),
                                           L.ptr(out), L.ptr(aggr), L.ptr(aggr_up), L.ptr(wsum), L.ptr(wsum_up),
                                           L.ptr(_scratch(B, C, V, f.device)), B, C, V, L.stream()), 'refine_forward')
    return out, aggr, aggr_up, wsum, wsum_up


def refine_backward_feature(feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, feature_out,
                            feature_aggr, feature_aggr_up, weight_sum, weight_sum_up, grad_out, levels=None):
    w, g = _f32(edge_weight), _f32(grad_out)
    idx, par, chd = _i32(sorted_index), _i32(sorted_parent), _i32(sorted_child)
    lvl, nlv = levels if levels is not None else levels_of(sorted_index, sorted_parent)
    B, C, V = g.shape
    gf = torch.empty_like(g)
    with torch.cuda.device(g.device):
        L.check(L.lib().bxs_refine_backward_feature(L.ptr(w), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv),
                                                    L.ptr(_f32(weight_sum)), L.ptr(g), L.ptr(gf),
                                                    L.ptr(_scratch(B, C, V, g.device)), B, C, V, L.stream()),
                'refine_backward_feature')
    return gf


def refine_backward_weight(feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, feature_out,
                           feature_aggr, feature_aggr_up, weight_sum, weight_sum_up, grad_out, levels=None):
    w, g = _f32(edge_weight), _f32(grad_out)
    idx, par, chd = _i32(sorted_index), _i32(sorted_parent), _i32(sorted_child)
    lvl, nlv = levels if levels is not None else levels_of(sorted_index, sorted_parent)
    B, C, V = g.shape
    gw = torch.empty((B, V), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        L.check(L.lib().bxs_refine_backward_weight(L.ptr(w), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv),
                                                   L.ptr(_f32(feature_out)), L.ptr(_f32(feature_aggr)),
                                                   L.ptr(_f32(feature_aggr_up)), L.ptr(_f32(weight_sum)),
                                                   L.ptr(_f32(weight_sum_up)), L.ptr(g), L.ptr(gw),
                                                   L.ptr(_scratch(B, C, V, g.device)), B, C, V, L.stream()),
                'refine_backward_weight')
    return gw

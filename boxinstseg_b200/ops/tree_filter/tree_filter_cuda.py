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


def bfs_forward(edge_index, max_adj_per_node, root=0):
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
        L.check(lib.bxs_bfs_forward_rooted(L.ptr(te), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv), L.ptr(ws), B, V,
                                           int(max_adj_per_node), int(root), L.stream()), 'bfs_forward')
    setattr(idx, _LEVELS, (lvl, nlv, idx._version))
    return idx, par, chd


def _cached_levels(sorted_index):
    """(level_start, num_levels) produced by bfs_forward for exactly this tensor in its current state, else None.
    The cache rides on the tensor object; it is trusted only while the tensor has not been written since
    (same `_version`), so a buffer that was modified in place or reused as an output is treated as foreign."""
    cached = getattr(sorted_index, _LEVELS, None)
    if cached is None or cached[2] != sorted_index._version:
        return None
    return cached[0], cached[1]


def levels_of(sorted_index, sorted_parent):
    """level boundaries of an order produced by bfs_forward (cached), else of the equivalent canonical order."""
    cached = _cached_levels(sorted_index)
    if cached is not None:
        return cached
    return _adopt(sorted_index, sorted_parent)[3]


def _adopt(sorted_index, sorted_parent):
    """Any parent-before-child order (e.g. the reference's own racy bfs_forward output, bfs.cu:46-98) -> the
    deterministic level order of the SAME tree that the refine kernels need (levels contiguous, children of a node
    adjacent): (idx, par, chd, (level_start, num_levels), perm) with perm[b, foreign position] = position here."""
    idx, par = _i32(sorted_index), _i32(sorted_parent)
    B, V = idx.shape
    up = torch.gather(idx, 1, par[:, 1:].long())                       # vertex id of each non-root node's parent
    edges = torch.stack((idx[:, 1:], up), dim=2).contiguous()          # the tree as an edge list [B,V-1,2]
    oidx, opar, ochd = bfs_forward(edges, 4)
    pos_of_vertex = torch.empty((B, V), dtype=torch.int64, device=idx.device)
    pos_of_vertex.scatter_(1, oidx.long(), torch.arange(V, device=idx.device).expand(B, V))
    perm = torch.gather(pos_of_vertex, 1, idx.long())
    return oidx, opar, ochd, _cached_levels(oidx), perm


def _resolve(edge_weight, sorted_index, sorted_parent, sorted_child, levels):
    """-> (w, idx, par, chd, levels, perm): the caller's order when it came from bfs_forward (perm None), else the
    adopted order with the per-position edge weights carried over by vertex."""
    idx, par, chd = _i32(sorted_index), _i32(sorted_parent), _i32(sorted_child)
    if levels is None:
        levels = _cached_levels(sorted_index)
    if levels is not None:
        return edge_weight, idx, par, chd, levels, None
    oidx, opar, ochd, levels, perm = _adopt(idx, par)
    w = torch.zeros_like(edge_weight)
    w.scatter_(1, perm, edge_weight)
    return w, oidx, opar, ochd, levels, perm


def _to_ours(t, perm):      # position-ordered [B,V] or [B,C,V] tensor: foreign order -> adopted order
    if perm is None or t is None:
        return t
    out = torch.empty_like(t)
    out.scatter_(t.dim() - 1, perm if t.dim() == 2 else perm.unsqueeze(1).expand_as(t), t)
    return out


def _to_theirs(t, perm):
    if perm is None:
        return t
    return torch.gather(t, t.dim() - 1, perm if t.dim() == 2 else perm.unsqueeze(1).expand_as(t))


def _scratch(B, C, V, dev):
    return torch.empty(max(L.lib().bxs_refine_scratch_bytes(B, C, V), 4), dtype=torch.uint8, device=dev)


def refine_forward(feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, levels=None):
    f, w = _f32(feature_in), _f32(edge_weight)
    L.require_cuda(f, w, sorted_index, sorted_parent, sorted_child)
    w, idx, par, chd, (lvl, nlv), perm = _resolve(w, sorted_index, sorted_parent, sorted_child, levels)
    B, C, V = f.shape
    out, aggr, aggr_up = torch.empty_like(f), torch.empty_like(f), torch.empty_like(f)
    wsum = torch.empty((B, V), dtype=torch.float32, device=f.device)
    wsum_up = torch.empty_like(wsum)
    with torch.cuda.device(f.device):
        L.check(L.lib().bxs_refine_forward(L.ptr(f), L.ptr(w), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv),
                                           L.ptr(out), L.ptr(aggr), L.ptr(aggr_up), L.ptr(wsum), L.ptr(wsum_up),
                                           L.ptr(_scratch(B, C, V, f.device)), B, C, V, L.stream()), 'refine_forward')
    # out / aggr / wsum are vertex-ordered; the two *_up tensors are position-ordered (refine.cu:201-260)
    return out, aggr, _to_theirs(aggr_up, perm), wsum, _to_theirs(wsum_up, perm)


def refine_backward_feature(feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, feature_out,
                            feature_aggr, feature_aggr_up, weight_sum, weight_sum_up, grad_out, levels=None):
    w, g = _f32(edge_weight), _f32(grad_out)
    w, idx, par, chd, (lvl, nlv), perm = _resolve(w, sorted_index, sorted_parent, sorted_child, levels)
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
    w, idx, par, chd, (lvl, nlv), perm = _resolve(w, sorted_index, sorted_parent, sorted_child, levels)
    feature_aggr_up, weight_sum_up = _to_ours(_f32(feature_aggr_up), perm), _to_ours(_f32(weight_sum_up), perm)
    B, C, V = g.shape
    gw = torch.empty((B, V), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        L.check(L.lib().bxs_refine_backward_weight(L.ptr(w), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv),
                                                   L.ptr(_f32(feature_out)), L.ptr(_f32(feature_aggr)),
                                                   L.ptr(_f32(feature_aggr_up)), L.ptr(_f32(weight_sum)),
                                                   L.ptr(_f32(weight_sum_up)), L.ptr(g), L.ptr(gw),
                                                   L.ptr(_scratch(B, C, V, g.device)), B, C, V, L.stream()),
                'refine_backward_weight')
    return _to_theirs(gw, perm)


# ----------------------------------------------------------------------------------------------------------------
# grouped forms (not part of the reference's pybind surface): n instances share G trees, tree_of [n] -> group
# ----------------------------------------------------------------------------------------------------------------
def refine_forward_grouped(feature_in, edge_weight, sorted_index, sorted_parent, sorted_child, levels, tree_of):
    f, w = _f32(feature_in), _f32(edge_weight)
    idx, par, chd = _i32(sorted_index), _i32(sorted_parent), _i32(sorted_child)
    L.require_cuda(f, w, idx, par, chd, tree_of)
    lvl, nlv = levels
    n, C, V = f.shape
    G = w.shape[0]
    out, aggr, aggr_up = torch.empty_like(f), torch.empty_like(f), torch.empty_like(f)
    wsum = torch.empty((G, V), dtype=torch.float32, device=f.device)
    wsum_up = torch.empty_like(wsum)
    with torch.cuda.device(f.device):
        L.check(L.lib().bxs_refine_forward_grouped(L.ptr(f), L.ptr(w), L.ptr(idx), L.ptr(par), L.ptr(chd), L.ptr(lvl), L.ptr(nlv),
                                                   L.ptr(tree_of), L.ptr(out), L.ptr(aggr), L.ptr(aggr_up), L.ptr(wsum),
                                                   L.ptr(wsum_up), L.ptr(_scratch(max(n, G), C, V, f.device)), n, G, C, V,
                                                   L.stream()), 'refine_forward_grouped')
    return out, aggr, aggr_up, wsum, wsum_up


def refine_backward_feature_grouped(edge_weight, sorted_index, sorted_parent, sorted_child, levels, tree_of, weight_sum,
                                    grad_out):
    w, g = _f32(edge_weight), _f32(grad_out)
    lvl, nlv = levels
    n, C, V = g.shape
    G = w.shape[0]
    gf = torch.empty_like(g)
    with torch.cuda.device(g.device):
        L.check(L.lib().bxs_refine_backward_feature_grouped(L.ptr(w), L.ptr(sorted_index), L.ptr(sorted_parent),
                                                            L.ptr(sorted_child), L.ptr(lvl), L.ptr(nlv), L.ptr(tree_of),
                                                            L.ptr(weight_sum), L.ptr(g), L.ptr(gf),
                                                            L.ptr(_scratch(max(n, G), C, V, g.device)), n, G, C, V, L.stream()),
                'refine_backward_feature_grouped')
    return gf


def refine_backward_weight_grouped(edge_weight, sorted_index, sorted_parent, sorted_child, levels, tree_of, feature_out,
                                   feature_aggr, feature_aggr_up, weight_sum, weight_sum_up, grad_out):
    w, g = _f32(edge_weight), _f32(grad_out)
    lvl, nlv = levels
    n, C, V = g.shape
    G = w.shape[0]
    gw = torch.empty((n, V), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        L.check(L.lib().bxs_refine_backward_weight_grouped(L.ptr(w), L.ptr(sorted_index), L.ptr(sorted_parent),
                                                           L.ptr(sorted_child), L.ptr(lvl), L.ptr(nlv), L.ptr(tree_of),
                                                           L.ptr(feature_out), L.ptr(feature_aggr), L.ptr(feature_aggr_up),
                                                           L.ptr(weight_sum), L.ptr(weight_sum_up), L.ptr(g), L.ptr(gw),
                                                           L.ptr(_scratch(max(n, G), C, V, g.device)), n, G, C, V, L.stream()),
                'refine_backward_weight_grouped')
    return gw

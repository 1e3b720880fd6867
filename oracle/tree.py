"""Oracle (test infrastructure only) for the tree-filter rows a12-a15 of SURVEY.md section 8.

Index/sequential work is in C (``oracle/csrc/tree_oracle.c`` -> ``liboracle_tree.so``); the
edge-weight arithmetic stays in torch exactly as the reference computes it.  When
``oracle/_ref/libboruvka_ref.so`` exists (the reference's own ``boruvka.cpp`` compiled where
it lies by ``oracle/Makefile``) it is exposed as ``mst_reference_boruvka`` to pin the
restatement.  Never imported by the product.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build():
    subprocess.run(['make', '-s', '-C', _HERE], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboracle_tree.so')
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_mst_kruskal.restype = ctypes.c_int
        _LIB.orc_bfs.restype = ctypes.c_int
    return _LIB


def _ip(a):
    return a.ctypes.data_as(_i32p)


def _fp(a):
    return a.ctypes.data_as(_f32p)


# ----------------------------------------------------------------------------------------
# a12: graph construction + MST
# ----------------------------------------------------------------------------------------
def grid_edges(h, w):
    """int32 [E,2]: vertical edges (v, v+w) row-major first, then horizontal (v, v+1).
    mmdet/ops/tree_filter/modules/tree_filter.py:15-25."""
    v = np.arange(h * w, dtype=np.int32).reshape(h, w)
    vert = np.stack([v[:-1, :], v[1:, :]], -1).reshape(-1, 2)
    hori = np.stack([v[:, :-1], v[:, 1:]], -1).reshape(-1, 2)
    return np.ascontiguousarray(np.concatenate([vert, hori], 0))


def grid_edge_weights(fm):
    """fm [B,C,H,W] -> float32 [B,E]: sum_c (f_a-f_b)^2 + 1 in the edge order of grid_edges.
    tree_filter.py:27-34 with norm2_distance (:78-82)."""
    b = fm.shape[0]
    dv = fm[:, :, :-1, :] - fm[:, :, 1:, :]
    dh = fm[:, :, :, :-1] - fm[:, :, :, 1:]
    wv = (dv * dv).sum(1).reshape(b, -1)
    wh = (dh * dh).sum(1).reshape(b, -1)
    return torch.cat([wv, wh], 1) + 1


def mst_edge_ids(edge_index, edge_weight, num_vertices):
    """Sorted int32 ids (into edge_index) of the unique MST under the (weight, id) order."""
    ei = np.ascontiguousarray(edge_index, dtype=np.int32)
    ew = np.ascontiguousarray(edge_weight, dtype=np.float32)
    out = np.empty(num_vertices - 1, dtype=np.int32)
    n = _lib().orc_mst_kruskal(_ip(ei), _fp(ew), num_vertices, ei.shape[0], _ip(out))
    assert n == num_vertices - 1, 'graph not connected'
    return np.sort(out)


class _RefEdge(ctypes.Structure):
    _fields_ = [('src', ctypes.c_int), ('dest', ctypes.c_int), ('weight', ctypes.c_float)]


class _RefGraph(ctypes.Structure):
    _fields_ = [('V', ctypes.c_int), ('E', ctypes.c_int), ('edge', ctypes.POINTER(_RefEdge))]


def have_reference_boruvka():
    return os.path.exists(os.path.join(_HERE, '_ref', 'libboruvka_ref.so'))


def mst_reference_boruvka(edge_index, edge_weight, num_vertices):
    """Runs the reference's own compiled boruvka.cpp (oracle/_ref); returns [V-1,2] int32 in
    the reference's discovery order (mst.cu:41-49,78-83 is what fills the Graph)."""
    global _REF
    if _REF is None:
        _REF = ctypes.CDLL(os.path.join(_HERE, '_ref', 'libboruvka_ref.so'))
        _REF._Z11createGraphii.restype = ctypes.POINTER(_RefGraph)
        _REF._Z11createGraphii.argtypes = [ctypes.c_int, ctypes.c_int]
        _REF._Z10boruvkaMSTP5GraphPi.argtypes = [ctypes.POINTER(_RefGraph), _i32p]
    ei = np.ascontiguousarray(edge_index, dtype=np.int32)
    ew = np.ascontiguousarray(edge_weight, dtype=np.float32)
    E = ei.shape[0]
    g = _REF._Z11createGraphii(num_vertices, E)
    packed = np.empty(E, dtype=[('src', 'i4'), ('dest', 'i4'), ('weight', 'f4')])
    packed['src'], packed['dest'], packed['weight'] = ei[:, 0], ei[:, 1], ew
    ctypes.memmove(g.contents.edge, packed.ctypes.data, packed.nbytes)
    out = np.zeros((num_vertices - 1, 2), dtype=np.int32)
    _REF._Z10boruvkaMSTP5GraphPi(g, _ip(out))
    return out            # (the Graph is leaked on purpose: the reference frees with delete[])


def edges_to_ids(tree_edges, h, w):
    """Map [V-1,2] (src,dest) grid edges back to ids in grid_edges order."""
    te = np.asarray(tree_edges, dtype=np.int64)
    a = np.minimum(te[:, 0], te[:, 1])
    b = np.maximum(te[:, 0], te[:, 1])
    vertical = (b - a) == w
    ids = np.where(vertical, a, (h - 1) * w + (a // w) * (w - 1) + (a % w))
    assert np.all(vertical | ((b - a) == 1))
    return np.sort(ids.astype(np.int32))


def mst(fm):
    """fm [B,C,H,W] float32 -> int32 [B,V-1,2] tree edges (ascending edge id)."""
    b, _, h, w = fm.shape
    ei = grid_edges(h, w)
    ew = grid_edge_weights(fm.float()).numpy()
    out = np.stack([ei[mst_edge_ids(ei, ew[i], h * w)] for i in range(b)])
    return torch.from_numpy(out)


# ----------------------------------------------------------------------------------------
# a13: BFS
# ----------------------------------------------------------------------------------------
def bfs(tree):
    """tree int32 [B,V-1,2] -> sorted_index [B,V], sorted_parent [B,V], sorted_child [B,V,4]."""
    t = np.ascontiguousarray(tree.numpy(), dtype=np.int32)
    b, vm1, _ = t.shape
    V = vm1 + 1
    idx = np.zeros((b, V), np.int32)
    par = np.zeros((b, V), np.int32)
    chd = np.zeros((b, V, 4), np.int32)
    for i in range(b):
        rc = _lib().orc_bfs(_ip(t[i]), V, _ip(idx[i]), _ip(par[i]), _ip(chd[i]))
        assert rc == 0
    return torch.from_numpy(idx), torch.from_numpy(par), torch.from_numpy(chd)


# ----------------------------------------------------------------------------------------
# a14: edge weights in sorted order
# ----------------------------------------------------------------------------------------
def build_edge_weight(embed, sorted_index, sorted_parent, low_tree, sigma=0.02):
    """w[pos] = exp(-|E(v_pos)-E(v_par(pos))|^2 / (sigma if low_tree else 1)); differentiable
    wrt embed.  tree_filter.py:91-108 (groups == 1)."""
    b, c = embed.shape[:2]
    fm = embed.reshape(b, c, -1)
    src = torch.gather(fm, 2, sorted_index.long()[:, None, :].expand(-1, c, -1))
    tgt = torch.gather(src, 2, sorted_parent.long()[:, None, :].expand(-1, c, -1))
    d = src - tgt
    dist = (d * d).sum(1)
    return torch.exp(-dist / sigma) if low_tree else torch.exp(-dist)


# ----------------------------------------------------------------------------------------
# a15: refine (autograd wrapper over the C passes)
# ----------------------------------------------------------------------------------------
class _Refine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, edge_weight, idx, par, chd, low_tree):
        f = np.ascontiguousarray(feature.detach().numpy(), dtype=np.float32)
        w = np.array(edge_weight.detach().numpy(), dtype=np.float32, copy=True)
        i_, p_, c_ = (np.ascontiguousarray(a.numpy(), dtype=np.int32) for a in (idx, par, chd))
        b, c, v = f.shape
        out, aggr, aggr_up = np.empty_like(f), np.empty_like(f), np.empty_like(f)
        ws, wsu = np.empty((b, v), np.float32), np.empty((b, v), np.float32)
        for i in range(b):
            _lib().orc_refine_forward(_fp(f[i]), _fp(w[i]), _ip(i_[i]), _ip(p_[i]), _ip(c_[i]),
                                      c, v, _fp(out[i]), _fp(aggr[i]), _fp(aggr_up[i]),
                                      _fp(ws[i]), _fp(wsu[i]))
        ctx.saved = (w, i_, p_, c_, out, aggr, aggr_up, ws, wsu)
        ctx.low_tree = low_tree
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_out):
        w, i_, p_, c_, out, aggr, aggr_up, ws, wsu = ctx.saved
        g = np.ascontiguousarray(grad_out.numpy(), dtype=np.float32)
        b, c, v = g.shape
        gf = np.empty_like(g)
        for i in range(b):
            _lib().orc_refine_backward_feature(_fp(g[i]), _fp(w[i]), _ip(i_[i]), _ip(p_[i]),
                                               _ip(c_[i]), _fp(ws[i]), c, v, _fp(gf[i]))
        gw = None
        if not ctx.low_tree:
            gw_np = np.empty((b, v), np.float32)
            for i in range(b):
                _lib().orc_refine_backward_weight(_fp(g[i]), _fp(w[i]), _ip(i_[i]), _ip(p_[i]),
                                                  _ip(c_[i]), _fp(out[i]), _fp(aggr[i]),
                                                  _fp(aggr_up[i]), _fp(ws[i]), _fp(wsu[i]),
                                                  c, v, _fp(gw_np[i]))
            gw = torch.from_numpy(gw_np)
        return torch.from_numpy(gf), gw, None, None, None, None


def refine(feature, edge_weight, idx, par, chd, low_tree):
    """feature [B,C,V] float32, edge_weight [B,V] -> [B,C,V].  functions/refine.py:9-41."""
    return _Refine.apply(feature, edge_weight, idx, par, chd, low_tree)


def tree_filter(feature_in, embed_in, tree, low_tree=True, sigma=0.02):
    """TreeFilter2D.forward (tree_filter.py:135-150), groups == 1."""
    shape = feature_in.shape
    idx, par, chd = bfs(tree)
    w = build_edge_weight(embed_in, idx, par, low_tree, sigma)
    out = refine(feature_in.reshape(shape[0], shape[1], -1), w, idx, par, chd, low_tree)
    return out.reshape(shape)


def tree_filter_f64(feature_in, embed_in, idx, par, low_tree=True, sigma=0.02):
    """Float64 arbiter of the same recursion (refine.cu:19-199 semantics) on a GIVEN
    parent-before-child order, differentiable by torch autograd wrt feature and embed (whatever low_tree is: callers
    compare d/d embed only where the reference propagates it).  Level by level, vectorised: any size the tests use.
        up   U[p] = x[p] + sum_children w[c] U[c]        down  A[0] = U[0]; A[p] = (1 - w^2) U[p] + w A[par]
        out = A / Z with Z the same for x == 1."""
    shape = feature_in.shape
    B, C = shape[0], shape[1]
    V = shape[2] * shape[3]
    idx_l, par_l = idx.long(), par.long()
    w_all = build_edge_weight(embed_in.double(), idx, par, low_tree, sigma)          # [B,V], entry 0 unused
    outs = []
    for b in range(B):
        ib, pb = idx_l[b], par_l[b]
        depth = torch.zeros(V, dtype=torch.long)
        for p in range(1, V):                                                         # parents precede children
            depth[p] = depth[pb[p]] + 1
        order = torch.argsort(depth, stable=True)
        bounds = torch.searchsorted(depth[order].contiguous(), torch.arange(int(depth.max()) + 2))
        levels = [order[bounds[l]:bounds[l + 1]] for l in range(int(depth.max()) + 1)]
        w = torch.cat([w_all[b, :1].detach() * 0, w_all[b, 1:]])
        x = torch.cat([feature_in[b].reshape(C, V).double()[:, ib], torch.ones(1, V, dtype=torch.float64)])   # + normaliser
        U = x
        for nodes in reversed(levels[1:]):                                            # deepest level first
            U = U.index_add(1, pb[nodes], U[:, nodes] * w[nodes])
        A = U
        for nodes in levels[1:]:
            wn = w[nodes]
            A = A.index_copy(1, nodes, (1 - wn * wn) * U[:, nodes] + wn * A[:, pb[nodes]])
        res = (A[:C] / A[C:]).new_zeros(C, V).index_copy(1, ib, A[:C] / A[C:])
        outs.append(res)
    return torch.stack(outs).reshape(shape)


def tree_filter_dense(feature, embed, tree, low_tree=True, sigma=0.02):
    """Closed form for tiny cases (pure python/numpy, float64):
    out[i] = sum_j prod_{edges on path i..j} w * x[j] / sum_j prod w.   SURVEY appendix A15."""
    b, c, h, w_ = feature.shape
    V = h * w_
    x = feature.reshape(b, c, V).double().numpy()
    e = embed.reshape(b, embed.shape[1], V).double().numpy()
    t = tree.numpy()
    out = np.zeros_like(x)
    for n in range(b):
        adj = [[] for _ in range(V)]
        for a, bb in t[n]:
            d = ((e[n, :, a] - e[n, :, bb]) ** 2).sum()
            wt = np.exp(-d / sigma) if low_tree else np.exp(-d)
            adj[a].append((bb, wt))
            adj[bb].append((a, wt))
        for i in range(V):
            prod = np.zeros(V)
            prod[i] = 1.0
            stack, seen = [i], {i}
            while stack:
                u = stack.pop()
                for (v2, wt) in adj[u]:
                    if v2 not in seen:
                        seen.add(v2)
                        prod[v2] = prod[u] * wt
                        stack.append(v2)
            out[n, :, i] = (x[n] * prod[None]).sum(1) / prod.sum()
    return torch.from_numpy(out.reshape(b, c, h, w_))

/*
 * CPU oracle (TEST INFRASTRUCTURE ONLY, never linked into the product) for the tree-filter
 * rows of SURVEY.md section 8: a12 (minimum spanning tree of the 4-connected grid), a13
 * (breadth-first ordering), a15 (two-pass tree aggregation, forward and both backwards).
 *
 * Independent restatement:
 *   - orc_mst_kruskal: the reference runs Boruvka (mmdet/ops/tree_filter/src/mst/boruvka.cpp:20-112)
 *     where every component picks its minimum edge under the strict total order
 *     (weight, edge index) (strict '>' at :75,:79).  Under a strict total order the MST is
 *     unique, so Kruskal over edges sorted by (weight, index) yields the identical edge SET.
 *   - orc_bfs: any breadth-first order rooted at vertex 0 is valid downstream (the reference's
 *     own order depends on atomics, src/bfs/bfs.cu:39-42,86); here children are visited in
 *     ascending vertex id so the oracle is deterministic.
 *   - orc_refine_*: src/refine/refine.cu:19-199 (kernels) and :201-370 (host wrappers).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float w; int32_t idx; } edge_key;

static int cmp_edge(const void* a, const void* b) {
    const edge_key* x = (const edge_key*)a;
    const edge_key* y = (const edge_key*)b;
    if (x->w < y->w) return -1;
    if (x->w > y->w) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static int32_t uf_find(int32_t* p, int32_t v) {
    while (p[v] != v) { p[v] = p[p[v]]; v = p[v]; }
    return v;
}

/* edge_index [E,2] int32, edge_weight [E] float, out edge ids chosen [V-1] (ascending by key).
 * returns number of tree edges written. */
int orc_mst_kruskal(const int32_t* edge_index, const float* edge_weight, int32_t V, int32_t E,
                    int32_t* out_edge_ids) {
    edge_key* keys = (edge_key*)malloc(sizeof(edge_key) * (size_t)E);
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)V);
    for (int32_t i = 0; i < E; ++i) { keys[i].w = edge_weight[i]; keys[i].idx = i; }
    for (int32_t v = 0; v < V; ++v) parent[v] = v;
    qsort(keys, (size_t)E, sizeof(edge_key), cmp_edge);
    int32_t n = 0;
    for (int32_t i = 0; i < E && n < V - 1; ++i) {
        int32_t e = keys[i].idx;
        int32_t a = uf_find(parent, edge_index[2 * e]);
        int32_t b = uf_find(parent, edge_index[2 * e + 1]);
        if (a == b) continue;
        parent[a] = b;
        out_edge_ids[n++] = e;
    }
    free(keys);
    free(parent);
    return n;
}

/* tree edges [V-1,2] -> sorted_index[V], sorted_parent[V] (positions), sorted_child[V,4]
 * (positions, 0-terminated).  Root = vertex 0 at position 0.  returns 0, or -1 if a vertex has
 * more than 4 neighbours / the edges do not span. */
int orc_bfs(const int32_t* tree_edges, int32_t V, int32_t* sorted_index, int32_t* sorted_parent,
            int32_t* sorted_child) {
    int32_t* deg = (int32_t*)calloc((size_t)V, sizeof(int32_t));
    int32_t* adj = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)V);
    int32_t* par_vertex = (int32_t*)malloc(sizeof(int32_t) * (size_t)V);
    int rc = 0;
    for (int32_t i = 0; i < V - 1; ++i) {
        int32_t a = tree_edges[2 * i], b = tree_edges[2 * i + 1];
        if (deg[a] >= 4 || deg[b] >= 4) { rc = -1; goto done; }
        adj[4 * a + deg[a]++] = b;
        adj[4 * b + deg[b]++] = a;
    }
    /* ascending neighbour ids (insertion sort of <=4) */
    for (int32_t v = 0; v < V; ++v)
        for (int i = 1; i < deg[v]; ++i)
            for (int j = i; j > 0 && adj[4 * v + j] < adj[4 * v + j - 1]; --j) {
                int32_t t = adj[4 * v + j]; adj[4 * v + j] = adj[4 * v + j - 1]; adj[4 * v + j - 1] = t;
            }
    memset(sorted_child, 0, sizeof(int32_t) * 4 * (size_t)V);
    sorted_index[0] = 0; sorted_parent[0] = 0; par_vertex[0] = -1;
    int32_t len = 1;
    for (int32_t pos = 0; pos < len; ++pos) {
        int32_t v = sorted_index[pos];
        int32_t nc = 0;
        for (int i = 0; i < deg[v]; ++i) {
            int32_t u = adj[4 * v + i];
            if (u == par_vertex[pos]) continue;
            sorted_index[len] = u;
            sorted_parent[len] = pos;
            par_vertex[len] = v;
            sorted_child[4 * pos + nc++] = len;
            ++len;
        }
    }
    if (len != V) rc = -1;
done:
    free(deg); free(adj); free(par_vertex);
    return rc;
}

/* up pass: U[pos] = x[v_pos] (or 1 when x==NULL) + sum_child w[child] U[child]  (sorted order)
 * down pass: A[v_0] = U[0]; A[v_pos] = (1-w^2) U[pos] + w A[v_par]             (vertex order)
 * refine.cu:70-134 and :19-68 */
static void up_pass(const float* x, const float* w, const int32_t* idx, const int32_t* child,
                    int32_t V, float* U) {
    for (int32_t pos = V - 1; pos >= 0; --pos) {
        float acc = x ? x[idx[pos]] : 1.0f;
        for (int j = 0; j < 4; ++j) {
            int32_t c = child[4 * pos + j];
            if (c <= 0) break;
            acc += U[c] * w[c];
        }
        U[pos] = acc;
    }
}

static void down_pass(const float* U, const float* w, const int32_t* idx, const int32_t* par,
                      int32_t V, float* A) {
    A[idx[0]] = U[0];
    for (int32_t pos = 1; pos < V; ++pos) {
        float ew = w[pos];
        A[idx[pos]] = U[pos] * (1.0f - ew * ew) + A[idx[par[pos]]] * ew;
    }
}

/* feature [C,V] (vertex order), w [V] (sorted order, w[0] ignored -> 0).
 * outputs: out [C,V] = A/Z, aggr A [C,V] (vertex order), aggr_up U [C,V] (sorted order),
 *          wsum Z [V] (vertex order), wsum_up [V] (sorted order).  refine.cu:201-249 */
void orc_refine_forward(const float* feature, float* w, const int32_t* idx, const int32_t* par,
                        const int32_t* child, int32_t C, int32_t V, float* out, float* aggr,
                        float* aggr_up, float* wsum, float* wsum_up) {
    w[0] = 0.0f;
    for (int32_t c = 0; c < C; ++c) {
        up_pass(feature + (size_t)c * V, w, idx, child, V, aggr_up + (size_t)c * V);
        down_pass(aggr_up + (size_t)c * V, w, idx, par, V, aggr + (size_t)c * V);
    }
    up_pass(NULL, w, idx, child, V, wsum_up);
    down_pass(wsum_up, w, idx, par, V, wsum);
    for (int32_t c = 0; c < C; ++c)
        for (int32_t v = 0; v < V; ++v) out[(size_t)c * V + v] = aggr[(size_t)c * V + v] / wsum[v];
}

/* grad wrt feature: the same filter applied to g/Z.  refine.cu:251-300 */
void orc_refine_backward_feature(const float* grad_out, const float* w, const int32_t* idx,
                                 const int32_t* par, const int32_t* child, const float* wsum,
                                 int32_t C, int32_t V, float* grad_feature) {
    float* gn = (float*)malloc(sizeof(float) * (size_t)V);
    float* up = (float*)malloc(sizeof(float) * (size_t)V);
    for (int32_t c = 0; c < C; ++c) {
        for (int32_t v = 0; v < V; ++v) gn[v] = grad_out[(size_t)c * V + v] / wsum[v];
        up_pass(gn, w, idx, child, V, up);
        down_pass(up, w, idx, par, V, grad_feature + (size_t)c * V);
    }
    free(gn); free(up);
}

/* one root->leaf gradient sweep (refine.cu:136-199): for pos>0
 *   grad[pos] = gup[pos]*(outd[v_par] - w*ind[pos]) + ind[pos]*(G[par] - w*gup[pos])
 *   G[pos]    = gup[pos]*(1-w^2) + G[par]*w      with G[0] = gup[0]
 * ind = up-pass data (sorted order), outd = down-pass data (vertex order), gup = up pass of
 * the upstream gradient (sorted order).  */
static void grad_sweep(const float* ind, const float* gup, const float* outd, const float* w,
                       const int32_t* idx, const int32_t* par, int32_t V, float* grad) {
    float* G = (float*)malloc(sizeof(float) * (size_t)V);
    G[0] = gup[0];
    grad[0] = 0.0f;
    for (int32_t pos = 1; pos < V; ++pos) {
        float ew = w[pos];
        int32_t p = par[pos];
        float left = gup[pos] * (outd[idx[p]] - ew * ind[pos]);
        float right = ind[pos] * (G[p] - ew * gup[pos]);
        grad[pos] = left + right;
        G[pos] = gup[pos] * (1.0f - ew * ew) + G[p] * ew;
    }
    free(G);
}

/* grad wrt edge weight [V] (sorted order).  refine.cu:302-370 */
void orc_refine_backward_weight(const float* grad_out, const float* w, const int32_t* idx,
                                const int32_t* par, const int32_t* child, const float* out,
                                const float* aggr, const float* aggr_up, const float* wsum,
                                const float* wsum_up, int32_t C, int32_t V, float* grad_w) {
    float* gn = (float*)malloc(sizeof(float) * (size_t)V);
    float* fg = (float*)malloc(sizeof(float) * (size_t)V);
    float* gn_up = (float*)malloc(sizeof(float) * (size_t)V);
    float* fg_up = (float*)malloc(sizeof(float) * (size_t)V);
    float* g1 = (float*)malloc(sizeof(float) * (size_t)V);
    float* g2 = (float*)malloc(sizeof(float) * (size_t)V);
    for (int32_t v = 0; v < V; ++v) grad_w[v] = 0.0f;
    for (int32_t c = 0; c < C; ++c) {
        for (int32_t v = 0; v < V; ++v) {
            gn[v] = grad_out[(size_t)c * V + v] / wsum[v];
            fg[v] = gn[v] * out[(size_t)c * V + v];
        }
        up_pass(gn, w, idx, child, V, gn_up);
        up_pass(fg, w, idx, child, V, fg_up);
        grad_sweep(aggr_up + (size_t)c * V, gn_up, aggr + (size_t)c * V, w, idx, par, V, g1);
        grad_sweep(wsum_up, fg_up, wsum, w, idx, par, V, g2);
        for (int32_t v = 0; v < V; ++v) grad_w[v] += g1[v] - g2[v];
    }
    free(gn); free(fg); free(gn_up); free(fg_up); free(g1); free(g2);
}

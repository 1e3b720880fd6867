// a6 + a7 + a8: the fused BoxInst mask loss (projection term + colour-pairwise term) and its
// backward.  Replaces the arithmetic of CondInstMaskHead.loss, condinst_head.py:1288-1343
// (sigmoid -> compute_project_term :134-143 -> pairwise_nlog -> threshold weights -> reduction).
//
// B200-first design (DESIGN.md "BoxInst loss"):
//   * the [N,8,H,W] pairwise tensor, the [N,8,H,W] gathered similarity and the [N,1,H,W] gathered
//     bitmask of the reference are never materialised.  The similarity enters as ONE BYTE per
//     pixel per image (bit c = sim[c] >= thresh) and the box bitmask as 4 integers per GT.
//   * forward kernel 1 streams the logits once (HBM-bound): row / column maxima of the scores
//     with first-index tie-breaking, combined across CTAs with order-independent 64-bit
//     atomicMax on (score bits, ~index) -> deterministic.
//   * forward kernel 2 evaluates pair terms ONLY where a weight can be non-zero (inside the GT
//     box +- dilation), walking a device-built work list of 16x64 tiles so every CTA has work.
//     Each unordered pair is evaluated once (pl is symmetric) with an integer multiplier 0/1/2.
//   * the finalize kernel turns the profiles into dice losses and into per-row/per-column
//     gradient coefficients; the last CTA (ticket counter) writes the scalars.
//   * backward is one tile kernel writing every g_logits element exactly once: zeros + arg-max
//     scatter outside the box, gather-form pairwise gradient inside.  No atomics on floats.
#include <algorithm>

#include "common.cuh"

namespace bxs {
namespace {

constexpr int TH = 16, TW = 64, NT = 256;
constexpr float kFastLimit = 40.f;
constexpr float kDiceEps = 1e-5f;        // condinst_head.py:124

struct Workspace {
  // zeroed at the start of every forward
  unsigned long long* row_packed;  // [N*H]  (score bits << 32) | ~x
  unsigned long long* col_packed;  // [N*W]  (score bits << 32) | ~y
  unsigned long long* weight_sum;  // [1]    sum of weights (integer)
  unsigned int* ticket;            // [1]
  size_t zero_bytes;
  // plain
  float* coef_row;     // [N*H] d loss_prj / d logit at the row arg-max
  float* coef_col;     // [N*W]
  int* row_arg;        // [N*H]
  int* col_arg;        // [N*W]
  int* tile_prefix;    // [N+1]
  float* pair_partial; // [N*tiles_full]
  float* inst_prj;     // [N]
  float* inst_num;     // [N]
  float* scale_pair;   // [1] warmup / max(weight_sum, 1)
  size_t total_bytes;
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

inline Workspace carve(void* base, int64_t N, int64_t H, int64_t W) {
  Workspace w{};
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p + off; off = align_up(off + bytes); return r; };
  w.row_packed = (unsigned long long*)take(sizeof(unsigned long long) * N * H);
  w.col_packed = (unsigned long long*)take(sizeof(unsigned long long) * N * W);
  w.weight_sum = (unsigned long long*)take(sizeof(unsigned long long));
  w.ticket = (unsigned int*)take(sizeof(unsigned int));
  w.zero_bytes = off;
  w.coef_row = (float*)take(sizeof(float) * N * H);
  w.coef_col = (float*)take(sizeof(float) * N * W);
  w.row_arg = (int*)take(sizeof(int) * N * H);
  w.col_arg = (int*)take(sizeof(int) * N * W);
  w.tile_prefix = (int*)take(sizeof(int) * (N + 1));
  const int64_t tiles_full = ceil_div(H, TH) * ceil_div(W, TW);
  w.pair_partial = (float*)take(sizeof(float) * N * tiles_full);
  w.inst_prj = (float*)take(sizeof(float) * N);
  w.inst_num = (float*)take(sizeof(float) * N);
  w.scale_pair = (float*)take(sizeof(float));
  w.total_bytes = off;
  return w;
}

struct Rect { int j0, j1, i0, i1; };

__device__ __forceinline__ Rect load_rect(const int32_t* rects, int g) {
  int4 r = *reinterpret_cast<const int4*>(rects + 4 * (int64_t)g);
  return Rect{r.x, r.y, r.z, r.w};
}
__device__ __forceinline__ bool rect_empty(const Rect& r) { return r.j0 > r.j1 || r.i0 > r.i1; }
__device__ __forceinline__ bool in_rect(const Rect& r, int y, int x) {
  return y >= r.j0 && y <= r.j1 && x >= r.i0 && x <= r.i1;
}

// tiles (map aligned) that contain a pixel p for which pair (p, p+forward offset) can carry weight:
// rows [j0-d, j1], cols [i0-d, i1+d], clipped to the map.
__device__ __forceinline__ bool fwd_tile_range(const Rect& r, int d, int H, int W, int& ty0, int& ty1,
                                               int& tx0, int& tx1) {
  if (rect_empty(r)) return false;
  int y0 = max(r.j0 - d, 0), y1 = min(r.j1, H - 1), x0 = max(r.i0 - d, 0), x1 = min(r.i1 + d, W - 1);
  if (y0 > y1 || x0 > x1) return false;
  ty0 = y0 / TH; ty1 = y1 / TH; tx0 = x0 / TW; tx1 = x1 / TW;
  return true;
}

// ---------------------------------------------------------------------------------------
// work list: exclusive prefix of per-instance tile counts (one CTA)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) prep_worklist(const int32_t* __restrict__ rects,
                                                    const int32_t* __restrict__ inst_gt, int N, int H, int W,
                                                    int d, int* __restrict__ tile_prefix) {
  __shared__ int s_sum[NT];
  const int per = (N + NT - 1) / NT;
  const int lo = threadIdx.x * per, hi = min(lo + per, N);
  int local = 0;
  for (int n = lo; n < hi; ++n) {
    int a, b, c, e;
    if (fwd_tile_range(load_rect(rects, inst_gt[n]), d, H, W, a, b, c, e)) local += (b - a + 1) * (e - c + 1);
  }
  s_sum[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {     // NT partial sums: a serial scan is cheaper than it looks (256 adds)
    int run = 0;
    for (int i = 0; i < NT; ++i) { int v = s_sum[i]; s_sum[i] = run; run += v; }
    tile_prefix[N] = run;
  }
  __syncthreads();
  int run = s_sum[threadIdx.x];
  for (int n = lo; n < hi; ++n) {
    tile_prefix[n] = run;
    int a, b, c, e;
    if (fwd_tile_range(load_rect(rects, inst_gt[n]), d, H, W, a, b, c, e)) run += (b - a + 1) * (e - c + 1);
  }
}

// ---------------------------------------------------------------------------------------
// forward 1: projection maxima.  grid (strips, N, panels); one warp per row.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_max(float score, int index) {
  return ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - (unsigned)index);
}

template <int NCHUNK, int V>
__global__ void __launch_bounds__(NT) prj_max_kernel(const float* __restrict__ logits, int H, int W,
                                                     int rows_per_block,
                                                     unsigned long long* __restrict__ row_packed,
                                                     unsigned long long* __restrict__ col_packed) {
  constexpr int PANEL = NCHUNK * 32 * V;
  constexpr int NWARP = NT / 32;
  __shared__ unsigned long long s_col[NWARP][PANEL];
  const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, H);
  const int panel0 = blockIdx.z * PANEL;
  const float* img = logits + (int64_t)n * H * W;

  float cbest[NCHUNK * V];
  int cy[NCHUNK * V];
#pragma unroll
  for (int i = 0; i < NCHUNK * V; ++i) { cbest[i] = -1.f; cy[i] = 0; }

  for (int y = r0 + warp; y < r1; y += NWARP) {
    const float* row = img + (int64_t)y * W;
    float v[NCHUNK * V];
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {           // issue all loads of the row first
      const int col = panel0 + (ch * 32 + lane) * V;
      if constexpr (V == 4) {
        float4 q = col < W ? __ldg(reinterpret_cast<const float4*>(row + col)) : make_float4(0, 0, 0, 0);
        v[ch * 4 + 0] = q.x; v[ch * 4 + 1] = q.y; v[ch * 4 + 2] = q.z; v[ch * 4 + 3] = q.w;
      } else {
        v[ch] = col < W ? __ldg(row + col) : 0.f;
      }
    }
    float rbest = -1.f;
    int rx = 0;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int col = panel0 + (ch * 32 + lane) * V + e;
        if (col < W) {
          const float s = sigmoid_fast(v[ch * V + e]);
          if (s > rbest) { rbest = s; rx = col; }                       // ascending cols: first wins
          if (s > cbest[ch * V + e]) { cbest[ch * V + e] = s; cy[ch * V + e] = y; }   // ascending rows
        }
      }
    }
    unsigned long long rp = rbest >= 0.f ? pack_max(rbest, rx) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(kFull, rp, o);
      rp = other > rp ? other : rp;
    }
    if (lane == 0 && rp) atomicMax(row_packed + (int64_t)n * H + y, rp);
  }
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch)
#pragma unroll
    for (int e = 0; e < V; ++e)
      s_col[warp][(ch * 32 + lane) * V + e] = cbest[ch * V + e] >= 0.f ? pack_max(cbest[ch * V + e], cy[ch * V + e]) : 0ull;
  __syncthreads();
  for (int c = threadIdx.x; c < PANEL; c += NT) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) best = s_col[w][c] > best ? s_col[w][c] : best;
    if (best && panel0 + c < W) atomicMax(col_packed + (int64_t)n * W + panel0 + c, best);
  }
}

// ---------------------------------------------------------------------------------------
// shared-memory tile of (x, s, n, edge byte) with an asymmetric halo
// ---------------------------------------------------------------------------------------
struct PairTile {
  float *x, *s, *n;
  uint8_t* e;
  int pitch;
};

__device__ __forceinline__ PairTile load_pair_tile(const float* __restrict__ img, const uint8_t* __restrict__ bits,
                                                   int H, int W, int y_lo, int x_lo, int ph, int pw,
                                                   float* smem) {
  PairTile t{smem, smem + ph * pw, smem + 2 * ph * pw, reinterpret_cast<uint8_t*>(smem + 3 * ph * pw), pw};
  for (int i = threadIdx.x; i < ph * pw; i += NT) {
    int ly = i / pw, lx = i - ly * pw;
    int gy = y_lo + ly, gx = x_lo + lx;
    float v = 0.f, s = 0.f, n = 0.f;
    uint8_t e = 0;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      v = __ldg(img + (int64_t)gy * W + gx);
      e = __ldg(bits + (int64_t)gy * W + gx);
      sigmoid_pair(v, s, n);
    }
    t.x[i] = v; t.s[i] = s; t.n[i] = n; t.e[i] = e;
  }
  __syncthreads();
  return t;
}

// ---------------------------------------------------------------------------------------
// forward 2: pairwise numerators over the work list
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) pair_fwd_kernel(const float* __restrict__ logits,
                                                      const uint8_t* __restrict__ edge_bits,
                                                      const int32_t* __restrict__ rects,
                                                      const int32_t* __restrict__ inst_gt,
                                                      const int32_t* __restrict__ gt_img, int N, int H, int W,
                                                      int d, const int* __restrict__ tile_prefix,
                                                      float* __restrict__ pair_partial,
                                                      unsigned long long* __restrict__ weight_sum) {
  extern __shared__ float smem[];
  __shared__ float s_red[NT / 32];
  __shared__ int s_redi[NT / 32];
  const int total = tile_prefix[N];
  const int ph = TH + d, pw = TW + 2 * d;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    int lo = 0, hi = N;                       // largest n with tile_prefix[n] <= item
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tile_prefix[mid] <= item) lo = mid; else hi = mid;
    }
    const int n = lo;
    const int g = inst_gt[n];
    const Rect r = load_rect(rects, g);
    int ty0, ty1, tx0, tx1;
    fwd_tile_range(r, d, H, W, ty0, ty1, tx0, tx1);
    const int local = item - tile_prefix[n];
    const int ntx = tx1 - tx0 + 1;
    const int y0 = (ty0 + local / ntx) * TH, x0 = (tx0 + local % ntx) * TW;
    const float* img = logits + (int64_t)n * H * W;
    const uint8_t* bits = edge_bits + (int64_t)gt_img[g] * H * W;
    __syncthreads();                          // previous item's readers are done with smem
    PairTile t = load_pair_tile(img, bits, H, W, y0, x0 - d, ph, pw, smem);

    float acc = 0.f;
    int wsum = 0;
    const int tx = threadIdx.x % TW;
    const int gx = x0 + tx;
    for (int ty = threadIdx.x / TW; ty < TH; ty += NT / TW) {
      const int gy = y0 + ty;
      if (gy >= H || gx >= W) continue;
      const int ci = ty * t.pitch + tx + d;
      const bool pin = in_rect(r, gy, gx);
      const unsigned ep = t.e[ci];
      if (pin) wsum += __popc(ep);
      const float xa = t.x[ci], sa = t.s[ci], na = t.n[ci];
      // forward half of the neighbourhood: c = 4..7 <-> (0,+d), (+d,-d), (+d,0), (+d,+d)
#pragma unroll
      for (int c = 4; c < 8; ++c) {
        const int dy = c == 4 ? 0 : d;
        const int dx = c == 4 ? d : (c - 6) * d;
        const int qy = gy + dy, qx = gx + dx;
        if (qy >= H || qx < 0 || qx >= W) continue;
        const int qi = ci + (c == 4 ? 0 : d * t.pitch) + dx;
        const unsigned eq = t.e[qi];
        const int m = (pin ? (ep >> c) & 1u : 0u) + (in_rect(r, qy, qx) ? (eq >> (7 - c)) & 1u : 0u);
        if (m) {
          const float xb = t.x[qi];
          float pl;
          if (fmaxf(fabsf(xa), fabsf(xb)) <= kFastLimit)
            pl = -__logf(sa * t.s[qi] + na * t.n[qi]);
          else
            pl = pair_nlog_logspace<float>(xa, xb, true);
          acc = fmaf((float)m, pl, acc);
        }
      }
    }
    const float tot = block_sum<float>(acc, s_red);
    const int wtot = block_sum<int>(wsum, s_redi);
    if (threadIdx.x == 0) {
      pair_partial[item] = tot;
      if (wtot) atomicAdd(weight_sum, (unsigned long long)wtot);
    }
  }
}

// ---------------------------------------------------------------------------------------
// finalize: dice terms, gradient coefficients, scalars (last CTA by ticket)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) finalize_kernel(const int32_t* __restrict__ rects,
                                                      const int32_t* __restrict__ inst_gt, int N, int H, int W,
                                                      Workspace ws, const float* __restrict__ iter_ptr,
                                                      float warmup_iters, float* __restrict__ losses_out) {
  __shared__ float s_red[NT / 32];
  __shared__ float s_bc[4];
  __shared__ bool s_last;
  const int n = blockIdx.x;
  const Rect r = load_rect(rects, inst_gt[n]);
  const bool empty = rect_empty(r);
  const float inv_n = 1.f / (float)N;
  float inst_loss = 0.f;
  // axis 0: row profile (max over x) of length H;  axis 1: column profile of length W
  for (int axis = 0; axis < 2; ++axis) {
    const int L = axis == 0 ? H : W;
    const int lo = axis == 0 ? r.j0 : r.i0, hi = axis == 0 ? r.j1 : r.i1;
    const unsigned long long* packed = (axis == 0 ? ws.row_packed + (int64_t)n * H : ws.col_packed + (int64_t)n * W);
    float* coef = axis == 0 ? ws.coef_row + (int64_t)n * H : ws.coef_col + (int64_t)n * W;
    int* arg = axis == 0 ? ws.row_arg + (int64_t)n * H : ws.col_arg + (int64_t)n * W;
    float inter = 0.f, x2 = 0.f;
    for (int i = threadIdx.x; i < L; i += NT) {
      const float s = __uint_as_float((unsigned)(packed[i] >> 32));
      const bool t = !empty && i >= lo && i <= hi;
      inter += t ? s : 0.f;
      x2 = fmaf(s, s, x2);
    }
    inter = block_sum<float>(inter, s_red);
    x2 = block_sum<float>(x2, s_red);
    if (threadIdx.x == 0) {
      const float t2 = empty ? 0.f : (float)(max(min(hi, L - 1) - max(lo, 0) + 1, 0));
      const float u = x2 + t2 + kDiceEps;
      s_bc[0] = inter; s_bc[1] = u;
      inst_loss += 1.f - 2.f * inter / u;
    }
    __syncthreads();
    const float I = s_bc[0], U = s_bc[1];
    for (int i = threadIdx.x; i < L; i += NT) {
      const unsigned long long p = packed[i];
      const float s = __uint_as_float((unsigned)(p >> 32));
      const float t = (!empty && i >= lo && i <= hi) ? 1.f : 0.f;
      // d dice / d s = -2 t / U + 4 I s / U^2 ; through the sigmoid: * s (1 - s); mean over N
      coef[i] = inv_n * (-2.f * t / U + 4.f * I * s / (U * U)) * s * (1.f - s);
      arg[i] = (int)(0xffffffffu - (unsigned)(p & 0xffffffffull));
    }
    __syncthreads();
  }
  // pairwise numerator of this instance: fixed-order sum of its tile partials
  float num = 0.f;
  for (int i = ws.tile_prefix[n] + threadIdx.x; i < ws.tile_prefix[n + 1]; i += NT) num += ws.pair_partial[i];
  num = block_sum<float>(num, s_red);
  if (threadIdx.x == 0) {
    ws.inst_prj[n] = inst_loss;
    ws.inst_num[n] = num;
    __threadfence();
    s_last = atomicAdd(ws.ticket, 1u) == (unsigned)(N - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float prj = 0.f, pn = 0.f;
  for (int i = threadIdx.x; i < N; i += NT) { prj += __ldcg(ws.inst_prj + i); pn += __ldcg(ws.inst_num + i); }
  prj = block_sum<float>(prj, s_red);
  pn = block_sum<float>(pn, s_red);
  if (threadIdx.x == 0) {
    const float wsum = (float)__ldcg(ws.weight_sum);
    const float warm = fminf(iter_ptr[0] / warmup_iters, 1.f);
    const float scale = warm / fmaxf(wsum, 1.f);
    losses_out[0] = prj * inv_n;
    losses_out[1] = pn * scale;
    losses_out[2] = pn;
    losses_out[3] = wsum;
    ws.scale_pair[0] = scale;
  }
}

// ---------------------------------------------------------------------------------------
// backward: one tile kernel, every g_logits element written exactly once
// ---------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(NT) loss_bwd_kernel(const float* __restrict__ logits,
                                                      const uint8_t* __restrict__ edge_bits,
                                                      const int32_t* __restrict__ rects,
                                                      const int32_t* __restrict__ inst_gt,
                                                      const int32_t* __restrict__ gt_img, int H, int W, int d,
                                                      Workspace ws, const float* __restrict__ g_losses,
                                                      float* __restrict__ g_logits) {
  extern __shared__ float smem[];
  const int n = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const int g = inst_gt[n];
  const Rect r = load_rect(rects, g);
  const float g_prj = g_losses[0];
  const float g_pair = g_losses[1] * ws.scale_pair[0];
  // pixels that can receive a pairwise gradient: the box grown by d
  const bool box_tile = !rect_empty(r) && y0 <= r.j1 + d && y0 + TH - 1 >= r.j0 - d && x0 <= r.i1 + d &&
                        x0 + TW - 1 >= r.i0 - d;
  PairTile t{};
  if (box_tile) {
    const float* img = logits + (int64_t)n * H * W;
    const uint8_t* bits = edge_bits + (int64_t)gt_img[g] * H * W;
    t = load_pair_tile(img, bits, H, W, y0 - d, x0 - d, TH + 2 * d, TW + 2 * d, smem);
  }
  const int ty = threadIdx.x / (TW / 4), tx = (threadIdx.x % (TW / 4)) * 4;
  const int gy = y0 + ty, gx = x0 + tx;
  if (gy >= H || gx >= W) return;
  const float* crow = ws.coef_row + (int64_t)n * H;
  const int* arow = ws.row_arg + (int64_t)n * H;
  const float* ccol = ws.coef_col + (int64_t)n * W;
  const int* acol = ws.col_arg + (int64_t)n * W;
  const int ra = arow[gy];
  const float rc = crow[gy] * g_prj;
  float out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = gx + k;
    float v = 0.f;
    if (x < W) {
      if (x == ra) v += rc;
      if (acol[x] == gy) v += ccol[x] * g_prj;
      if (box_tile) {
        const int ci = (ty + d) * t.pitch + (tx + k + d);
        const bool pin = in_rect(r, gy, x);
        const unsigned ep = t.e[ci];
        const float xa = t.x[ci], sa = t.s[ci], na = t.n[ci];
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cc = c < 4 ? c : c + 1;              // skip the centre of the 3x3 stencil
          const int dy = (cc / 3 - 1) * d, dx = (cc % 3 - 1) * d;
          const int qy = gy + dy, qx = x + dx;
          if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
          const int qi = ci + dy * t.pitch + dx;
          const unsigned eq = t.e[qi];
          const int m = (pin ? (ep >> c) & 1u : 0u) + (in_rect(r, qy, qx) ? (eq >> (7 - c)) & 1u : 0u);
          if (m) {
            const float xb = t.x[qi];
            float dd;
            if (fmaxf(fabsf(xa), fabsf(xb)) <= kFastLimit) {
              const float sb = t.s[qi], nb = t.n[qi];
              dd = -(sb - nb) * __frcp_rn(sa * sb + na * nb);
              dd *= sa * na;
            } else {
              dd = pair_nlog_grad_a_logspace<float>(xa, xb, true, pair_nlog_logspace<float>(xa, xb, true));
            }
            acc = fmaf((float)m, dd, acc);
          }
        }
        v = fmaf(acc, g_pair, v);
      }
    }
    out[k] = v;
  }
  float* dst = g_logits + (int64_t)n * H * W + (int64_t)gy * W + gx;
  if (VEC) {
    *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (gx + k < W) dst[k] = out[k];
  }
}

inline bool args_ok(int64_t N, int64_t H, int64_t W, int d) {
  return N > 0 && N < 65536 && H > 0 && W > 0 && H * W < (int64_t(1) << 31) && d >= 1 && d <= 16;
}

}  // namespace
}  // namespace bxs

using namespace bxs;

extern "C" int64_t bxs_boxinst_loss_workspace_bytes(int64_t N, int64_t H, int64_t W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)carve(nullptr, N, H, W).total_bytes;
}

extern "C" int bxs_boxinst_loss_forward(const float* logits, const uint8_t* edge_bits, const int32_t* rects,
                                        const int32_t* inst_gt, const int32_t* gt_img, const float* iter_ptr,
                                        float warmup_iters, void* workspace, float* losses_out, int64_t N,
                                        int64_t H, int64_t W, int dilation, bxs_stream_t stream) {
  if (!logits || !edge_bits || !rects || !inst_gt || !gt_img || !iter_ptr || !workspace || !losses_out ||
      !args_ok(N, H, W, dilation) || !(warmup_iters > 0.f))
    return BXS_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  Workspace ws = carve(workspace, N, H, W);
  const int d = dilation;
  cudaMemsetAsync(workspace, 0, ws.zero_bytes, st);
  prep_worklist<<<1, NT, 0, st>>>(rects, inst_gt, (int)N, (int)H, (int)W, d, ws.tile_prefix);

  // forward 1: ~4 CTAs per SM in flight, at least 8 rows (one per warp) per CTA
  const int sms = sm_count();
  int strips = (int)ceil_div((int64_t)sms * 4, N);
  strips = (int)std::max<int64_t>(1, std::min<int64_t>(strips, ceil_div(H, 8)));
  const int rows = (int)ceil_div(H, strips);
  strips = (int)ceil_div(H, rows);
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec && W <= 128) {
    prj_max_kernel<1, 4><<<dim3(strips, (unsigned)N, 1), NT, 0, st>>>(logits, (int)H, (int)W, rows, ws.row_packed, ws.col_packed);
  } else if (vec && W <= 256) {
    prj_max_kernel<2, 4><<<dim3(strips, (unsigned)N, 1), NT, 0, st>>>(logits, (int)H, (int)W, rows, ws.row_packed, ws.col_packed);
  } else if (vec) {
    prj_max_kernel<4, 4><<<dim3(strips, (unsigned)N, (unsigned)ceil_div(W, 512)), NT, 0, st>>>(logits, (int)H, (int)W, rows, ws.row_packed, ws.col_packed);
  } else {
    prj_max_kernel<8, 1><<<dim3(strips, (unsigned)N, (unsigned)ceil_div(W, 256)), NT, 0, st>>>(logits, (int)H, (int)W, rows, ws.row_packed, ws.col_packed);
  }
  int rc = check_launch();
  if (rc) return rc;

  // forward 2: persistent CTAs over the tile work list
  const size_t sm2 = (size_t)(TH + d) * (TW + 2 * d) * (3 * sizeof(float) + 1) + 16;
  pair_fwd_kernel<<<sms * 4, NT, sm2, st>>>(logits, edge_bits, rects, inst_gt, gt_img, (int)N, (int)H, (int)W, d,
                                            ws.tile_prefix, ws.pair_partial, ws.weight_sum);
  rc = check_launch();
  if (rc) return rc;
  finalize_kernel<<<(unsigned)N, NT, 0, st>>>(rects, inst_gt, (int)N, (int)H, (int)W, ws, iter_ptr, warmup_iters,
                                              losses_out);
  return check_launch();
}

extern "C" int bxs_boxinst_loss_backward(const float* logits, const uint8_t* edge_bits, const int32_t* rects,
                                         const int32_t* inst_gt, const int32_t* gt_img, const void* workspace,
                                         const float* g_losses, float* g_logits, int64_t N, int64_t H,
                                         int64_t W, int dilation, bxs_stream_t stream) {
  if (!logits || !edge_bits || !rects || !inst_gt || !gt_img || !workspace || !g_losses || !g_logits ||
      !args_ok(N, H, W, dilation))
    return BXS_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  Workspace ws = carve(const_cast<void*>(workspace), N, H, W);
  const int d = dilation;
  const size_t sm = (size_t)(TH + 2 * d) * (TW + 2 * d) * (3 * sizeof(float) + 1) + 16;
  dim3 grid((unsigned)ceil_div(W, TW), (unsigned)ceil_div(H, TH), (unsigned)N);
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(g_logits) & 15) == 0);
  if (vec)
    loss_bwd_kernel<true><<<grid, NT, sm, st>>>(logits, edge_bits, rects, inst_gt, gt_img, (int)H, (int)W, d, ws,
                                                g_losses, g_logits);
  else
    loss_bwd_kernel<false><<<grid, NT, sm, st>>>(logits, edge_bits, rects, inst_gt, gt_img, (int)H, (int)W, d, ws,
                                                 g_losses, g_logits);
  return check_launch();
}

// a6 + a7 + a8, single pass: the fused BoxInst mask loss AND its gradient from ONE read of the logits.
// Same arithmetic as boxinst_loss.cu (CondInstMaskHead.loss, condinst_head.py:1288-1343), different
// schedule: the pairwise normaliser (sum of weights) does not depend on the logits and the projection
// term back-propagates only to H + W arg-max positions per instance, so the dense part of the gradient
// can be produced while the logits stream through the SM the first time.
//
//   onepass_main_kernel   persistent CTAs, dynamic tile scheduler (one atomic per work item).  A work
//       item is a strip of 16 rows of one instance: ONE cp.async.bulk (TMA, 1-D) brings the strip
//       (+ the dilation halo when the strip meets the box) into shared memory, double buffered on
//       mbarriers, so the next strip is in flight while this one is processed.  From shared memory:
//       row maxima (integer redux), column maxima, zero stores of the gradient outside the box
//       span, and -- inside the span -- the pair terms and the RAW (unscaled) pairwise gradient,
//       gathered from a (sigmoid, 1 - sigmoid, edge bits) tile.  HBM traffic: logits read once,
//       gradient written once.
//   onepass_finalize_kernel   one CTA per instance: dice terms and their gradient coefficients, the
//       global weight sum; the last CTA writes the losses.  Does not touch the gradient.
//   onepass_backward_kernel   one CTA per instance, in place: the (small) box span of the gradient is
//       scaled by g_pair * warmup / weights and the projection terms are added at the H + W arg-max
//       positions.
//
// Everything is summed in a fixed order: results do not depend on which CTA processed which strip.
#include <algorithm>

#include "boxinst_common.cuh"

namespace bxs {
namespace {

constexpr int OP_NT = 256;          // threads per CTA of the main kernel
constexpr int OP_NW = OP_NT / 32;
constexpr int OP_R = 16;            // rows per strip
constexpr int OP_FIN_NT = 512;      // threads per CTA of the finalize / backward kernels
constexpr int OP_MAX_N = 2048;      // instance records live in shared memory (12 B each)

struct OpWorkspace {
  unsigned long long* row_packed;  // [N*H]    (key(max logit) << 32) | ~x
  unsigned long long* col_part;    // [N*S*W]  per strip (key(max logit) << 32) | ~y
  float* num_part;                 // [N*S]    per strip pairwise numerator
  int* den_part;                   // [N*S]    per strip weight count
  float* coef_row;                 // [N*H]    d loss_prj / d logit at the row arg-max
  float* coef_col;                 // [N*W]
  int* arg_row;                    // [N*H]
  int* arg_col;                    // [N*W]
  int* span;                       // [N*4]    y_lo, y_hi, c_lo, c_hi of the gradient span (y_lo > y_hi: none)
  float* inst_prj;                 // [N]
  float* inst_num;                 // [N]
  float* scale;                    // [1]      warmup / max(weight sum, 1)
  size_t total_bytes;
};

inline size_t op_align(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

inline OpWorkspace op_carve(void* base, int64_t N, int64_t H, int64_t W) {
  OpWorkspace w{};
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p + off; off = op_align(off + bytes); return r; };
  const int64_t S = ceil_div(H, OP_R);
  w.row_packed = (unsigned long long*)take(8 * N * H);
  w.col_part = (unsigned long long*)take(8 * N * S * W);
  w.num_part = (float*)take(4 * N * S);
  w.den_part = (int*)take(4 * N * S);
  w.coef_row = (float*)take(4 * N * H);
  w.coef_col = (float*)take(4 * N * W);
  w.arg_row = (int*)take(4 * N * H);
  w.arg_col = (int*)take(4 * N * W);
  w.span = (int*)take(16 * N);
  w.inst_prj = (float*)take(4 * N);
  w.inst_num = (float*)take(4 * N);
  w.scale = (float*)take(4);
  w.total_bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk copy (TMA engine, no tensor map needed for a contiguous strip)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t op_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void op_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(op_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void op_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(op_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void op_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "OP_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra OP_DONE;\n\t"
      "bra OP_WAIT;\n\t"
      "OP_DONE:\n\t"
      "}" ::"r"(op_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void op_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   op_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(op_smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------
// instance records
// ---------------------------------------------------------------------------------------
struct SRec { short j0, j1, i0, i1; int img; };     // box rectangle clipped to the map; canonical empty = (1,0,1,0)

__device__ __forceinline__ SRec make_srec(const int32_t* __restrict__ rects, const int32_t* __restrict__ inst_gt,
                                          const int32_t* __restrict__ gt_img, int n, int H, int W) {
  const int g = __ldg(inst_gt + n);
  const int4 r = __ldg(reinterpret_cast<const int4*>(rects + 4 * (int64_t)g));
  int j0 = max(r.x, 0), j1 = min(r.y, H - 1), i0 = max(r.z, 0), i1 = min(r.w, W - 1);
  if (r.x > r.y || r.z > r.w || j0 > j1 || i0 > i1) { j0 = 1; j1 = 0; i0 = 1; i1 = 0; }
  SRec v;
  v.j0 = (short)j0; v.j1 = (short)j1; v.i0 = (short)i0; v.i1 = (short)i1;
  v.img = __ldg(gt_img + g);
  return v;
}

// pixels that can receive a pairwise gradient: the box grown by D, columns rounded out to float4 groups
struct OpSpan { int y_lo, y_hi, c_lo, c_hi; };
template <int D>
__device__ __forceinline__ OpSpan op_span(const SRec& r, int H, int W) {
  OpSpan s;
  if (r.j0 > r.j1) { s.y_lo = 1; s.y_hi = 0; s.c_lo = 4; s.c_hi = 3; return s; }
  s.y_lo = max(r.j0 - D, 0);
  s.y_hi = min(r.j1 + D, H - 1);
  s.c_lo = max(r.i0 - D, 0) & ~3;
  s.c_hi = min(r.i1 + D, W - 1) | 3;      // < W because W % 4 == 0
  return s;
}

__device__ __forceinline__ int op_div(int a, int b) {      // exact for 0 <= a < 2^20, 0 < b < 2^10
  return __float2int_rz((__int2float_rn(a) + 0.5f) * rcp_approx(__int2float_rn(b)));
}
__device__ __forceinline__ float op_lg2(float v) {         // MUFU.LG2 (v is a normal number here)
  float o;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(o) : "f"(v));
  return o;
}
__device__ __forceinline__ float op_ex2(float v) {
  float o;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(o) : "f"(v));
  return o;
}
// (sigmoid(x), sigmoid(-x)), both to a few ulp relative (no 1 - s cancellation); exact enough for |x| <= kFastLimit
__device__ __forceinline__ void op_sigmoid_pair(float x, float& s, float& n) {
  const float e = op_ex2(-1.4426950408889634f * fabsf(x));
  const float big = rcp_approx(1.f + e);
  const float small = e * big;
  s = x >= 0.f ? big : small;
  n = x >= 0.f ? small : big;
}

// ---------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------
struct OpSched { unsigned next, done, ticket, pad; };

template <int D>
struct OpTile {
  static constexpr int ROWS = OP_R + 2 * D;      // rows of one staged strip (halo above / strip / halo below)
  static constexpr int TW = 64;                  // pitch of the pair tile: two columns per lane
  static constexpr int TC = TW - 2 * D;          // owner columns per tile
};

// FULLW: W == NCHUNK * 128 (every lane of a row warp owns NCHUNK full float4 groups; W folds to a constant)
template <int NCHUNK, int D, bool FULLW>
__global__ void __launch_bounds__(OP_NT, NCHUNK <= 2 ? 4 : 2)
onepass_main_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ edge_bits,
                    const int32_t* __restrict__ rects, const int32_t* __restrict__ inst_gt,
                    const int32_t* __restrict__ gt_img, int N, int H, int W_rt, int S, OpWorkspace ws,
                    OpSched* __restrict__ sched, float* __restrict__ g_logits) {
  constexpr int ROWS = OpTile<D>::ROWS, TW = OpTile<D>::TW, TC = OpTile<D>::TC;
  constexpr int TROWS_PER_WARP = (ROWS + OP_NW - 1) / OP_NW;
  const int W = FULLW ? NCHUNK * 128 : W_rt;
  extern __shared__ __align__(128) unsigned char op_smem[];
  const int stage_floats = ROWS * W;
  float* xbuf = reinterpret_cast<float*>(op_smem);                               // [2][ROWS][W]
  float2* t_sn = reinterpret_cast<float2*>(xbuf + 2 * stage_floats);             // [ROWS][TW] (sigmoid, 1 - sigmoid)
  uint8_t* t_e = reinterpret_cast<uint8_t*>(t_sn + ROWS * TW);                   // [ROWS][TW] effective edge bits
  SRec* s_rec = reinterpret_cast<SRec*>(t_e + ROWS * TW);                        // [N]
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ int s_item[2];
  __shared__ float s_redf[OP_NW];
  __shared__ int s_redi[OP_NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = N * S;

  // strip `item` -> stage: ONE bulk copy; the D halo rows above / below ride along when the strip meets the box span
  auto issue = [&](int item, int stage, bool force_halo) {
    const int n = op_div(item, S), s = item - n * S, y0 = s * OP_R;
    const int rows = min(OP_R, H - y0);
    bool halo = force_halo;
    if (!halo) {
      const OpSpan sp = op_span<D>(s_rec[n], H, W);
      halo = sp.y_lo <= sp.y_hi && y0 <= sp.y_hi && y0 + rows - 1 >= sp.y_lo;
    }
    const int top = (halo && y0 >= D) ? D : 0;
    const int bot = halo ? min(D, H - (y0 + rows)) : 0;
    const uint32_t bytes = (uint32_t)(top + rows + bot) * (uint32_t)W * 4u;
    op_mbar_expect_tx(&s_bar[stage], bytes);
    op_bulk_g2s(xbuf + (size_t)stage * stage_floats + (size_t)(D - top) * W,
                logits + ((int64_t)n * H + (y0 - top)) * W, bytes, &s_bar[stage]);
  };

  unsigned fetched = 0;          // thread 0: next dynamic item (fetched one iteration ahead of its use)
  if (tid == 0) {
    op_mbar_init(&s_bar[0], 1);
    op_mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int first = blockIdx.x;
    s_item[0] = first;
    if (first < total) issue(first, 0, true);
    fetched = gridDim.x + atomicAdd(&sched->next, 1u);
  }
  for (int n = tid; n < N; n += OP_NT) s_rec[n] = make_srec(rects, inst_gt, gt_img, n, H, W);
  __syncthreads();

  for (int k = 0;; ++k) {
    const int stage = k & 1;
    const int item = s_item[stage];
    if (item >= total) break;
    if (tid == 0) {              // prefetch item k + 1 into the other stage (its readers finished before the last barrier)
      const int nxt = (int)min(fetched, (unsigned)total);
      s_item[stage ^ 1] = nxt;
      if (nxt < total) {
        issue(nxt, stage ^ 1, false);
        fetched = gridDim.x + atomicAdd(&sched->next, 1u);
      }
    }
    const int n = op_div(item, S), s = item - n * S, y0 = s * OP_R;
    const int rows = min(OP_R, H - y0);
    const SRec rec = s_rec[n];
    const OpSpan sp = op_span<D>(rec, H, W);
    const int ya = max(y0, sp.y_lo), yb = min(y0 + rows - 1, sp.y_hi);
    const bool has_pair = ya <= yb;
    float* ginst = g_logits + (int64_t)n * H * W;
    const uint8_t* bits = edge_bits + (int64_t)rec.img * H * W;
    const int nrow = yb - ya + 1, trows = nrow + 2 * D;   // tile row tr <-> map row ya - D + tr <-> xbuf row (ya - y0) + tr

    // edge bytes of the first pair tile: issued now, consumed after the streaming passes (hides the L2 latency)
    unsigned eb[TROWS_PER_WARP][2];
    auto load_bits = [&](int cx0, int ccols) {
#pragma unroll
      for (int j = 0; j < TROWS_PER_WARP; ++j) {
        const int yy = ya - D + warp + j * OP_NW;
        const bool rowok = warp + j * OP_NW < trows && yy >= rec.j0 && yy <= rec.j1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int xx = cx0 - D + lane + 32 * h;
          eb[j][h] = (rowok && lane + 32 * h < ccols + 2 * D && xx >= rec.i0 && xx <= rec.i1) ? (unsigned)__ldg(bits + yy * W + xx) : 0u;
        }
      }
    };
    if (has_pair) load_bits(sp.c_lo, min(TC, sp.c_hi - sp.c_lo + 1));

    op_mbar_wait(&s_bar[stage], (k >> 1) & 1);
    const float* xb = xbuf + (size_t)stage * stage_floats;       // strip row r at xb[(D + r) * W + col]

    // ---- row maxima (one warp per row) + zero stores outside the span ----
    for (int r = warp; r < rows; r += OP_NW) {
      const int y = y0 + r;
      const float* row = xb + (D + r) * W;
      float v[NCHUNK * 4];
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const int col0 = (ch * 32 + lane) * 4;
        float4 q;
        if (FULLW || col0 < W) q = *reinterpret_cast<const float4*>(row + col0);
        else q = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        v[ch * 4] = q.x; v[ch * 4 + 1] = q.y; v[ch * 4 + 2] = q.z; v[ch * 4 + 3] = q.w;
      }
      float m = v[0];
#pragma unroll
      for (int i = 1; i < NCHUNK * 4; ++i) m = fmaxf(m, v[i]);
      const unsigned kmax = __reduce_max_sync(kFull, fkey(m));
      const float mv = fkey_inv(kmax);
      int cand = 0x7fffffff;
#pragma unroll
      for (int i = NCHUNK * 4 - 1; i >= 0; --i)
        if (v[i] == mv) cand = ((i >> 2) * 32 + lane) * 4 + (i & 3);
      const int amin = __reduce_min_sync(kFull, cand);
      if (lane == 0) ws.row_packed[n * H + y] = pack_key(kmax, amin == 0x7fffffff ? 0 : amin);
      float* grow = ginst + y * W + lane * 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_pair && y >= ya && y <= yb) {                      // warp-uniform: a span row
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
          const int col0 = (ch * 32 + lane) * 4;
          if ((FULLW || col0 < W) && (col0 < sp.c_lo || col0 > sp.c_hi)) *reinterpret_cast<float4*>(grow + ch * 128) = z;
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch)
          if (FULLW || (ch * 32 + lane) * 4 < W) *reinterpret_cast<float4*>(grow + ch * 128) = z;
      }
    }
    // ---- column maxima of the strip (one thread per column; first row wins ties) ----
    for (int col = tid; col < W; col += OP_NT) {
      const float* p = xb + D * W + col;
      float best = p[0];
      int brow = 0;
      if (rows == OP_R) {
#pragma unroll
        for (int r = 1; r < OP_R; ++r) {
          const float val = p[r * W];
          if (val > best) { best = val; brow = r; }
        }
      } else {
        for (int r = 1; r < rows; ++r) {
          const float val = p[r * W];
          if (val > best) { best = val; brow = r; }
        }
      }
      ws.col_part[(n * S + s) * W + col] = pack_key(fkey(best), y0 + brow);
    }

    // ---- pair terms of the span rows [ya, yb]: tiles of TC owner columns, two columns per lane ----
    float acc_v = 0.f;           // sum of w * lg2(denominator) (scaled by -ln 2 at the end)
    float acc_slow = 0.f;        // slow-path values (natural log)
    int acc_w = 0;
    if (has_pair) {              // CTA-uniform
      const int lr0 = ya - y0;
      for (int cx0 = sp.c_lo; cx0 <= sp.c_hi; cx0 += TC) {
        const int ccols = min(TC, sp.c_hi - cx0 + 1);
        const int twc = ccols + 2 * D;
        __syncthreads();                               // the previous tile has been consumed
        bool ext = false;
#pragma unroll
        for (int j = 0; j < TROWS_PER_WARP; ++j) {
          const int tr = warp + j * OP_NW;
          if (tr < trows) {                            // warp-uniform
            const int yy = ya - D + tr;
            const bool rowok = yy >= 0 && yy < H;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int lc = lane + 32 * h, xx = cx0 - D + lc;
              if (lc < twc) {
                float sg = 1.f, ng = 1.f;              // outside the map: -log(s_a * 1 + n_a * 1) = 0 (padded neighbour)
                if (rowok && xx >= 0 && xx < W) {
                  const float x = xb[(lr0 + tr) * W + xx];
                  op_sigmoid_pair(x, sg, ng);
                  ext |= fabsf(x) > kFastLimit;
                }
                t_sn[tr * TW + lc] = make_float2(sg, ng);
                t_e[tr * TW + lc] = (uint8_t)eb[j][h];
              }
            }
          }
        }
        ext = __syncthreads_or(ext);
        if (cx0 + TC <= sp.c_hi) load_bits(cx0 + TC, min(TC, sp.c_hi - cx0 - TC + 1));   // next tile's edge bytes
        for (int ry = warp; ry < nrow; ry += OP_NW) {
          const int y = ya + ry;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cx = lane + 32 * h;
            if (cx < ccols) {
              const int ci = (ry + D) * TW + cx + D;
              const float2 a = t_sn[ci];
              const unsigned ea = t_e[ci];
              const int x = cx0 + cx;
              float g = 0.f;
              if (!ext) {
                float lg = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const int cc = c < 4 ? c : c + 1;
                  const int dy = (cc / 3 - 1) * D, dx = (cc % 3 - 1) * D;
                  const float2 q = t_sn[ci + dy * TW + dx];
                  const unsigned eq = t_e[ci + dy * TW + dx];
                  const bool pa = (ea & (1u << c)) != 0u, pq = (eq & (1u << (7 - c))) != 0u;
                  const float den = fmaf(a.x, q.x, a.y * q.y);
                  const float t = (q.y - q.x) * rcp_approx(den);
                  const float l = op_lg2(den);
                  lg += pa ? l : 0.f;
                  g += pa ? t : 0.f;
                  g += pq ? t : 0.f;
                }
                acc_v += lg;
                g *= a.x * a.y;
              } else {                                     // a logit beyond +-40 in the tile: log-space formulas
                const float xa = xb[(lr0 + ry + D) * W + x];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const int cc = c < 4 ? c : c + 1;
                  const int dy = (cc / 3 - 1) * D, dx = (cc % 3 - 1) * D;
                  const unsigned eq = t_e[ci + dy * TW + dx];
                  const unsigned wa = (ea >> c) & 1u;
                  const unsigned mm = wa + ((eq >> (7 - c)) & 1u);
                  if (mm) {
                    const int qy = y + dy, qx = x + dx;
                    const bool has_b = qy >= 0 && qy < H && qx >= 0 && qx < W;
                    const float xq = has_b ? xb[(lr0 + ry + D) * W + x + dy * W + dx] : 0.f;
                    const float pl = pair_nlog_logspace<float>(xa, xq, has_b);
                    acc_slow = fmaf((float)wa, pl, acc_slow);
                    g = fmaf((float)mm, pair_nlog_grad_a_logspace<float>(xa, xq, has_b, pl), g);
                  }
                }
              }
              acc_w += __popc(ea);
              ginst[y * W + x] = g;
            }
          }
        }
      }
      // per-strip partials: fixed-order block reduction
      const float v = warp_sum(fmaf(acc_v, -0.69314718055994531f, acc_slow));
      const int w = warp_sum(acc_w);
      if (lane == 0) { s_redf[warp] = v; s_redi[warp] = w; }
    }
    __syncthreads();             // stage `stage` and the tile are free; s_item[stage ^ 1] is visible
    if (tid == 0) {
      float v = 0.f;
      int w = 0;
      if (has_pair) {
#pragma unroll
        for (int i = 0; i < OP_NW; ++i) { v += s_redf[i]; w += s_redi[i]; }
      }
      ws.num_part[item] = v;
      ws.den_part[item] = w;
    }
  }
  // ---- scheduler reset by the last CTA (every fetch of a CTA precedes its `done` increment) ----
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&sched->done, 1u) == gridDim.x - 1) {
      sched->next = 0u;
      sched->done = 0u;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------
// finalize: one CTA per instance.  Profiles -> dice terms and gradient coefficients; global weight sum -> scale;
// the last CTA (ticket) reduces the per-instance terms in a fixed order and writes the losses.  Does not touch
// the gradient.  All global loads of a phase are issued before the first use.
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(OP_FIN_NT)
onepass_finalize_kernel(const int32_t* __restrict__ rects, const int32_t* __restrict__ inst_gt,
                        const int32_t* __restrict__ gt_img, int N, int H, int W, int S, OpWorkspace ws,
                        OpSched* __restrict__ sched, const float* __restrict__ iter_ptr, float warmup_iters,
                        float* __restrict__ losses_out) {
  constexpr int NWF = OP_FIN_NT / 32;
  __shared__ float s_f[4][NWF];
  __shared__ int s_i[NWF];
  __shared__ bool s_last;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // independent loads first: this thread's row key, its column keys of every strip, weight counts, the record
  const int row_i = tid, col_i = tid;                   // H, W <= 512 = OP_FIN_NT
  unsigned long long rp = 0ull, cp = 0ull;
  if (row_i < H) rp = ws.row_packed[n * H + row_i];
  if (col_i < W) {
    const unsigned long long* src = ws.col_part + (int64_t)n * S * W + col_i;
#pragma unroll 4
    for (int s = 0; s < S; ++s) {
      const unsigned long long p = src[s * W];
      cp = p > cp ? p : cp;
    }
  }
  int den = 0;
  for (int i = tid; i < N * S; i += OP_FIN_NT) den += ws.den_part[i];
  float num = 0.f;
  if (wid == 0 && lane < S) num = ws.num_part[n * S + lane];        // S <= 32
  const float warm = fminf(iter_ptr[0] / warmup_iters, 1.f);
  const SRec rec = make_srec(rects, inst_gt, gt_img, n, H, W);
  const OpSpan sp = op_span<D>(rec, H, W);
  const bool empty = rec.j0 > rec.j1;

  const float sr = row_i < H ? sigmoid_exact(fkey_inv((unsigned)(rp >> 32))) : 0.f;
  const float sc = col_i < W ? sigmoid_exact(fkey_inv((unsigned)(cp >> 32))) : 0.f;
  const bool tr = !empty && row_i >= rec.j0 && row_i <= rec.j1, tc = !empty && col_i >= rec.i0 && col_i <= rec.i1;
  float r0 = warp_sum(tr ? sr : 0.f), r1 = warp_sum(sr * sr), r2 = warp_sum(tc ? sc : 0.f), r3 = warp_sum(sc * sc);
  den = warp_sum(den);
  if (lane == 0) { s_f[0][wid] = r0; s_f[1][wid] = r1; s_f[2][wid] = r2; s_f[3][wid] = r3; s_i[wid] = den; }
  __syncthreads();
  float Ir = 0.f, Xr = 0.f, Ic = 0.f, Xc = 0.f;
  int wtot = 0;
#pragma unroll
  for (int i = 0; i < NWF; ++i) { Ir += s_f[0][i]; Xr += s_f[1][i]; Ic += s_f[2][i]; Xc += s_f[3][i]; wtot += s_i[i]; }
  const float inv_n = 1.f / (float)N;
  const float Ur = Xr + (empty ? 0.f : (float)(rec.j1 - rec.j0 + 1)) + kDiceEps;
  const float Uc = Xc + (empty ? 0.f : (float)(rec.i1 - rec.i0 + 1)) + kDiceEps;
  // d dice / d s = -2 t / U + 4 I s / U^2 ; through the sigmoid: * s (1 - s); mean over N
  if (row_i < H) {
    ws.coef_row[n * H + row_i] = inv_n * (-2.f * (tr ? 1.f : 0.f) / Ur + 4.f * Ir * sr / (Ur * Ur)) * sr * (1.f - sr);
    ws.arg_row[n * H + row_i] = (int)(0xffffffffu - (unsigned)(rp & 0xffffffffull));
  }
  if (col_i < W) {
    ws.coef_col[n * W + col_i] = inv_n * (-2.f * (tc ? 1.f : 0.f) / Uc + 4.f * Ic * sc / (Uc * Uc)) * sc * (1.f - sc);
    ws.arg_col[n * W + col_i] = (int)(0xffffffffu - (unsigned)(cp & 0xffffffffull));
  }
  if (wid == 0) {
    // fixed-order sum of the strip numerators (sequential over lanes -> same order as a serial loop)
    float tot = 0.f;
    for (int s = 0; s < S; ++s) tot += __shfl_sync(kFull, num, s);
    if (lane == 0) {
      ws.inst_prj[n] = (1.f - 2.f * Ir / Ur) + (1.f - 2.f * Ic / Uc);
      ws.inst_num[n] = tot;
      reinterpret_cast<int4*>(ws.span)[n] = make_int4(sp.y_lo, sp.y_hi, sp.c_lo, sp.c_hi);
      __threadfence();
      s_last = atomicAdd(&sched->ticket, 1u) == (unsigned)(N - 1);
    }
  }
  __syncthreads();
  if (!s_last || tid >= 32) return;
  __threadfence();
  float prj = 0.f, pn = 0.f;
  for (int i = lane; i < N; i += 32) { prj += __ldcg(ws.inst_prj + i); pn += __ldcg(ws.inst_num + i); }
  prj = warp_sum(prj);
  pn = warp_sum(pn);
  if (lane == 0) {
    const float scale = warm / fmaxf((float)wtot, 1.f);
    losses_out[0] = prj * inv_n;
    losses_out[1] = pn * scale;
    losses_out[2] = pn;
    losses_out[3] = (float)wtot;
    ws.scale[0] = scale;
    sched->ticket = 0u;
  }
}

// ---------------------------------------------------------------------------------------
// backward: one CTA per instance turns the raw pairwise gradient into the final gradient, in place:
//   span pixels *= g_pair * scale;  arg-max positions += g_prj * coefficient.  A position that is both a row and
//   a column arg-max is written once (by its column), so no two threads touch the same address in a phase.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OP_FIN_NT)
onepass_backward_kernel(int H, int W, OpWorkspace ws, const float* __restrict__ g_prj_p,
                        const float* __restrict__ g_pair_p, float* __restrict__ g_logits) {
  __shared__ int s_ar[512], s_ac[512];
  __shared__ float s_cr[512];
  const int n = blockIdx.x, tid = threadIdx.x;
  float* ginst = g_logits + (int64_t)n * H * W;
  const float gp = g_prj_p[0], gq = g_pair_p[0] * ws.scale[0];
  const int4 sp = reinterpret_cast<const int4*>(ws.span)[n];
  int ar = 0, ac = 0;
  float cr = 0.f, cc = 0.f;
  if (tid < H) { ar = ws.arg_row[n * H + tid]; cr = ws.coef_row[n * H + tid] * gp; s_ar[tid] = ar; s_cr[tid] = cr; }
  if (tid < W) { ac = ws.arg_col[n * W + tid]; cc = ws.coef_col[n * W + tid] * gp; s_ac[tid] = ac; }
  const int sw4 = (sp.w - sp.z + 1) >> 2, srows = sp.y - sp.x + 1;
  if (srows > 0) {
    for (int i = tid; i < srows * sw4; i += OP_FIN_NT) {
      const int ry = op_div(i, sw4), c4 = i - ry * sw4;
      float4* p = reinterpret_cast<float4*>(ginst + (sp.x + ry) * W + sp.z + 4 * c4);
      float4 q = *p;
      q.x *= gq; q.y *= gq; q.z *= gq; q.w *= gq;
      *p = q;
    }
  }
  __syncthreads();
  if (tid < H && s_ac[ar] != tid) ginst[tid * W + ar] += cr;
  if (tid < W) {
    float* p = ginst + ac * W + tid;
    *p = s_ar[ac] == tid ? (*p + s_cr[ac]) + cc : *p + cc;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
inline bool op_supported(int64_t N, int64_t H, int64_t W, int d) {
  return N > 0 && N <= OP_MAX_N && H > 0 && H <= 512 && W > 0 && W <= 512 && (W % 4 == 0) && d >= 1 && d <= 4;
}

template <int D>
inline size_t op_smem_bytes(int64_t N, int64_t W) {
  constexpr int ROWS = OpTile<D>::ROWS, TW = OpTile<D>::TW;
  return (size_t)2 * ROWS * W * 4 + (size_t)ROWS * TW * 8 + (size_t)ROWS * TW + (size_t)N * sizeof(SRec);
}

template <int NCHUNK, int D, bool FULLW>
int op_launch_main(cudaStream_t st, const float* logits, const uint8_t* edge_bits, const int32_t* rects,
                   const int32_t* inst_gt, const int32_t* gt_img, int N, int H, int W, OpWorkspace ws, OpSched* sched,
                   float* g_logits) {
  const size_t smem = op_smem_bytes<D>(N, W);
  auto kern = onepass_main_kernel<NCHUNK, D, FULLW>;
  static thread_local size_t configured = 0;      // per instantiation
  static thread_local int occ_dev = -1, occ = 0;
  static thread_local size_t occ_smem = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_error(cudaGetLastError());
      return BXS_ERR_UNSUPPORTED;
    }
    configured = smem;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != occ_dev || smem != occ_smem) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OP_NT, smem) != cudaSuccess || occ < 1) {
      set_last_error(cudaGetLastError());
      return BXS_ERR_UNSUPPORTED;
    }
    occ_dev = dev;
    occ_smem = smem;
  }
  const int S = (int)ceil_div(H, OP_R);
  const int64_t total = (int64_t)N * S;
  const int grid = (int)std::min<int64_t>(total, (int64_t)sm_count() * occ);
  kern<<<grid, OP_NT, smem, st>>>(logits, edge_bits, rects, inst_gt, gt_img, N, H, W, S, ws, sched, g_logits);
  return check_launch();
}

}  // namespace
}  // namespace bxs

using namespace bxs;

extern "C" int bxs_boxinst_loss_fused_supported(int64_t N, int64_t H, int64_t W, int dilation) {
  return op_supported(N, H, W, dilation) ? 1 : 0;
}

extern "C" int64_t bxs_boxinst_loss_fused_workspace_bytes(int64_t N, int64_t H, int64_t W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)op_carve(nullptr, N, H, W).total_bytes;
}

extern "C" int64_t bxs_boxinst_loss_fused_sched_bytes(void) { return (int64_t)sizeof(OpSched); }

extern "C" int bxs_boxinst_loss_fused_forward(const float* logits, const uint8_t* edge_bits, const int32_t* rects,
                                              const int32_t* inst_gt, const int32_t* gt_img, const float* iter_ptr,
                                              float warmup_iters, void* workspace, void* sched_state,
                                              float* losses_out, float* g_logits, int64_t N, int64_t H, int64_t W,
                                              int dilation, bxs_stream_t stream) {
  if (!logits || !edge_bits || !rects || !inst_gt || !gt_img || !iter_ptr || !workspace || !sched_state ||
      !losses_out || !g_logits || N <= 0 || H <= 0 || W <= 0 || !(warmup_iters > 0.f))
    return BXS_ERR_INVALID_ARG;
  if (!op_supported(N, H, W, dilation) || (reinterpret_cast<uintptr_t>(logits) & 15) ||
      (reinterpret_cast<uintptr_t>(g_logits) & 15))
    return BXS_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  OpWorkspace ws = op_carve(workspace, N, H, W);
  OpSched* sched = reinterpret_cast<OpSched*>(sched_state);
  const int S = (int)ceil_div(H, OP_R);
  int rc = BXS_ERR_UNSUPPORTED;
#define BXS_OP_CASE(NC, DD)                                                                                          \
  rc = (W == NC * 128) ? op_launch_main<NC, DD, true>(st, logits, edge_bits, rects, inst_gt, gt_img, (int)N, (int)H,  \
                                                      (int)W, ws, sched, g_logits)                                  \
                       : op_launch_main<NC, DD, false>(st, logits, edge_bits, rects, inst_gt, gt_img, (int)N, (int)H, \
                                                       (int)W, ws, sched, g_logits);                                \
  if (rc == BXS_OK) {                                                                                               \
    onepass_finalize_kernel<DD><<<(unsigned)N, OP_FIN_NT, 0, st>>>(rects, inst_gt, gt_img, (int)N, (int)H, (int)W, S, \
                                                                   ws, sched, iter_ptr, warmup_iters, losses_out);  \
    rc = check_launch();                                                                                            \
  }
#define BXS_OP_D(NC)                                                  \
  switch (dilation) {                                                 \
    case 1: { BXS_OP_CASE(NC, 1) } break;                             \
    case 2: { BXS_OP_CASE(NC, 2) } break;                             \
    case 3: { BXS_OP_CASE(NC, 3) } break;                             \
    default: { BXS_OP_CASE(NC, 4) } break;                            \
  }
  if (W <= 128) { BXS_OP_D(1) }
  else if (W <= 256) { BXS_OP_D(2) }
  else { BXS_OP_D(4) }
#undef BXS_OP_D
#undef BXS_OP_CASE
  return rc;
}

extern "C" int bxs_boxinst_loss_fused_backward(const void* workspace, const float* g_prj, const float* g_pair,
                                               float* g_logits, int64_t N, int64_t H, int64_t W,
                                               bxs_stream_t stream) {
  if (!workspace || !g_prj || !g_pair || !g_logits || N <= 0 || H <= 0 || W <= 0) return BXS_ERR_INVALID_ARG;
  OpWorkspace ws = op_carve(const_cast<void*>(workspace), N, H, W);
  onepass_backward_kernel<<<(unsigned)N, OP_FIN_NT, 0, as_stream(stream)>>>((int)H, (int)W, ws, g_prj, g_pair, g_logits);
  return check_launch();
}

// a2 / a3 / a4: dense dynamic 1x1 convolution  out[b,i,p] = sum_c kernel[b,i,c] * feat[b,c,p]
// (K = C = 256 channels) as a TMA-fed tcgen05 tensor-core kernel, TF32 inputs, FP32 accumulation in TMEM.
// Replaces the F.conv2d / einsum calls at
//   mmdet/models/dense_heads/box_solov2_head.py:209-211   (all S^2 cells, groups=B)
//   mmdet/models/dense_heads/discobox_head.py:1219        (positive kernels of one image)
//   mmdet/models/dense_heads/box2mask_head.py:345         (einsum 'bqc,bchw->bqhw')
// which run as cuDNN/cuBLAS GEMMs (cuDNN convolutions use TF32 by default on Ampere and later).
//
// GEMM view per image:  D[M = pixels, N = kernels] = A[M, K] * B[N, K]^T
//   A = feat^T : the pixel index is contiguous in memory -> "MN-major" operand, staged by TMA as
//       [k-row][32 pixels] boxes with the 128-byte swizzle on 32-BYTE atoms (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
//       <-> UMMA SWIZZLE_128B_BASE32B: the layout the tensor core requires for an MN-major 32-bit operand;
//       the ordinary 128B swizzle silently yields zeros), LBO = 4 KB between 32-pixel groups, SBO = 512 B between
//       groups of 4 k-rows;
//   B = kernel : K-major, [row = kernel][32 channels] boxes with the 128-byte swizzle (SBO = 1 KB);
//   D          : 128 TMEM lanes (pixels) x N <= 128 fp32 columns, double buffered (256 columns).
// The op is HBM bound (each feat tile is read once: 128 KB per 128 pixels; B stays resident in shared
// memory for the whole CTA lifetime), so one CTA per SM streams pixel tiles:
//   warp 0   : TMA producer (B once, then a 4-stage ring of 16 KB A stages)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (4 MMAs of K = 8 per stage)
//   warps 2-9: epilogue -- tcgen05.ld 32x32b (one pixel per thread; the two warps of a lane quarter split the
//              columns) and coalesced stores
// synchronised with mbarriers only (full/empty per stage, tmem_full/tmem_empty per accumulator).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace bxs {
namespace {

constexpr int BLOCK_M = 128;        // pixels per tile (TMEM lanes)
constexpr int BLOCK_N = 128;        // kernels per CTA (TMEM columns per accumulator)
constexpr int BLOCK_K = 32;         // channels per stage = one 128-byte swizzle row of tf32
constexpr int UMMA_K = 8;           // tf32: 32 bytes per instruction
constexpr int MAX_KBLOCKS = 8;      // C <= 256
constexpr int STAGES = 6;            // 96 KB of feat in flight per SM (4 stages measured 0.54 of HBM peak: latency bound)
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 4;      // 16 KB
constexpr int B_KBLOCK_BYTES = BLOCK_N * BLOCK_K * 4;     // 16 KB
constexpr int NUM_THREADS = 320;     // forward: TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int WG_THREADS = 192;      // d/d kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int TMEM_COLS = 2 * BLOCK_N;                    // 256 (power of two)

struct SmemLayout {
  alignas(1024) uint8_t a[STAGES][A_STAGE_BYTES];
  alignas(1024) uint8_t b[MAX_KBLOCKS][B_KBLOCK_BYTES];
  alignas(8) uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t b_full;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

// shared-memory matrix descriptor (SM100): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | swizzle-128B (2) << 61
// layout_type: 2 = SWIZZLE_128B (16-byte atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms; the only swizzled layout the
// tensor core accepts for an MN-major 32-bit operand)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// instruction descriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, a_major MN (1) @15, b_major K (0) @16,
// N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// grid: (ctas_per_group, B * n_chunks).  CTA (x, g): image b = g / n_chunks, kernel chunk g % n_chunks,
// pixel tiles x, x + gridDim.x, ...
__global__ void __launch_bounds__(NUM_THREADS, 1)
dynconv_tf32_kernel(const __grid_constant__ CUtensorMap tm_feat, const __grid_constant__ CUtensorMap tm_kern,
                    float* __restrict__ out, int C, int P, int I, int n_chunks) {
  extern __shared__ uint8_t smem_raw[];
  SmemLayout& S = *reinterpret_cast<SmemLayout*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y / n_chunks, chunk = blockIdx.y % n_chunks;
  const int kblocks = (C + BLOCK_K - 1) / BLOCK_K;       // channels beyond C are zero-filled by the 3-D tensor map
  const int tiles_m = (P + BLOCK_M - 1) / BLOCK_M;
  const int n_here = min(BLOCK_N, ((I - chunk * BLOCK_N) + 15) / 16 * 16);     // multiple of 16, <= 128

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.b_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&S.tmem_full[a], 1); mbar_init(&S.tmem_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                      // TMEM allocation by one full warp; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(&S.b_full, (uint32_t)kblocks * B_KBLOCK_BYTES);
      for (int kb = 0; kb < kblocks; ++kb)
        tma_load_2d(&tm_kern, &S.b_full, S.b[kb], kb * BLOCK_K, b * I + chunk * BLOCK_N);
      int stage = 0, phase = 0;
      for (int mt = blockIdx.x; mt < tiles_m; mt += gridDim.x) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&S.empty[stage], phase ^ 1);
          mbar_expect_tx(&S.full[stage], A_STAGE_BYTES);
#pragma unroll
          for (int mg = 0; mg < BLOCK_M / 32; ++mg)
            tma_load_3d(&tm_feat, &S.full[stage], S.a[stage] + mg * (BLOCK_K * 128), mt * BLOCK_M + mg * 32,
                        kb * BLOCK_K, b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(n_here);
      mbar_wait(&S.b_full, 0);
      tc_fence_after();
      int stage = 0, phase = 0, t = 0;
      for (int mt = blockIdx.x; mt < tiles_m; mt += gridDim.x, ++t) {
        const int acc = t & 1;
        mbar_wait(&S.tmem_empty[acc], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&S.full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(S.a[stage]);
          const uint32_t b_base = smem_u32(S.b[kb]);
#pragma unroll
          for (int k4 = 0; k4 < BLOCK_K / UMMA_K; ++k4) {
            // A (MN-major, 128B swizzle with 32-byte atoms): 8 k-rows = two 512-byte atoms (SBO) per 32-pixel group;
            // the 32-pixel groups (one TMA box each) are 4 KB apart (LBO)
            const uint64_t adesc = make_desc(a_base + k4 * 1024, BLOCK_K * 128, 512, 1);
            // B (K-major, 128B swizzle): advance 32 bytes inside the 128-byte swizzled row; 8-row groups 1 KB apart
            const uint64_t bdesc = make_desc(b_base + k4 * (UMMA_K * 4), 16, 1024, 2);
            umma_tf32(tmem_d, adesc, bdesc, idesc, (kb | k4) != 0 ? 1u : 0u);
          }
          umma_commit(&S.empty[stage]);            // frees this A stage once the MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&S.tmem_full[acc]);            // accumulator complete
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int quarter = warp & 3;                  // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int half = (warp - 2) >> 2;              // warps 2-5: columns [0, 64), warps 6-9: columns [64, 128)
    const int c_lo = half * (BLOCK_N / 2), c_hi = min(n_here, c_lo + BLOCK_N / 2);
    int t = 0;
    for (int mt = blockIdx.x; mt < tiles_m; mt += gridDim.x, ++t) {
      const int acc = t & 1;
      mbar_wait(&S.tmem_full[acc], (t >> 1) & 1);
      tc_fence_after();
      const int pixel = mt * BLOCK_M + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      float* dst = out + ((int64_t)b * I + chunk * BLOCK_N) * P + pixel;
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t r[32];
        const bool two = c0 + 16 < c_hi;
        tmem_ld16_async(taddr + c0, r);
        if (two) tmem_ld16_async(taddr + c0 + 16, r + 16);
        tmem_ld_wait();
        if (pixel < P) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if ((j < 16 || two) && chunk * BLOCK_N + c0 + j < I) dst[(int64_t)(c0 + j) * P] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows, cols] float32 matrix, box = [box_rows, 32 floats], 128-byte swizzle, zero fill out of bounds
bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [batch, rows, cols] float32 tensor, box = [1, box_rows, 32 floats]: rows beyond `rows` of a batch entry are OUT OF
// BOUNDS and read as zeros (a 2-D [batch * rows, cols] view would run into the next entry instead)
bool make_map3(CUtensorMap* map, const float* base, int64_t batch, int64_t rows, int64_t cols, int box_rows,
               CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * (cuuint64_t)rows * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// kT[b, c, i] = kernels[b, i, c] for i < I, 0 for I <= i < Ipad   (the K-major B operand of d/d feat = K^T . G)
__global__ void transpose_pad_kernel(const float* __restrict__ k, float* __restrict__ kt, int I, int C, int Ipad, int64_t total) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % Ipad);
    const int c = (int)((idx / Ipad) % C);
    const int64_t b = idx / ((int64_t)Ipad * C);
    kt[idx] = i < I ? __ldg(k + (b * I + i) * C + c) : 0.f;
  }
}

// =====================================================================================================================
// d/d kernel:  dK[b,i,c] = sum_p g[b,i,p] * feat[b,c,p]  -- a GEMM with a LONG reduction (K = pixels, 51 200 .. 65 536) and a
// tiny output (I x C).  Both operands are K-major (the pixel index is contiguous in g and in feat), i.e. the plain
// 128-byte-swizzle layout of the forward kernel's B operand on both sides.  Split-K: CTA (split, b * chunks + chunk) reduces
// pixels [p0, p1) of image b for 128 rows of g into ONE 128 x C accumulator in TMEM (C <= 256 columns) and stores it as a
// partial; dynconv_wgrad_reduce sums the partials in split order (deterministic).
//   warp 0: TMA producer (A = g tile 128 x 32, B = feat tile C x 32 per stage, 4 stages)
//   warp 1: TMEM allocator + single-thread tcgen05.mma issuer (4 MMAs of K = 8 per stage)
//   warps 2-5: epilogue (tcgen05.ld, one g row per thread)
// =====================================================================================================================
constexpr int WG_STAGES = 4;
constexpr int WG_A_BYTES = 128 * BLOCK_K * 4;      // 16 KB
constexpr int WG_B_BYTES = 256 * BLOCK_K * 4;      // 32 KB

struct WgSmem {
  alignas(1024) uint8_t a[WG_STAGES][WG_A_BYTES];
  alignas(1024) uint8_t b[WG_STAGES][WG_B_BYTES];
  alignas(8) uint64_t full[WG_STAGES];
  uint64_t empty[WG_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// instruction descriptor, both operands K-major: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t make_idesc_kk(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
dynconv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_feat,
                     float* __restrict__ partial, int C, int P, int I, int chunks, int splits, int kb_per_split) {
  extern __shared__ uint8_t smem_raw[];
  WgSmem& S = *reinterpret_cast<WgSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, b = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
  const int kb_total = (P + BLOCK_K - 1) / BLOCK_K;
  const int kb0 = split * kb_per_split, kb1 = min(kb0 + kb_per_split, kb_total);
  const int n_cols = (C + 15) / 16 * 16;                               // UMMA N (multiple of 16, <= 256)
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&S.empty[stage], phase ^ 1);
        mbar_expect_tx(&S.full[stage], WG_A_BYTES + (uint32_t)C * BLOCK_K * 4);
        tma_load_3d(&tm_g, &S.full[stage], S.a[stage], kb * BLOCK_K, chunk * 128, b);       // rows >= I: zeros
        tma_load_3d(&tm_feat, &S.full[stage], S.b[stage], kb * BLOCK_K, 0, b);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_kk(n_cols);
      int stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&S.full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(S.a[stage]), b_base = smem_u32(S.b[stage]);
#pragma unroll
        for (int k4 = 0; k4 < BLOCK_K / UMMA_K; ++k4) {
          // K-major, 128-byte swizzle: advance 32 bytes inside the swizzled row; 8-row groups 1 KB apart
          const uint64_t adesc = make_desc(a_base + k4 * (UMMA_K * 4), 16, 1024, 2);
          const uint64_t bdesc = make_desc(b_base + k4 * (UMMA_K * 4), 16, 1024, 2);
          umma_tf32(tmem_base, adesc, bdesc, idesc, (kb > kb0 || k4 != 0) ? 1u : 0u);
        }
        umma_commit(&S.empty[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&S.acc_full);
    }
  } else {
    const int quarter = warp & 3;
    mbar_wait(&S.acc_full, 0);
    tc_fence_after();
    const int row = chunk * 128 + quarter * 32 + lane;             // g row (kernel index i)
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* dst = partial + (((int64_t)b * splits + split) * I + row) * C;
    for (int c0 = 0; c0 < n_cols; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + c0, r);
      if (row < I && kb1 > kb0) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (c0 + j < C) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                                __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

__global__ void dynconv_wgrad_reduce(const float* __restrict__ partial, float* __restrict__ out, int splits, int64_t per_image,
                                     int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_image, r = i - b * per_image;
    const float* p = partial + b * splits * per_image + r;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += p[(int64_t)s * per_image];          // split order: deterministic
    out[i] = acc;
  }
}

}  // namespace
}  // namespace bxs

using namespace bxs;

// feat [B,C,P], kernels [B,I,C] -> out [B,I,P] (all float32, contiguous).  Requirements of the TMA/UMMA path:
// C % 32 == 0, C <= 256, P % 4 == 0, 16-byte aligned bases.  BXS_ERR_UNSUPPORTED otherwise.
extern "C" int bxs_dynconv1x1_forward(const float* feat, const float* kernels, float* out, int64_t B, int64_t C, int64_t P,
                                      int64_t I, bxs_stream_t stream) {
  if (!feat || !kernels || !out || B <= 0 || C <= 0 || P <= 0 || I <= 0) return BXS_ERR_INVALID_ARG;
  if (C % BLOCK_K || C > BLOCK_K * MAX_KBLOCKS || P % 4 || (reinterpret_cast<uintptr_t>(feat) & 15) ||
      (reinterpret_cast<uintptr_t>(kernels) & 15) || B * C >= (int64_t(1) << 31) || B * I >= (int64_t(1) << 31))
    return BXS_ERR_UNSUPPORTED;
  CUtensorMap tm_feat, tm_kern;
  if (!make_map3(&tm_feat, feat, B, C, P, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map(&tm_kern, kernels, B * I, C, BLOCK_N, CU_TENSOR_MAP_SWIZZLE_128B)) {
    set_last_error(cudaErrorNotSupported);
    return BXS_ERR_LAUNCH;
  }
  const int n_chunks = (int)ceil_div(I, BLOCK_N);
  const int64_t groups = B * n_chunks;
  if (groups > 65535) return BXS_ERR_UNSUPPORTED;
  const int tiles_m = (int)ceil_div(P, BLOCK_M);
  int per_group = (int)std::max<int64_t>(1, sm_count() / groups);
  per_group = std::min(per_group, tiles_m);
  const size_t smem = sizeof(SmemLayout) + 1024;
  cudaFuncSetAttribute(dynconv_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dynconv_tf32_kernel<<<dim3(per_group, (unsigned)groups), NUM_THREADS, smem, as_stream(stream)>>>(
      tm_feat, tm_kern, out, (int)C, (int)P, (int)I, n_chunks);
  return check_launch();
}


// ---------------------------------------------------------------------------------------
// backward of the dynamic 1x1 convolution (the autograd of box_solov2_head.py:209-211, discobox_head.py:1219,
// box2mask_head.py:345; the reference gets it from cuDNN / cuBLAS)
//   d/d feat  [B,C,P] = K^T . G : the FORWARD kernel with the roles swapped -- "feat" := g_out [B,I,P] (its rows beyond I
//       read as zeros through the 3-D tensor map), "kernels" := K^T zero-padded to a multiple of 32 instances
//       (transpose_pad_kernel, in the workspace).  I <= 256.
//   d/d kernel [B,I,C] = G . F^T : dynconv_wgrad_kernel (split-K) + dynconv_wgrad_reduce.
// workspace: bxs_dynconv1x1_backward_workspace_bytes(B, C, P, I) bytes.
// ---------------------------------------------------------------------------------------
namespace {
inline int wg_splits(int64_t B, int64_t P, int64_t I) {
  const int64_t chunks = ceil_div(I, 128), kb_total = ceil_div(P, BLOCK_K);
  int64_t s = std::max<int64_t>(1, (2 * (int64_t)sm_count()) / (B * chunks));
  s = std::min<int64_t>(s, std::max<int64_t>(1, kb_total / 8));          // at least 8 k-blocks (256 pixels) per CTA
  return (int)s;
}
}  // namespace

extern "C" int64_t bxs_dynconv1x1_backward_workspace_bytes(int64_t B, int64_t C, int64_t P, int64_t I) {
  if (B <= 0 || C <= 0 || P <= 0 || I <= 0) return 0;
  const int64_t ipad = ceil_div(I, 32) * 32;
  return 4 * B * C * ipad + 1024 + 4 * B * (int64_t)wg_splits(B, P, I) * I * C + 1024;
}

extern "C" int bxs_dynconv1x1_backward(const float* feat, const float* kernels, const float* g_out, float* g_feat,
                                       float* g_kernels, void* workspace, int64_t B, int64_t C, int64_t P, int64_t I,
                                       bxs_stream_t stream) {
  if (!feat || !kernels || !g_out || (!g_feat && !g_kernels) || !workspace || B <= 0 || C <= 0 || P <= 0 || I <= 0)
    return BXS_ERR_INVALID_ARG;
  if (C % BLOCK_K || C > BLOCK_K * MAX_KBLOCKS || P % 4 || (g_feat && I > 256) || (reinterpret_cast<uintptr_t>(feat) & 15) ||
      (reinterpret_cast<uintptr_t>(g_out) & 15) || (reinterpret_cast<uintptr_t>(kernels) & 15))
    return BXS_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int64_t ipad = ceil_div(I, 32) * 32;
  float* kt = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(kt + B * C * ipad) + 1023) & ~uintptr_t(1023));
  if (g_feat) {
    const int64_t total = B * C * ipad;
    transpose_pad_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 1024), 256, 0, st>>>(kernels, kt, (int)I, (int)C,
                                                                                                  (int)ipad, total);
    CUtensorMap tm_g, tm_kt;
    if (!make_map3(&tm_g, g_out, B, I, P, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map(&tm_kt, kt, B * C, ipad, BLOCK_N, CU_TENSOR_MAP_SWIZZLE_128B)) {
      set_last_error(cudaErrorNotSupported);
      return BXS_ERR_LAUNCH;
    }
    const int n_chunks = (int)ceil_div(C, BLOCK_N);
    const int64_t groups = B * n_chunks;
    if (groups > 65535) return BXS_ERR_UNSUPPORTED;
    const int tiles_m = (int)ceil_div(P, BLOCK_M);
    const int per_group = std::min((int)std::max<int64_t>(1, sm_count() / groups), tiles_m);
    const size_t smem = sizeof(SmemLayout) + 1024;
    cudaFuncSetAttribute(dynconv_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // out[b, c, p] = sum_i kT[b, c, i] * g[b, i, p]: "C" := I (zero-filled up to ipad), "I" := C
    dynconv_tf32_kernel<<<dim3(per_group, (unsigned)groups), NUM_THREADS, smem, st>>>(tm_g, tm_kt, g_feat, (int)ipad, (int)P,
                                                                                      (int)C, n_chunks);
    int rc = check_launch();
    if (rc != BXS_OK) return rc;
  }
  if (g_kernels) {
    const int chunks = (int)ceil_div(I, 128), splits = wg_splits(B, P, I);
    const int kb_total = (int)ceil_div(P, BLOCK_K), kb_per = (int)ceil_div(kb_total, splits);
    CUtensorMap tm_g, tm_f;
    if (!make_map3(&tm_g, g_out, B, I, P, 128, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map3(&tm_f, feat, B, C, P, (int)C, CU_TENSOR_MAP_SWIZZLE_128B)) {
      set_last_error(cudaErrorNotSupported);
      return BXS_ERR_LAUNCH;
    }
    const size_t smem = sizeof(WgSmem) + 1024;
    cudaFuncSetAttribute(dynconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dynconv_wgrad_kernel<<<dim3((unsigned)splits, (unsigned)(B * chunks)), WG_THREADS, smem, st>>>(
        tm_g, tm_f, partial, (int)C, (int)P, (int)I, chunks, splits, kb_per);
    int rc = check_launch();
    if (rc != BXS_OK) return rc;
    const int64_t total = B * I * C;
    dynconv_wgrad_reduce<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 1024), 256, 0, st>>>(partial, g_kernels, splits,
                                                                                                  I * C, total);
    rc = check_launch();
    if (rc != BXS_OK) return rc;
  }
  return BXS_OK;
}

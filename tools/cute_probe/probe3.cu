// Diagnostic matrix: which tcgen05.mma variant produces non-zero results from hand-filled shared memory?
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;

__device__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// VAR 0: tf32, A MN-major, no mask   1: tf32, A MN-major, mask form   2: tf32, A K-major, mask form
// VAR 3: bf16 K-major both, mask form   4: tf32 A K-major, no mask, extra fences
template <int VAR>
__global__ void probe(float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  using T = std::conditional_t<VAR == 3, cutlass::bfloat16_t, cutlass::tfloat32_t>;
  constexpr int KT = VAR == 3 ? 64 : 32;            // tile K (one 128-byte swizzle row)
  constexpr int KI = VAR == 3 ? 16 : 8;             // K per instruction
  constexpr bool A_MN = VAR <= 1 || VAR == 5;
  T* pa = reinterpret_cast<T*>(smem);
  T* pb = reinterpret_cast<T*>(smem + 32768);
  auto la = [] {
    if constexpr (VAR == 5) return tile_to_shape(UMMA::Layout_MN_SW128_32B_Atom<T>{}, Shape<_128, Int<KT>>{}, Step<_2, _1>{});
    else if constexpr (A_MN) return tile_to_shape(UMMA::Layout_MN_SW128_Atom<T>{}, Shape<_128, Int<KT>>{});
    else return tile_to_shape(UMMA::Layout_K_SW128_Atom<T>{}, Shape<_128, Int<KT>>{});
  }();
  auto lb = tile_to_shape(UMMA::Layout_K_SW128_Atom<T>{}, Shape<_128, Int<KT>>{});
  auto ta = make_tensor(make_smem_ptr(pa), la);
  auto tb = make_tensor(make_smem_ptr(pb), lb);
  for (int i = threadIdx.x; i < 128 * KT; i += blockDim.x) {
    int m = i % 128, k = i / 128;
    ta(m, k) = T(float((m * 3 + k * 5) % 17) / 16.f);
    tb(m, k) = T(m < 16 ? float((m * 7 + k) % 13) / 8.f : 0.f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&tbase)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) {
    constexpr auto AM = A_MN ? UMMA::Major::MN : UMMA::Major::K;
    auto idesc = UMMA::make_runtime_instr_desc<T, T, float, 128, 16, AM, UMMA::Major::K>();
    if (cute::elect_one_sync()) {
      for (int k = 0; k < KT / KI; ++k) {
        auto da = UMMA::make_umma_desc<AM>(local_tile(ta, Shape<_128, Int<KI>>{}, make_coord(0, k)));
        auto db = UMMA::make_umma_desc<UMMA::Major::K>(local_tile(tb, Shape<_128, Int<KI>>{}, make_coord(0, k)));
        uint32_t acc = k > 0, z = 0;
        if (VAR == 5) printf("v5 k%d adesc %016llx lbo %u sbo %u layout %u\n", k, (unsigned long long)uint64_t(da), (unsigned)da.leading_byte_offset_, (unsigned)da.stride_byte_offset_, (unsigned)da.layout_type_);
        if constexpr (VAR == 0 || VAR == 4) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tbase), "l"(uint64_t(da)), "l"(uint64_t(db)), "r"(uint32_t(idesc >> 32)), "r"(acc) : "memory");
        } else if constexpr (VAR == 3) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                       ::"r"(tbase), "l"(uint64_t(da)), "l"(uint64_t(db)), "r"(uint32_t(idesc >> 32)), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                       ::"r"(tbase), "l"(uint64_t(da)), "l"(uint64_t(db)), "r"(uint32_t(idesc >> 32)), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(&bar)) : "memory");
    }
  }
  asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
               ::"r"(s2u(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  uint32_t taddr = tbase + ((uint32_t)((threadIdx.x / 32) * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(32));
}

template <int VAR>
void run() {
  constexpr int KT = VAR == 3 ? 64 : 32;
  float* out;
  cudaMallocManaged(&out, 128 * 16 * 4);
  for (int i = 0; i < 128 * 16; ++i) out[i] = -1.f;
  cudaFuncSetAttribute(probe<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<VAR><<<1, 128, 64 * 1024>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
    double ref = 0;
    for (int k = 0; k < KT; ++k) ref += double((m * 3 + k * 5) % 17) / 16.0 * double((n * 7 + k) % 13) / 8.0;
    maxerr = fmax(maxerr, fabs(ref - out[m * 16 + n])); maxref = fmax(maxref, fabs(ref));
  }
  printf("variant %d: %s  max|err| %.4g (max ref %.4g) out[0..3] %.4f %.4f %.4f %.4f\n", VAR, cudaGetErrorString(e), maxerr, maxref,
         out[0], out[1], out[2], out[3]);
}
int main() { run<5>(); return 0; }

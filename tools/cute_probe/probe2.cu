// Diagnostic: one tcgen05 TF32 MMA (M128 N16 K8 x4) on hand-filled shared memory in CuTe's canonical layouts,
// A MN-major / B K-major, using CuTe's own descriptor builders -- checks the result against a CPU product.
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
#include <cute/arch/mma_sm100_umma.hpp>
using namespace cute;
using T = cutlass::tfloat32_t;

__device__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ORDER>
__global__ void probe(float* out, unsigned long long* descs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  T* pa = reinterpret_cast<T*>(smem);
  T* pb = reinterpret_cast<T*>(smem + 32768);
  auto la = [] {
    if constexpr (ORDER == 0) return tile_to_shape(UMMA::Layout_MN_SW128_Atom<T>{}, Shape<_128, _32>{});
    else return tile_to_shape(UMMA::Layout_MN_SW128_Atom<T>{}, Shape<_128, _32>{}, Step<_2, _1>{});
  }();
  auto lb = tile_to_shape(UMMA::Layout_K_SW128_Atom<T>{}, Shape<_128, _32>{});
  auto ta = make_tensor(make_smem_ptr(pa), la);
  auto tb = make_tensor(make_smem_ptr(pb), lb);
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
    int m = i % 128, k = i / 128;
    ta(m, k) = T(float((m * 3 + k * 5) % 17) / 16.f);
    tb(m, k) = T(m < 16 ? float((m * 7 + k) % 13) / 8.f : 0.f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> async proxy (UMMA)
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&tbase)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    auto idesc = UMMA::make_runtime_instr_desc<T, T, float, 128, 16, UMMA::Major::MN, UMMA::Major::K>();
    for (int k = 0; k < 4; ++k) {
      auto da = UMMA::make_umma_desc<UMMA::Major::MN>(local_tile(ta, Shape<_128, _8>{}, make_coord(0, k)));
      auto db = UMMA::make_umma_desc<UMMA::Major::K>(local_tile(tb, Shape<_128, _8>{}, make_coord(0, k)));
      descs[2 * k] = uint64_t(da); descs[2 * k + 1] = uint64_t(db);
      uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tbase), "l"(uint64_t(da)), "l"(uint64_t(db)), "r"(uint32_t(idesc >> 32)), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(&bar)) : "memory");
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

template <int ORDER>
void run() {
  float* out; unsigned long long* descs;
  cudaMallocManaged(&out, 128 * 16 * 4); cudaMallocManaged(&descs, 64);
  cudaFuncSetAttribute(probe<ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<ORDER><<<1, 128, 64 * 1024>>>(out, descs);
  printf("order %d sync: %s\n", ORDER, cudaGetErrorString(cudaDeviceSynchronize()));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
    double ref = 0;
    for (int k = 0; k < 32; ++k) ref += double((m * 3 + k * 5) % 17) / 16.0 * double((n * 7 + k) % 13) / 8.0;
    maxerr = fmax(maxerr, fabs(ref - out[m * 16 + n])); maxref = fmax(maxref, fabs(ref));
  }
  printf("  max |err| %.4g (max ref %.4g)  out[0..3] %.4f %.4f %.4f %.4f\n", maxerr, maxref, out[0], out[1], out[2], out[3]);
  for (int k = 0; k < 4; ++k) printf("  k%d: adesc %016llx bdesc %016llx\n", k, descs[2 * k], descs[2 * k + 1]);
}
int main() { run<0>(); run<1>(); return 0; }

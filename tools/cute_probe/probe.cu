// Prints the UMMA shared-memory descriptors CuTe builds for the two operand layouts of dynconv_tcgen05.cu
// (diagnostic only; uses the CUTLASS headers vendored in the image, not part of the product).
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;

__global__ void probe() {
  extern __shared__ __align__(1024) uint8_t smem[];
  using T = cutlass::tfloat32_t;
  {
    // A: MN-major, 128 (M) x 32 (K) tile built from the SW128 MN atom
    auto layout = tile_to_shape(UMMA::Layout_MN_SW128_Atom<T>{}, Shape<_128, _32>{});
    auto t = make_tensor(make_smem_ptr(reinterpret_cast<T*>(smem)), layout);
    auto tk = local_tile(t, Shape<_128, _8>{}, make_coord(0, 0));
    auto d0 = UMMA::make_umma_desc<UMMA::Major::MN>(tk);
    auto tk1 = local_tile(t, Shape<_128, _8>{}, make_coord(0, 1));
    auto d1 = UMMA::make_umma_desc<UMMA::Major::MN>(tk1);
    printf("A MN-major 128x32 tf32: k0 desc %016llx  k1 desc %016llx\n", (unsigned long long)uint64_t(d0), (unsigned long long)uint64_t(d1));
    printf("  start %u lbo %u sbo %u ver %u base_off %u lbo_mode %u layout %u\n", (unsigned)d0.start_address_, (unsigned)d0.leading_byte_offset_,
           (unsigned)d0.stride_byte_offset_, (unsigned)d0.version_, (unsigned)d0.base_offset_, (unsigned)d0.lbo_mode_, (unsigned)d0.layout_type_);
    print(layout); printf("\n");
  }
  {
    auto layout = tile_to_shape(UMMA::Layout_K_SW128_Atom<T>{}, Shape<_128, _32>{});
    auto t = make_tensor(make_smem_ptr(reinterpret_cast<T*>(smem + 65536)), layout);
    auto tk = local_tile(t, Shape<_128, _8>{}, make_coord(0, 0));
    auto d0 = UMMA::make_umma_desc<UMMA::Major::K>(tk);
    auto tk1 = local_tile(t, Shape<_128, _8>{}, make_coord(0, 1));
    auto d1 = UMMA::make_umma_desc<UMMA::Major::K>(tk1);
    printf("B K-major 128x32 tf32: k0 desc %016llx  k1 desc %016llx\n", (unsigned long long)uint64_t(d0), (unsigned long long)uint64_t(d1));
    printf("  start %u lbo %u sbo %u ver %u layout %u\n", (unsigned)d0.start_address_, (unsigned)d0.leading_byte_offset_,
           (unsigned)d0.stride_byte_offset_, (unsigned)d0.version_, (unsigned)d0.layout_type_);
    print(layout); printf("\n");
  }
  auto id = UMMA::make_instr_desc<T, T, float, 128, 16, UMMA::Major::MN, UMMA::Major::K>();
  printf("idesc M128 N16 tf32 A=MN B=K: %08x\n", (unsigned)uint32_t(id));
}
int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  probe<<<1, 1, 160 * 1024>>>();
  printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

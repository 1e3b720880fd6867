// TEST INFRASTRUCTURE: host twin of corr_solve_kernel / corr_transfer_kernel / corr_mean_kernel
// (boxinstseg_b200/csrc/corr.cu).  It compiles the very header the kernels are made of (corr_core.cuh: corr_solve,
// corr_weighted, corr_transfer_pixel; a phase is a plain loop on the host) with g++ -ffp-contract=off, so that the CPU suite
// can check the kernels' arithmetic, indexing and phase structure against the oracle on a box without a GPU.  Never shipped.
#include <cmath>
#include <vector>

#include "../../boxinstseg_b200/csrc/corr_core.cuh"

extern "C" int host_corr_solve(const float* Cu, float* T, int64_t K, int64_t h, int64_t w, int dist_kernel, int num_iter,
                               int num_smooth) {
  const int64_t P = h * w, PP = P * P;
  std::vector<float> a(PP), b(PP), rs((1 + bxs::kCorrLanes) * P);
  for (int64_t k = 0; k < K; ++k)
    bxs::corr_solve(Cu + k * PP, T + k * PP, a.data(), b.data(), rs.data(), (int)h, (int)w, dist_kernel, num_iter, num_smooth);
  return 0;
}

extern "C" int host_corr_transfer(const float* T, const float* Cu, const float* m0, const float* m1, float* fg_ci, float* bg_ci,
                                  int64_t K, int64_t h, int64_t w, int64_t Hm, int64_t Wm) {
  const int64_t P = h * w, PP = P * P, M = Hm * Wm;
  std::vector<float> t2(PP), rs(P), mx(P), R(P), part(K * 2 * M);
  std::vector<bxs::CorrTap> tapV(Hm), tapU(Wm);
  for (int64_t i = 0; i < Hm; ++i) tapV[i] = bxs::corr_tap((int)i, (int)h, (int)Hm);
  for (int64_t i = 0; i < Wm; ++i) tapU[i] = bxs::corr_tap((int)i, (int)w, (int)Wm);
  for (int64_t k = 0; k < K; ++k) {
    bxs::corr_weighted(T + k * PP, Cu + k * PP, t2.data(), rs.data(), mx.data(), (int)P, [](float v) { return std::exp(v); });
    for (int64_t pq = 0; pq < M; ++pq)
      bxs::corr_transfer_pixel(t2.data(), m0[pq], m1 + k * M, R.data(), tapV.data(), tapU.data(), (int)h, (int)w, (int)Hm, (int)Wm, (int)pq,
                               &part[(k * 2 + 0) * M + pq], &part[(k * 2 + 1) * M + pq]);
  }
  for (int64_t pq = 0; pq < M; ++pq) {
    float a = 0.f, b = 0.f;
    for (int64_t k = 0; k < K; ++k) {
      a += part[(k * 2 + 0) * M + pq];
      b += part[(k * 2 + 1) * M + pq];
    }
    fg_ci[pq] = a / (float)K;
    bg_ci[pq] = b / (float)K;
  }
  return 0;
}

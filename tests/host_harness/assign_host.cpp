// TEST INFRASTRUCTURE: host twin of fcos_targets_kernel (boxinstseg_b200/csrc/assign_targets.cu).  It compiles the very
// header the kernel is made of (assign_core.cuh: fcos_point_target, fcos_emit, fcos_levels_from_host) with g++
// (-ffp-contract=off) and loops over (image, location) on the host, so that the arithmetic, the level lookup and the
// output layout of the kernel can be checked bit for bit against the oracle on a box without a GPU.  Never shipped: the
// product library exports only the CUDA path.
#include "../../boxinstseg_b200/csrc/assign_core.cuh"

extern "C" int host_fcos_targets(const float* points, const float* gt_boxes, const int64_t* gt_labels, const int64_t* gt_off,
                                 int64_t* labels, float* bbox_targets, int64_t* gt_inds, int64_t B, int64_t num_levels,
                                 const int64_t* level_off, const float* range_lo, const float* range_hi,
                                 const float* stride_radius, const float* stride, int center_sampling, int norm_on_bbox,
                                 int64_t num_classes) {
  bxs::FcosLevels lv;
  if (!bxs::fcos_levels_from_host(lv, num_levels, level_off, range_lo, range_hi, stride_radius, stride, center_sampling,
                                  norm_on_bbox, num_classes))
    return -1;
  const int64_t P = lv.level_off[num_levels];
  for (int64_t b = 0; b < B; ++b) {
    const int64_t g0 = gt_off[b];
    const int G = (int)(gt_off[b + 1] - g0);
    for (int64_t p = 0; p < P; ++p)
      bxs::fcos_emit(b, p, B, lv, points, gt_boxes + g0 * 4, gt_labels + g0, G, g0, labels, bbox_targets, gt_inds);
  }
  return 0;
}

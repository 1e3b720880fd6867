// f2: CondInstBoxHead.get_targets (mmdet/models/dense_heads/condinst_head.py:477-548) + _get_target_single (:550-633)
// for all images and all FPN levels of a batch in ONE launch.
//
// The reference runs, per image, ~25 element-wise ATen kernels over [points, gts] tensors (17 064 x G at 800 x 1024),
// then splits per level, concatenates per image and divides by the stride.  Here one thread owns one (image, point):
// the image's boxes sit in shared memory, the point's level record comes from the kernel parameters, and the three
// results are written directly in the reference's output layout (level-major, image-major inside a level), the index of
// the chosen ground truth already offset by the ground truths of the preceding images (:519-522).
#include "assign_core.cuh"
#include "common.cuh"

namespace bxs {
namespace {

constexpr int NT = 256;
constexpr int kSmemGts = 1024;      // 24 KB of boxes + labels; larger images read global memory directly
constexpr int kMaxImages = 256;

struct FcosImages { int64_t off[kMaxImages + 1]; };      // ground-truth offsets per image, by value (2 KB of parameters)

__global__ void __launch_bounds__(NT) fcos_targets_kernel(const float* __restrict__ points, const float* __restrict__ gt_boxes,
                                                          const int64_t* __restrict__ gt_labels,
                                                          const FcosImages im, int64_t* __restrict__ labels,
                                                          float* __restrict__ bbox_targets, int64_t* __restrict__ gt_inds,
                                                          int64_t B, const FcosLevels lv) {
  __shared__ float s_box[kSmemGts * 4];
  __shared__ int64_t s_lab[kSmemGts];
  const int64_t b = blockIdx.y;
  const int64_t g0 = im.off[b];
  const int G = (int)(im.off[b + 1] - g0);
  const bool staged = G <= kSmemGts;
  if (staged) {
    for (int i = threadIdx.x; i < G * 4; i += NT) s_box[i] = __ldg(gt_boxes + g0 * 4 + i);
    for (int i = threadIdx.x; i < G; i += NT) s_lab[i] = __ldg(gt_labels + g0 + i);
    __syncthreads();
  }
  const float* boxes = staged ? s_box : gt_boxes + g0 * 4;
  const int64_t* labs = staged ? s_lab : gt_labels + g0;
  const int64_t P = lv.level_off[lv.num_levels];
  for (int64_t p = blockIdx.x * (int64_t)NT + threadIdx.x; p < P; p += (int64_t)gridDim.x * NT)
    fcos_emit(b, p, B, lv, points, boxes, labs, G, g0, labels, bbox_targets, gt_inds);
}

}  // namespace
}  // namespace bxs

using namespace bxs;

extern "C" int bxs_fcos_targets(const float* points, const float* gt_boxes, const int64_t* gt_labels,
                                const int64_t* gt_off_host, int64_t* labels, float* bbox_targets, int64_t* gt_inds, int64_t B, int64_t num_levels,
                                const int64_t* level_off_host, const float* range_lo_host, const float* range_hi_host,
                                const float* stride_radius_host, const float* stride_host, int center_sampling,
                                int norm_on_bbox, int64_t num_classes, bxs_stream_t stream) {
  if (!points || !gt_off_host || !labels || !bbox_targets || !gt_inds || !level_off_host || !range_lo_host || !range_hi_host ||
      !stride_radius_host || !stride_host || B <= 0 || num_levels <= 0 || num_classes < 0)
    return BXS_ERR_INVALID_ARG;
  if (num_levels > kFcosMaxLevels || B > kMaxImages) return BXS_ERR_UNSUPPORTED;
  FcosImages im;
  for (int64_t b = 0; b <= kMaxImages; ++b) im.off[b] = gt_off_host[b <= B ? b : B];
  if (im.off[0] != 0) return BXS_ERR_INVALID_ARG;
  for (int64_t b = 0; b < B; ++b)
    if (im.off[b + 1] < im.off[b]) return BXS_ERR_INVALID_ARG;
  if (im.off[B] > 0 && (!gt_boxes || !gt_labels)) return BXS_ERR_INVALID_ARG;
  FcosLevels lv;
  if (!fcos_levels_from_host(lv, num_levels, level_off_host, range_lo_host, range_hi_host, stride_radius_host, stride_host,
                             center_sampling, norm_on_bbox, num_classes))
    return BXS_ERR_INVALID_ARG;
  const int64_t P = lv.level_off[num_levels];
  if (P == 0) return BXS_OK;
  const int64_t want = ceil_div(P, NT);
  const int gx = (int)(want < 4096 ? want : 4096);
  fcos_targets_kernel<<<dim3(gx, (unsigned)B), NT, 0, as_stream(stream)>>>(points, gt_boxes, gt_labels, im, labels,
                                                                           bbox_targets, gt_inds, B, lv);
  return check_launch();
}

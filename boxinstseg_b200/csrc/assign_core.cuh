// f2: per-point core of CondInstBoxHead._get_target_single (mmdet/models/dense_heads/condinst_head.py:550-633), the
// FCOS point assignment: for one location and the ground truths of its image, the ground truth of minimal area among
// those that contain the location (inside the centre-sampling box, :583-618) and regress it within the level's range
// (:620-624).  The reference materialises ~25 [points, gts] float tensors per image; here it is one pass in registers.
//
// Every floating-point value is the result of ONE correctly rounded fp32 operation of the reference, in the reference's
// order (differences, the area product, (a + b) / 2, centre -/+ stride, where-clamps, min/max), so the comparison
// results -- and with them every label, index and regression target -- are bit-identical.  The explicit *_rn
// intrinsics keep the compiler from contracting a product into a following sum.
//
// The body is BXS_HD so that tests/host_harness/assign_host.cpp can compile the very same lines for the host and
// check them against the oracle without a GPU (test infrastructure; the library itself exports only the CUDA path).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BXS_HD __host__ __device__ __forceinline__
#else
#define BXS_HD inline
#endif

namespace bxs {

#if defined(__CUDA_ARCH__)
#define BXS_FADD(a, b) __fadd_rn((a), (b))
#define BXS_FSUB(a, b) __fsub_rn((a), (b))
#define BXS_FMUL(a, b) __fmul_rn((a), (b))
#define BXS_FDIV(a, b) __fdiv_rn((a), (b))
#else                                     // host twin: built with -ffp-contract=off
#define BXS_FADD(a, b) ((a) + (b))
#define BXS_FSUB(a, b) ((a) - (b))
#define BXS_FMUL(a, b) ((a) * (b))
#define BXS_FDIV(a, b) ((a) / (b))
#endif

constexpr int kFcosMaxLevels = 8;
constexpr float kFcosInf = 1e8f;          // INF of condinst_head.py:16

struct FcosLevels {                        // by value in the kernel parameters: no H2D copy, graph-capturable
  int num_levels;
  int center_sampling;
  int norm_on_bbox;
  int num_classes;
  int64_t level_off[kFcosMaxLevels + 1];   // first point of every level in the concatenated point list; [L] = P
  float range_lo[kFcosMaxLevels], range_hi[kFcosMaxLevels];   // regress_ranges
  float stride_radius[kFcosMaxLevels];     // strides[l] * center_sample_radius, rounded to fp32 like the tensor assignment :595
  float stride[kFcosMaxLevels];            // strides[l] for norm_on_bbox (:542-543)
};

struct FcosTarget {
  int64_t label;      // class of the chosen ground truth, num_classes = background
  int64_t gt_ind;     // index into the image's ground truths, -1 = none
  float l, t, r, b;   // regression target w.r.t. the chosen ground truth (ground truth 0 when none, like :631)
};

BXS_HD float fmin4(float a, float b, float c, float d) {
  float m = a < b ? a : b;
  m = c < m ? c : m;
  return d < m ? d : m;
}
BXS_HD float fmax4(float a, float b, float c, float d) {
  float m = a > b ? a : b;
  m = c > m ? c : m;
  return d > m ? d : m;
}

// x, y: the location; boxes [G,4] (x1,y1,x2,y2), labels [G]; lo/hi: the level's regress range; sr: stride * radius.
BXS_HD FcosTarget fcos_point_target(float x, float y, const float* boxes, const int64_t* labels, int G, float lo, float hi,
                                    float sr, bool center_sampling, int num_classes) {
  FcosTarget out;
  out.label = num_classes;
  out.gt_ind = -1;
  out.l = out.t = out.r = out.b = 0.f;
  if (G <= 0) return out;                                   // :554-556 (no ground truth: background, zero targets)
  float best = 0.f;
  int best_g = -1;
  for (int g = 0; g < G; ++g) {
    const float x1 = boxes[4 * g], y1 = boxes[4 * g + 1], x2 = boxes[4 * g + 2], y2 = boxes[4 * g + 3];
    const float l = BXS_FSUB(x, x1), r = BXS_FSUB(x2, x), t = BXS_FSUB(y, y1), b = BXS_FSUB(y2, y);     // :571-574
    bool inside;
    if (center_sampling) {                                                                               // :577-615
      const float cx = BXS_FADD(x1, x2) / 2.f, cy = BXS_FADD(y1, y2) / 2.f;
      const float xmin = BXS_FSUB(cx, sr), ymin = BXS_FSUB(cy, sr), xmax = BXS_FADD(cx, sr), ymax = BXS_FADD(cy, sr);
      const float c0 = xmin > x1 ? xmin : x1, c1 = ymin > y1 ? ymin : y1;
      const float c2 = xmax > x2 ? x2 : xmax, c3 = ymax > y2 ? y2 : ymax;
      inside = fmin4(BXS_FSUB(x, c0), BXS_FSUB(y, c1), BXS_FSUB(c2, x), BXS_FSUB(c3, y)) > 0.f;
    } else {
      inside = fmin4(l, t, r, b) > 0.f;                                                                  // :618
    }
    const float far = fmax4(l, t, r, b);                                                                 // :621
    const bool in_range = far >= lo && far <= hi;
    float area = BXS_FMUL(BXS_FSUB(x2, x1), BXS_FSUB(y2, y1));                                           // :558-559
    if (!inside || !in_range) area = kFcosInf;                                                           // :626-627
    if (best_g < 0 || area < best) {          // areas.min(dim=1): the FIRST minimal entry
      best = area;
      best_g = g;
      out.l = l; out.t = t; out.r = r; out.b = b;
    }
  }
  if (best != kFcosInf) {                                                                                // :629-632
    out.label = labels[best_g];
    out.gt_ind = best_g;
  }
  return out;
}

// One (image b, location p): level lookup, assignment, and the three stores in the reference's output layout (level-major,
// image-major inside a level: what torch.cat of its per-level lists gives).  `boxes` / `labels` are the image's ground truths
// (G of them, the first one is entry g0 of the concatenated list).
BXS_HD void fcos_emit(int64_t b, int64_t p, int64_t B, const FcosLevels& lv, const float* points, const float* boxes,
                      const int64_t* labels, int G, int64_t g0, int64_t* out_labels, float* out_targets, int64_t* out_inds) {
  int l = 0;
  for (int k = 1; k < kFcosMaxLevels; ++k)
    if (k < lv.num_levels && p >= lv.level_off[k]) l = k;
  const FcosTarget t = fcos_point_target(points[2 * p], points[2 * p + 1], boxes, labels, G, lv.range_lo[l], lv.range_hi[l],
                                         lv.stride_radius[l], lv.center_sampling != 0, lv.num_classes);
  const int64_t Pl = lv.level_off[l + 1] - lv.level_off[l];
  const int64_t o = B * lv.level_off[l] + b * Pl + (p - lv.level_off[l]);
  out_labels[o] = t.label;
  out_inds[o] = t.gt_ind < 0 ? -1 : t.gt_ind + g0;
  float v0 = t.l, v1 = t.t, v2 = t.r, v3 = t.b;
  if (lv.norm_on_bbox) {                                    // :542-543, IEEE division
    const float s = lv.stride[l];
    v0 = BXS_FDIV(v0, s); v1 = BXS_FDIV(v1, s); v2 = BXS_FDIV(v2, s); v3 = BXS_FDIV(v3, s);
  }
  float* q = out_targets + 4 * o;
  q[0] = v0; q[1] = v1; q[2] = v2; q[3] = v3;
}

// Fills the by-value level record from the host arrays of the C ABI; false = invalid.
inline bool fcos_levels_from_host(FcosLevels& lv, int64_t num_levels, const int64_t* level_off, const float* range_lo,
                                  const float* range_hi, const float* stride_radius, const float* stride, int center_sampling,
                                  int norm_on_bbox, int64_t num_classes) {
  if (num_levels <= 0 || num_levels > kFcosMaxLevels) return false;
  lv.num_levels = (int)num_levels;
  lv.center_sampling = center_sampling;
  lv.norm_on_bbox = norm_on_bbox;
  lv.num_classes = (int)num_classes;
  for (int l = 0; l <= kFcosMaxLevels; ++l) lv.level_off[l] = level_off[l <= num_levels ? l : num_levels];
  for (int l = 0; l < kFcosMaxLevels; ++l) {
    const int k = l < num_levels ? l : (int)num_levels - 1;
    lv.range_lo[l] = range_lo[k];
    lv.range_hi[l] = range_hi[k];
    lv.stride_radius[l] = stride_radius[k];
    lv.stride[l] = stride[k];
  }
  if (lv.level_off[0] != 0) return false;
  for (int l = 0; l < num_levels; ++l)
    if (lv.level_off[l + 1] < lv.level_off[l]) return false;
  return true;
}

}  // namespace bxs

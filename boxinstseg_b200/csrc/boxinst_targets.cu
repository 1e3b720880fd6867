// a5: BoxInst target building on the GPU -- de-normalise -> uint8 -> stride x stride mean ->
// uint8 -> CIE-LAB -> 8-neighbour colour similarity -> threshold bits; box -> grid rectangle.
// Replaces CondInstMaskHead.get_targets / get_bitmasks_from_boxes / get_original_image /
// get_image_color_similarity (condinst_head.py:170-246, 1345-1448), which round-trip every image
// through the host (mmcv.tensor2imgs, skimage.color.rgb2lab) and loop over GT boxes in Python.
//
// Exactness: everything up to the uint8 downsampled image is integer work and is bit-exact
// (two separately rounded fp32 ops reproduce OpenCV's multiply/add; sums of <=255*s*s are exact).
// LAB is evaluated in float64 exactly as scikit-image does and only then cast to float32.
#include "common.cuh"

namespace bxs {
namespace {

struct Norm { float mean[3]; float std[3]; };

// scikit-image rgb2lab constants: sRGB D65 matrix, D65 2-degree white point
__device__ __forceinline__ void rgb_u8_to_lab(int r, int g, int b, float* lab) {
  double c[3] = {r / 255.0, g / 255.0, b / 255.0};
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = c[k] > 0.04045 ? pow((c[k] + 0.055) / 1.055, 2.4) : c[k] / 12.92;
  double x = (0.412453 * c[0] + 0.357580 * c[1] + 0.180423 * c[2]) / 0.95047;
  double y = (0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]) / 1.0;
  double z = (0.019334 * c[0] + 0.119193 * c[1] + 0.950227 * c[2]) / 1.08883;
  double f[3] = {x, y, z};
#pragma unroll
  for (int k = 0; k < 3; ++k) f[k] = f[k] > 0.008856 ? cbrt(f[k]) : 7.787 * f[k] + 16.0 / 116.0;
  lab[0] = (float)(116.0 * f[1] - 16.0);
  lab[1] = (float)(500.0 * (f[0] - f[1]));
  lab[2] = (float)(200.0 * (f[1] - f[2]));
}

// Python slice semantics of `full[int(y1):int(y2)+1, int(x1):int(x2)+1] = 1` followed by
// `[s//2::s, s//2::s]`  (condinst_head.py:1429-1432)
__device__ __forceinline__ void slice_range(float lo_f, float hi_f, int len, int s, int& g0, int& g1) {
  int start = (int)lo_f, stop = (int)hi_f + 1;       // int(): truncation toward zero
  if (start < 0) start += len;
  if (stop < 0) stop += len;
  start = min(max(start, 0), len);
  stop = min(max(stop, 0), len);
  // sampled positions s/2 + k*s inside [start, stop)
  const int h = s / 2;
  g0 = start <= h ? 0 : (start - h + s - 1) / s;
  g1 = stop - 1 >= h ? (stop - 1 - h) / s : -1;
  const int glen = (len - h + s - 1) / s;
  g1 = min(g1, glen - 1);
}

// one output pixel (b, j, i) of the stride-s LAB image and its validity
__device__ __forceinline__ void lab_pixel(const float* __restrict__ img, const Norm& nm, int ih, int iw, int removed,
                                          float* __restrict__ lab, uint8_t* __restrict__ valid, int b, int j, int i, int H,
                                          int W, int Hp, int Wp, int s) {
  int u8[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* plane = img + ((int64_t)b * 3 + c) * Hp * Wp;
    int sum = 0;
    for (int dy = 0; dy < s; ++dy) {
      const int y = j * s + dy;
      if (y >= ih) break;                                       // zero padding below the image
      for (int dx = 0; dx < s; ++dx) {
        const int x = i * s + dx;
        if (x >= iw) break;
        // tensor2imgs: float32 multiply, float32 add (separately rounded), astype(uint8)
        float f = __fadd_rn(__fmul_rn(__ldg(plane + (int64_t)y * Wp + x), nm.std[c]), nm.mean[c]);
        sum += (int)fminf(fmaxf(truncf(f), 0.f), 255.f);
      }
    }
    // avg_pool2d in fp32 (exact integer sum), then .byte() truncation
    u8[c] = (int)truncf(__fdiv_rn((float)sum, (float)(s * s)));
  }
  float l3[3];
  rgb_u8_to_lab(u8[0], u8[1], u8[2], l3);
  const int64_t plane_sz = (int64_t)H * W, o = (int64_t)j * W + i;
  lab[((int64_t)b * 3 + 0) * plane_sz + o] = l3[0];
  lab[((int64_t)b * 3 + 1) * plane_sz + o] = l3[1];
  lab[((int64_t)b * 3 + 2) * plane_sz + o] = l3[2];
  const int sy = j * s + s / 2, sx = i * s + s / 2;
  valid[(int64_t)b * plane_sz + o] = (sy < ih - removed && sx < iw) ? 1 : 0;
}

__global__ void lab_kernel(const float* __restrict__ img, const int32_t* __restrict__ img_hw,
                           const int32_t* __restrict__ removed_rows, Norm nm, float* __restrict__ lab,
                           uint8_t* __restrict__ valid, int B, int Hp, int Wp, int s) {
  const int H = Hp / s, W = Wp / s;
  const int64_t total = (int64_t)B * H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = idx % W, j = (idx / W) % H, b = idx / ((int64_t)W * H);
    lab_pixel(img, nm, img_hw[2 * b], img_hw[2 * b + 1], removed_rows[b], lab, valid, b, j, i, H, W, Hp, Wp, s);
  }
}

// The per-image metadata of a batch travels BY VALUE in the kernel parameters (no host-to-device copy, nothing to keep
// alive, capturable in a CUDA graph): image sizes, removed bottom rows, and the running GT count (gt_end[b] = number of
// GT boxes of images 0..b).
constexpr int kMaxMetaImages = 64;
struct TargetsMeta {
  int ih[kMaxMetaImages], iw[kMaxMetaImages], removed[kMaxMetaImages], gt_end[kMaxMetaImages];
  Norm nm;
};

// lab + validity of every stride-s pixel; the tail of the grid also turns the GT boxes into grid rectangles and writes
// the image index of every GT (get_bitmasks_from_boxes' per-image loop, condinst_head.py:1417-1448)
__global__ void lab_rects_kernel(const float* __restrict__ img, const float* __restrict__ boxes,
                                 const __grid_constant__ TargetsMeta meta, float* __restrict__ lab, uint8_t* __restrict__ valid,
                                 int32_t* __restrict__ rects, int32_t* __restrict__ gt_img, int B, int G, int Hp, int Wp, int s) {
  const int H = Hp / s, W = Wp / s;
  const int64_t total = (int64_t)B * H * W;
  const int64_t tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = tid0; idx < total; idx += nth) {
    const int i = idx % W, j = (idx / W) % H, b = idx / ((int64_t)W * H);
    lab_pixel(img, meta.nm, meta.ih[b], meta.iw[b], meta.removed[b], lab, valid, b, j, i, H, W, Hp, Wp, s);
  }
  for (int64_t g = nth - 1 - tid0; g < G; g += nth) {            // the last threads of the grid: least pixel work
    const float x1 = boxes[4 * g], y1 = boxes[4 * g + 1], x2 = boxes[4 * g + 2], y2 = boxes[4 * g + 3];
    int j0, j1, i0, i1;
    slice_range(y1, y2, Hp, s, j0, j1);
    slice_range(x1, x2, Wp, s, i0, i1);
    *reinterpret_cast<int4*>(rects + 4 * g) = make_int4(j0, j1, i0, i1);
    int b = 0;
    while (b < B - 1 && g >= meta.gt_end[b]) ++b;
    gt_img[g] = b;
  }
}

__global__ void similarity_kernel(const float* __restrict__ lab, const uint8_t* __restrict__ valid,
                                  float* __restrict__ sim, uint8_t* __restrict__ edge_bits, int B, int H, int W,
                                  int size, int dil, float thresh) {
  const int R = (size / 2) * dil, K = size * size - 1;
  const int64_t plane = (int64_t)H * W, total = (int64_t)B * plane;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = idx % W, y = (idx / W) % H, b = idx / plane;
    const float* L = lab + (int64_t)b * 3 * plane;
    const float p0 = L[(int64_t)y * W + x], p1 = L[plane + (int64_t)y * W + x], p2 = L[2 * plane + (int64_t)y * W + x];
    unsigned bits = 0;
    int c = 0;
    for (int dy = -R; dy <= R; dy += dil)
      for (int dx = -R; dx <= R; dx += dil) {
        if (dy == 0 && dx == 0) continue;
        const int qy = y + dy, qx = x + dx;
        float s = 0.f;      // out-of-image neighbour: the unfolded mask is zero padded
        if (qy >= 0 && qy < H && qx >= 0 && qx < W && valid[(int64_t)b * plane + (int64_t)qy * W + qx]) {
          const int64_t q = (int64_t)qy * W + qx;
          const float d0 = p0 - L[q], d1 = p1 - L[plane + q], d2 = p2 - L[2 * plane + q];
          const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
          s = expf(-__fmul_rn(__fsqrt_rn(n2), 0.5f));
        }
        if (sim) sim[(((int64_t)b * K + c) * H + y) * W + x] = s;
        if (s >= thresh && c < 8) bits |= 1u << c;
        ++c;
      }
    if (edge_bits) edge_bits[idx] = (uint8_t)bits;
  }
}

__global__ void rects_kernel(const float* __restrict__ boxes, int32_t* __restrict__ rects, int G, int Hp, int Wp,
                             int s) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const float x1 = boxes[4 * g], y1 = boxes[4 * g + 1], x2 = boxes[4 * g + 2], y2 = boxes[4 * g + 3];
  int j0, j1, i0, i1;
  slice_range(y1, y2, Hp, s, j0, j1);
  slice_range(x1, x2, Wp, s, i0, i1);
  rects[4 * g + 0] = j0; rects[4 * g + 1] = j1; rects[4 * g + 2] = i0; rects[4 * g + 3] = i1;
}

__global__ void bitmask_kernel(const int32_t* __restrict__ rects, float* __restrict__ out, int64_t G, int H, int W) {
  const int64_t total = G * H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = idx % W, y = (idx / W) % H;
    const int64_t g = idx / ((int64_t)W * H);
    const int4 r = *reinterpret_cast<const int4*>(rects + 4 * g);
    out[idx] = (y >= r.x && y <= r.y && x >= r.z && x <= r.w) ? 1.f : 0.f;
  }
}

inline int grid_for(int64_t total, int block = 256) {
  int64_t g = (total + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace
}  // namespace bxs

using namespace bxs;

extern "C" int bxs_boxinst_lab(const float* img, const int32_t* img_hw, const int32_t* removed_rows,
                               const float* mean3_host, const float* std3_host, float* lab, uint8_t* valid,
                               int64_t B, int64_t Hp, int64_t Wp, int stride, bxs_stream_t stream) {
  if (!img || !img_hw || !removed_rows || !mean3_host || !std3_host || !lab || !valid || B <= 0 || Hp <= 0 ||
      Wp <= 0 || stride <= 0 || stride > 256 || Hp % stride || Wp % stride)
    return BXS_ERR_INVALID_ARG;
  Norm nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3_host[c]; nm.std[c] = std3_host[c]; }
  const int64_t total = B * (Hp / stride) * (Wp / stride);
  lab_kernel<<<grid_for(total, 128), 128, 0, as_stream(stream)>>>(img, img_hw, removed_rows, nm, lab, valid, (int)B,
                                                                  (int)Hp, (int)Wp, stride);
  return check_launch();
}

extern "C" int bxs_boxinst_similarity(const float* lab, const uint8_t* valid, float* sim, uint8_t* edge_bits,
                                      int64_t B, int64_t H, int64_t W, int size, int dilation, float thresh,
                                      bxs_stream_t stream) {
  if (!lab || !valid || (!sim && !edge_bits) || B <= 0 || H <= 0 || W <= 0 || size < 3 || !(size & 1) ||
      dilation < 1 || (edge_bits && size != 3))
    return BXS_ERR_INVALID_ARG;
  similarity_kernel<<<grid_for(B * H * W, 128), 128, 0, as_stream(stream)>>>(lab, valid, sim, edge_bits, (int)B, (int)H,
                                                                             (int)W, size, dilation, thresh);
  return check_launch();
}

extern "C" int bxs_boxinst_rects(const float* boxes, int32_t* rects, int64_t G, int64_t Hp, int64_t Wp, int stride,
                                 bxs_stream_t stream) {
  if (G == 0) return BXS_OK;
  if (!boxes || !rects || G < 0 || Hp <= 0 || Wp <= 0 || stride <= 0) return BXS_ERR_INVALID_ARG;
  rects_kernel<<<(unsigned)ceil_div(G, 128), 128, 0, as_stream(stream)>>>(boxes, rects, (int)G, (int)Hp, (int)Wp, stride);
  return check_launch();
}

extern "C" int bxs_boxinst_bitmasks(const int32_t* rects, float* bitmasks, int64_t G, int64_t H, int64_t W,
                                    bxs_stream_t stream) {
  if (G == 0) return BXS_OK;
  if (!rects || !bitmasks || G < 0 || H <= 0 || W <= 0) return BXS_ERR_INVALID_ARG;
  bitmask_kernel<<<grid_for(G * H * W), 256, 0, as_stream(stream)>>>(rects, bitmasks, G, (int)H, (int)W);
  return check_launch();
}


// ---------------------------------------------------------------------------------------
// Whole target build of a batch in two launches, metadata by value (<= 64 images): LAB + validity + grid rectangles +
// GT image index, then the colour similarity / edge bits.  Same results as bxs_boxinst_lab + bxs_boxinst_similarity +
// bxs_boxinst_rects; no host-to-device copies, capturable in a CUDA graph.
// ---------------------------------------------------------------------------------------
extern "C" int bxs_boxinst_targets_forward(const float* img, const float* boxes, const int32_t* img_hw_host,
                                           const int32_t* removed_rows_host, const int32_t* num_gts_host,
                                           const float* mean3_host, const float* std3_host, float* lab, uint8_t* valid,
                                           float* sim, uint8_t* edge_bits, int32_t* rects, int32_t* gt_img, int64_t B,
                                           int64_t Hp, int64_t Wp, int stride, int size, int dilation, float thresh,
                                           bxs_stream_t stream) {
  if (!img || !img_hw_host || !removed_rows_host || !num_gts_host || !mean3_host || !std3_host || !lab || !valid ||
      (!sim && !edge_bits) || B <= 0 || Hp <= 0 || Wp <= 0 || stride <= 0 || stride > 256 || Hp % stride || Wp % stride ||
      size < 3 || !(size & 1) || dilation < 1 || (edge_bits && size != 3))
    return BXS_ERR_INVALID_ARG;
  if (B > kMaxMetaImages) return BXS_ERR_UNSUPPORTED;
  TargetsMeta meta;
  int64_t G = 0;
  for (int b = 0; b < kMaxMetaImages; ++b) {
    const bool in = b < B;
    if (in && num_gts_host[b] < 0) return BXS_ERR_INVALID_ARG;
    if (in) G += num_gts_host[b];
    meta.ih[b] = in ? img_hw_host[2 * b] : 0;
    meta.iw[b] = in ? img_hw_host[2 * b + 1] : 0;
    meta.removed[b] = in ? removed_rows_host[b] : 0;
    meta.gt_end[b] = (int)G;
  }
  if (G > 0 && (!boxes || !rects || !gt_img)) return BXS_ERR_INVALID_ARG;
  for (int c = 0; c < 3; ++c) { meta.nm.mean[c] = mean3_host[c]; meta.nm.std[c] = std3_host[c]; }
  const int64_t H = Hp / stride, W = Wp / stride, total = B * H * W;
  cudaStream_t st = as_stream(stream);
  lab_rects_kernel<<<(unsigned)std::max<int64_t>(ceil_div(total, 128), ceil_div(std::min<int64_t>(G, total), 128)), 128, 0, st>>>(
      img, boxes, meta, lab, valid, rects, gt_img, (int)B, (int)G, (int)Hp, (int)Wp, stride);
  int rc = check_launch();
  if (rc != BXS_OK) return rc;
  similarity_kernel<<<grid_for(total, 128), 128, 0, st>>>(lab, valid, sim, edge_bits, (int)B, (int)H, (int)W, size, dilation,
                                                           thresh);
  return check_launch();
}

// CIE-LAB of packed uint8 RGB triplets (test hook for the colour conversion: all 2^24 colours against the oracle)
namespace bxs { namespace {
__global__ void lab_of_u8_kernel(const uint8_t* __restrict__ rgb, float* __restrict__ lab, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    rgb_u8_to_lab(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], lab + 3 * i);
}
} }
extern "C" int bxs_rgb_u8_to_lab(const uint8_t* rgb, float* lab, int64_t n, bxs_stream_t stream) {
  if (n == 0) return BXS_OK;
  if (!rgb || !lab || n < 0) return BXS_ERR_INVALID_ARG;
  lab_of_u8_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(rgb, lab, n);
  return check_launch();
}

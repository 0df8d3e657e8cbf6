// Validator path: the device side of DetectionValidator.update_metrics.
//   b2y_box_iou           <- metrics.box_iou (ultralytics/utils/metrics.py:52-71)
//   b2y_match_predictions <- DetectionValidator._process_batch (models/yolo/detect/val.py:209-228) =
//                            box_iou(gt, det) + BaseValidator.match_predictions (engine/validator.py:222-258, use_scipy=False)
//   b2y_preprocess_chw    <- DetectionValidator.preprocess (val.py:47-50: batch["img"].float() / 255 of the dataloader's
//                            letterboxed (B,3,H,W) uint8 RGB tensor), written in the network's space-to-depth bf16 layout
// All arithmetic is the reference's fp32 op sequence with IEEE rounding (no FMA contraction: this file is compiled with
// -fmad=false and uses the _rn intrinsics), so the IoU matrix and the true-positive matrix are bit-exact.
//
// match_predictions restated.  For one threshold t the reference lists all (label, detection) pairs of equal class with
// IoU >= t, orders them by IoU descending, keeps the first pair of every detection (np.unique on the detection column:
// the detection's best label) -- which leaves the list ordered by detection index -- and then the first pair of every label
// (np.unique on the label column: the label's LOWEST-INDEX detection, not its best one; the re-sort that would change
// this is commented out at validator.py:254).  A detection's best label does not depend on t, only whether its IoU passes,
// so one pass finds (best label, best IoU) per detection and the ten thresholds reuse it:
//     tp[d][i] = best_iou[d] >= t_i  and  d == min{ d' : best[d'] == best[d], best_iou[d'] >= t_i }.
// Exact IoU ties between two labels of one detection go to the higher label index (numpy's argsort()[::-1] order for the
// short match lists it insertion-sorts; for longer lists numpy's order among exact ties is unspecified).
#include "common.cuh"

namespace {
constexpr int NT = 256;
constexpr int MAX_IOUV = 16;

__device__ __forceinline__ float iou_xyxy(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2,
                                          float by2, float eps) {
  // inter = (min(a2, b2) - max(a1, b1)).clamp_(0).prod(2);  iou = inter / (area_a + area_b - inter + eps)
  const float iw = fmaxf(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float area_a = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
  const float area_b = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
  return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), eps));
}

__global__ void __launch_bounds__(NT) box_iou_kernel(const float *__restrict__ box1, int n, const float *__restrict__ box2,
                                                     int m, float eps, float *__restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * NT + threadIdx.x;
  if (idx >= (long long)n * m) return;
  const int i = (int)(idx / m), j = (int)(idx % m);
  const float4 a = __ldg(reinterpret_cast<const float4 *>(box1) + i);
  const float4 b = __ldg(reinterpret_cast<const float4 *>(box2) + j);
  out[idx] = iou_xyxy(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, eps);
}

struct MatchParams {
  const float *dets;     // B x max_det x 6 [x1,y1,x2,y2,conf,cls]
  const int *counts;     // B
  const float *labels;   // B x max_labels x 5 [cls,x1,y1,x2,y2]
  const int *nlabels;    // B
  float iouv[MAX_IOUV];
  int niou, B, max_det, max_labels;
  uint8_t *tp;           // B x max_det x niou
};

// One CTA per image.  Shared memory: labels (5 floats each), first[label][threshold] (lowest claiming detection), and
// per detection its best label / best IoU.
__global__ void __launch_bounds__(NT) match_predictions_kernel(MatchParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  const int b = blockIdx.x, tid = threadIdx.x;
  float *lab = reinterpret_cast<float *>(smem_raw);
  int *first = reinterpret_cast<int *>(lab + (size_t)p.max_labels * 5);
  int *best = first + (size_t)p.max_labels * p.niou;
  float *best_iou = reinterpret_cast<float *>(best + p.max_det);
  pdl_wait();
  const int n = min(max(p.counts[b], 0), p.max_det), m = min(max(p.nlabels[b], 0), p.max_labels);
  for (int i = tid; i < m * 5; i += NT) lab[i] = p.labels[(size_t)b * p.max_labels * 5 + i];
  for (int i = tid; i < m * p.niou; i += NT) first[i] = 0x7fffffff;
  __syncthreads();
  const float *det = p.dets + (size_t)b * p.max_det * 6;
  for (int d = tid; d < n; d += NT) {
    const float x1 = det[d * 6], y1 = det[d * 6 + 1], x2 = det[d * 6 + 2], y2 = det[d * 6 + 3], cls = det[d * 6 + 5];
    int bl = -1;
    float bi = 0.f;
    for (int l = 0; l < m; ++l) {
      const float *g = lab + l * 5;
      if (g[0] != cls) continue;  // correct_class = true_classes[:, None] == pred_classes (validator.py:238)
      const float v = iou_xyxy(g[1], g[2], g[3], g[4], x1, y1, x2, y2, 1e-7f);
      if (v >= bi && v > 0.f) { bi = v; bl = l; }
    }
    best[d] = bl;
    best_iou[d] = bi;
    if (bl >= 0)
      for (int i = 0; i < p.niou; ++i)
        if (bi >= p.iouv[i]) atomicMin(&first[bl * p.niou + i], d);
  }
  __syncthreads();
  uint8_t *tp = p.tp + (size_t)b * p.max_det * p.niou;
  for (int e = tid; e < p.max_det * p.niou; e += NT) {
    const int d = e / p.niou, i = e - d * p.niou;
    uint8_t v = 0;
    if (d < n) {
      const int bl = best[d];
      v = (bl >= 0 && best_iou[d] >= p.iouv[i] && first[bl * p.niou + i] == d) ? 1 : 0;
    }
    tp[e] = v;
  }
}

// (B,3,H,W) uint8 RGB planes -> (B, H/2, W/2, 16) bf16 space-to-depth, channel ((y&1)*2 + (x&1))*4 + {R,G,B,0}; value / 255
// correctly rounded, then rounded to bf16.  One thread = 4 x 2 pixels = two super-pixels = 64 contiguous output bytes.
__global__ void __launch_bounds__(NT) preprocess_chw_kernel(const uint8_t *__restrict__ src, uint4 *__restrict__ dst, int B, int H,
                                                            int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int W4 = W >> 2, H2 = H >> 1;
  const long long total = (long long)B * H2 * W4;
  for (long long idx = (long long)blockIdx.x * NT + threadIdx.x; idx < total; idx += (long long)gridDim.x * NT) {
    const int xq = (int)(idx % W4);
    const int ys = (int)((idx / W4) % H2);
    const int b = (int)(idx / ((long long)W4 * H2));
    uchar4 px[3][2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        px[c][r] = __ldg(reinterpret_cast<const uchar4 *>(src + (((size_t)b * 3 + c) * H + 2 * ys + r) * W) + xq);
    uint4 *o = dst + (((size_t)b * H2 + ys) * (W >> 1) + 2 * xq) * 2;
#pragma unroll
    for (int s = 0; s < 2; ++s) {  // super-pixel s: source columns 2s, 2s+1
      float f[16];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uchar4 v = px[c][r];
            const int k = 2 * s + q;
            const unsigned char byte = k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
            f[(r * 2 + q) * 4 + c] = __fdiv_rn((float)byte, 255.f);
          }
          f[(r * 2 + q) * 4 + 3] = 0.f;
        }
      o[2 * s] = pack8(f);
      o[2 * s + 1] = pack8(f + 8);
    }
  }
}
}  // namespace

extern "C" int b2y_box_iou(const void *box1, int32_t n, const void *box2, int32_t m, float eps, void *iou, void *stream) {
  B2Y_REQUIRE(n >= 0 && m >= 0 && (n == 0 || m == 0 || (box1 && box2 && iou)), B2Y_ERR_INVALID, "box_iou: bad arguments");
  if (n == 0 || m == 0) return B2Y_OK;
  B2Y_REQUIRE(((reinterpret_cast<uintptr_t>(box1) | reinterpret_cast<uintptr_t>(box2)) & 15) == 0, B2Y_ERR_INVALID,
              "box_iou: boxes must be 16-byte aligned (N,4) fp32");
  B2YDeviceGuard guard(b2y_device_of(iou));
  const long long total = (long long)n * m;
  B2Y_REQUIRE(total < (1ll << 31) * (long long)NT, B2Y_ERR_UNSUPPORTED, "box_iou: matrix too large");
  b2y_launch(box_iou_kernel, dim3((unsigned)((total + NT - 1) / NT)), dim3(NT), 0, (cudaStream_t)stream,
             reinterpret_cast<const float *>(box1), n, reinterpret_cast<const float *>(box2), m, eps,
             reinterpret_cast<float *>(iou));
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

extern "C" int b2y_match_predictions(const void *dets, const void *counts, const void *labels, const void *label_counts,
                                     const float *iouv, int32_t niou, int32_t B, int32_t max_det, int32_t max_labels,
                                     void *tp, void *stream) {
  B2Y_REQUIRE(dets && counts && label_counts && iouv && tp && B > 0 && max_det > 0 && max_labels >= 0 &&
                  (labels || max_labels == 0),
              B2Y_ERR_INVALID, "match_predictions: bad arguments");
  B2Y_REQUIRE(niou > 0 && niou <= MAX_IOUV, B2Y_ERR_UNSUPPORTED, "match_predictions: 1..%d IoU thresholds", MAX_IOUV);
  MatchParams p;
  for (int i = 0; i < niou; ++i) {
    B2Y_REQUIRE(iouv[i] > 0.f, B2Y_ERR_INVALID, "match_predictions: IoU thresholds must be positive");
    p.iouv[i] = iouv[i];
  }
  const size_t smem = ((size_t)max_labels * (5 + niou) + 2 * (size_t)max_det) * 4;
  B2Y_REQUIRE(smem <= 200 * 1024, B2Y_ERR_UNSUPPORTED,
              "match_predictions: %d labels x %d detections per image do not fit shared memory", max_labels, max_det);
  B2YDeviceGuard guard(b2y_device_of(dets));
  p.dets = reinterpret_cast<const float *>(dets);
  p.counts = reinterpret_cast<const int *>(counts);
  p.labels = reinterpret_cast<const float *>(labels);
  p.nlabels = reinterpret_cast<const int *>(label_counts);
  p.niou = niou; p.B = B; p.max_det = max_det; p.max_labels = max_labels;
  p.tp = reinterpret_cast<uint8_t *>(tp);
  B2Y_CUDA_CHECK(cudaFuncSetAttribute(match_predictions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  b2y_launch(match_predictions_kernel, dim3(B), dim3(NT), smem, (cudaStream_t)stream, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

extern "C" int b2y_preprocess_chw(const void *src, int32_t B, int32_t H, int32_t W, void *dst, void *stream) {
  B2Y_REQUIRE(src && dst && B > 0 && H > 0 && W > 0, B2Y_ERR_INVALID, "preprocess_chw: bad arguments");
  B2Y_REQUIRE(H % 2 == 0 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              B2Y_ERR_UNSUPPORTED, "preprocess_chw: H must be even, W a multiple of 4, src 4-byte and dst 16-byte aligned");
  B2YDeviceGuard guard(b2y_device_of(dst));
  const long long total = (long long)B * (H / 2) * (W / 4);
  const long long blocks = (total + NT - 1) / NT;
  b2y_launch(preprocess_chw_kernel, dim3((unsigned)(blocks < 148 * 8 ? blocks : 148 * 8)), dim3(NT), 0, (cudaStream_t)stream,
             reinterpret_cast<const uint8_t *>(src), reinterpret_cast<uint4 *>(dst), B, H, W);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

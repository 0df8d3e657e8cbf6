// Detect-derived heads: the per-anchor extra columns of a Pose head.
//
// The reference decodes the keypoint branch for ALL anchors (Pose.kpts_decode, ultralytics/nn/modules/head.py:252-279),
// concatenates the (B, nk, A) rows under y (head.py:250) and drags them through non_max_suppression as "mask" columns
// (utils/ops.py:245-246, 268-273); PosePredictor.postprocess then reshapes the kept rows and maps them to the original
// image with ops.scale_coords (models/yolo/pose/predict.py:39-59, utils/ops.py:740-772, clip_coords :341-358).
// Here NMS runs on (B, 4+nc, A) exactly as for Detect and reports the anchor of every kept row (b2y_nms kept_anchor);
// this kernel gathers those <= max_det rows from the branch's raw fp32 NHWC maps and decodes / rescales only them:
// one launch, B * max_det * ne elements instead of three passes over B * A * ne.
//
// Arithmetic follows the reference's fp32 operation order with IEEE rounding (no FMA contraction: this file is compiled
// with -fmad=false and uses the _rn intrinsics), so given equal raw maps the x/y coordinates are bit-identical;
// the visibility sigmoid uses expf and is compared with a tolerance.
#include "common.cuh"

namespace {
constexpr int MAX_LEVELS = 4;
struct GatherParams {
  const float *raw[MAX_LEVELS];
  int hw[MAX_LEVELS], w[MAX_LEVELS], a0[MAX_LEVELS + 1];
  float stride[MAX_LEVELS];
  int n_levels, pitch, ne, ndim, mode, B, max_det;
  const int *kept, *counts;
  const float *geom;  // optional (B, 5): gain, padx, pady, w0, h0 of ops.scale_coords / clip_coords
  float *out;         // (B, max_det, ne)
  float *round_dets;  // optional (B, max_det, 6): box corners rounded half-to-even in place (pose/predict.py:52 `.round()`)
};

// one warp per (image, kept row); lanes stride over the ne extra channels (contiguous in the NHWC map and in `out`)
__global__ void __launch_bounds__(256) gather_extras_kernel(GatherParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= p.max_det) return;
  float *dst = p.out + ((size_t)b * p.max_det + row) * p.ne;
  if (row >= p.counts[b]) {  // rows past the count are zero, as b2y_nms leaves dets
    for (int k = lane; k < p.ne; k += 32) dst[k] = 0.f;
    return;
  }
  if (p.round_dets && lane < 4) {
    float *d = p.round_dets + ((size_t)b * p.max_det + row) * 6 + lane;
    *d = rintf(*d);
  }
  const int a = p.kept[(size_t)b * p.max_det + row];
  if (a < 0 || a >= p.a0[p.n_levels]) {  // cannot happen with b2y_nms's kept_anchor; never read out of bounds
    for (int k = lane; k < p.ne; k += 32) dst[k] = 0.f;
    return;
  }
  int lvl = 0;
  while (lvl + 1 < p.n_levels && a >= p.a0[lvl + 1]) ++lvl;
  const int ai = a - p.a0[lvl];  // anchor index within the level, row-major (utils/tal.py:371-383)
  const float *src = p.raw[lvl] + ((size_t)b * p.hw[lvl] + ai) * p.pitch;
  const float ax = (float)(ai % p.w[lvl]), ay = (float)(ai / p.w[lvl]);  // (anchor + 0.5) - 0.5, exact in fp32
  const float st = p.stride[lvl];
  float gain = 1.f, padx = 0.f, pady = 0.f, w0 = 0.f, h0 = 0.f;
  if (p.geom) {
    gain = p.geom[b * 5 + 0]; padx = p.geom[b * 5 + 1]; pady = p.geom[b * 5 + 2];
    w0 = p.geom[b * 5 + 3]; h0 = p.geom[b * 5 + 4];
  }
  for (int k = lane; k < p.ne; k += 32) {
    float v = src[k];
    if (p.mode == 1) {
      const int d = k % p.ndim;
      if (d == 0 || d == 1) {
        v = __fmul_rn(__fadd_rn(__fmul_rn(v, 2.0f), d == 0 ? ax : ay), st);  // head.py:277-278
        if (p.geom)  // ops.py:763-768, 353-354
          v = fminf(fmaxf(__fdiv_rn(__fsub_rn(v, d == 0 ? padx : pady), gain), 0.f), d == 0 ? w0 : h0);
      } else if (d == 2) {
        v = 1.f / (1.f + expf(-v));  // head.py:276
      }
    }
    dst[k] = v;
  }
}
}  // namespace

extern "C" int b2y_gather_extras(const void *const *raw, const int32_t *hs, const int32_t *ws, const float *strides,
                                 int32_t n_levels, int32_t pitch, int32_t ne, int32_t ndim, int32_t mode,
                                 const void *kept_anchor, const void *counts, const void *geom, int32_t B,
                                 int32_t max_det, void *out, void *round_dets, void *stream) {
  B2Y_REQUIRE(raw && hs && ws && strides && kept_anchor && counts && out, B2Y_ERR_INVALID, "gather_extras: null argument");
  B2Y_REQUIRE(n_levels >= 1 && n_levels <= MAX_LEVELS && B > 0 && max_det > 0 && ne > 0 && pitch >= ne, B2Y_ERR_INVALID,
              "gather_extras: bad sizes (levels %d, B %d, max_det %d, ne %d, pitch %d)", n_levels, B, max_det, ne, pitch);
  B2Y_REQUIRE(mode == 0 || (mode == 1 && (ndim == 2 || ndim == 3) && ne % ndim == 0), B2Y_ERR_INVALID,
              "gather_extras: mode %d / ndim %d (keypoints are (x, y) or (x, y, visibility) groups)", mode, ndim);
  B2YDeviceGuard guard(b2y_device_of(out));
  GatherParams p;
  int a0 = 0;
  for (int l = 0; l < n_levels; ++l) {
    B2Y_REQUIRE(raw[l] != nullptr && hs[l] > 0 && ws[l] > 0, B2Y_ERR_INVALID, "gather_extras: level %d", l);
    p.raw[l] = reinterpret_cast<const float *>(raw[l]);
    p.hw[l] = hs[l] * ws[l];
    p.w[l] = ws[l];
    p.a0[l] = a0;
    p.stride[l] = strides[l];
    a0 += p.hw[l];
  }
  p.a0[n_levels] = a0;
  p.n_levels = n_levels; p.pitch = pitch; p.ne = ne; p.ndim = ndim; p.mode = mode; p.B = B; p.max_det = max_det;
  p.kept = reinterpret_cast<const int *>(kept_anchor);
  p.counts = reinterpret_cast<const int *>(counts);
  p.geom = reinterpret_cast<const float *>(geom);
  p.out = reinterpret_cast<float *>(out);
  p.round_dets = reinterpret_cast<float *>(round_dets);
  b2y_launch(gather_extras_kernel, dim3(ceil_div(max_det, 8), B), dim3(256), 0, (cudaStream_t)stream, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

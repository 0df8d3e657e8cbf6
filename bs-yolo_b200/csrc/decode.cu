// Detect._inference fused into one pass (ultralytics/nn/modules/head.py:100-131):
//   DFL softmax-expectation over 16 bins x 4 sides (nn/modules/block.py:73-76), make_anchors (+0.5 cell offsets,
//   row-major, levels concatenated; utils/tal.py:371-383), dist2bbox xywh (utils/tal.py:386-395), * stride,
//   class sigmoid, and the NHWC -> channel-major (B, 4+nc, A) transpose through shared memory.
// All math fp32 on fp32 head maps (decoding in bf16 moves boxes by up to ~4 px).
// block = 32 consecutive anchors of one level of one image; global reads and writes are 128-byte coalesced.
#include "common.cuh"

namespace {
constexpr int MAX_LEVELS = 4;
struct DecodeParams {
  const float *raw[MAX_LEVELS];
  int hw[MAX_LEVELS], w[MAX_LEVELS], blk0[MAX_LEVELS + 1], a0[MAX_LEVELS];
  float stride[MAX_LEVELS];
  int n_levels, nc, A;
  float *y;
};

__global__ void __launch_bounds__(128) decode_kernel(DecodeParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float tile[];  // [32][no+1] raw values, then [32][4] distances
  const int no = 64 + p.nc, pitch = no + 1;
  float *dist = tile + 32 * pitch;
  int lvl = 0;
  while (lvl + 1 < p.n_levels && (int)blockIdx.x >= p.blk0[lvl + 1]) ++lvl;
  const int b = blockIdx.y;
  const int first = ((int)blockIdx.x - p.blk0[lvl]) * 32;
  const int nval = min(32, p.hw[lvl] - first);
  const float *src = p.raw[lvl] + ((size_t)b * p.hw[lvl] + first) * no;
  for (int i = threadIdx.x; i < nval * no; i += blockDim.x) tile[(i / no) * pitch + (i % no)] = src[i];
  __syncthreads();
  {  // DFL: thread = (anchor, side)
    const int a = threadIdx.x >> 2, side = threadIdx.x & 3;
    if (a < nval) {
      const float *x = tile + a * pitch + side * 16;
      float m = x[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) m = fmaxf(m, x[k]);
      float e[16], s = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        e[k] = expf(x[k] - m);
        s += e[k];
      }
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) d = fmaf(e[k] / s, (float)k, d);
      dist[a * 4 + side] = d;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane >= nval) return;
  const int ai = first + lane;  // anchor index within the level
  float *yb = p.y + (size_t)b * (4 + p.nc) * p.A + p.a0[lvl] + ai;
  if (warp == 0) {
    const float ax = (float)(ai % p.w[lvl]) + 0.5f, ay = (float)(ai / p.w[lvl]) + 0.5f;
    const float x1 = ax - dist[lane * 4 + 0], y1 = ay - dist[lane * 4 + 1];
    const float x2 = ax + dist[lane * 4 + 2], y2 = ay + dist[lane * 4 + 3];
    const float st = p.stride[lvl];
    yb[0] = ((x1 + x2) / 2.f) * st;
    yb[(size_t)p.A] = ((y1 + y2) / 2.f) * st;
    yb[(size_t)2 * p.A] = (x2 - x1) * st;
    yb[(size_t)3 * p.A] = (y2 - y1) * st;
  }
  for (int c = warp; c < p.nc; c += 4) {
    const float v = tile[lane * pitch + 64 + c];
    yb[(size_t)(4 + c) * p.A] = 1.f / (1.f + expf(-v));
  }
}
}  // namespace

int decode_launch(const void *const *raw, const int32_t *hs, const int32_t *ws, const float *strides, int32_t n_levels,
                  int32_t B, int32_t nc, void *y, cudaStream_t st) {
  B2Y_REQUIRE(n_levels >= 1 && n_levels <= MAX_LEVELS && B > 0 && nc > 0 && y, B2Y_ERR_INVALID, "decode: bad arguments");
  DecodeParams p;
  int blk = 0, a0 = 0;
  for (int l = 0; l < n_levels; ++l) {
    B2Y_REQUIRE(raw[l] != nullptr, B2Y_ERR_INVALID, "decode: level %d pointer is null", l);
    p.raw[l] = reinterpret_cast<const float *>(raw[l]);
    p.hw[l] = hs[l] * ws[l];
    p.w[l] = ws[l];
    p.blk0[l] = blk;
    p.a0[l] = a0;
    p.stride[l] = strides[l];
    blk += ceil_div(p.hw[l], 32);
    a0 += p.hw[l];
  }
  p.blk0[n_levels] = blk;
  p.n_levels = n_levels;
  p.nc = nc;
  p.A = a0;
  p.y = reinterpret_cast<float *>(y);
  const size_t smem = (size_t)(32 * (64 + nc + 1) + 128) * sizeof(float);
  B2Y_REQUIRE(smem <= 200 * 1024, B2Y_ERR_UNSUPPORTED, "decode: nc=%d too large", nc);
  if (smem > 48 * 1024)
    B2Y_CUDA_CHECK(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  b2y_launch(decode_kernel, dim3(dim3(blk, B)), dim3(128), smem, st, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

extern "C" int b2y_decode(const void *const *raw, const int32_t *hs, const int32_t *ws, const float *strides,
                          int32_t n_levels, int32_t B, int32_t nc, void *y, void *stream) {
  B2YDeviceGuard guard(b2y_device_of(y));
  return decode_launch(raw, hs, ws, strides, n_levels, B, nc, y, (cudaStream_t)stream);
}

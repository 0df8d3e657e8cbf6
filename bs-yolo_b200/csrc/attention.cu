// C2PSA attention core (ultralytics/nn/modules/block.py:4280-4286):
//   per image b and head h:  S = (Q^T K) * kd^-0.5,  P = softmax(S, keys),  O = V P^T
// qkv is the output of the 1x1 `qkv` conv in NHWC: per pixel, head h owns channels
//   [h*(2kd+hd), +kd) = q, [.. +kd, +kd) = k, [.. +2kd, +hd) = v     (the .view(B, nh, 2kd+hd, N).split of :4280-4282)
// out channel h*hd + d = O[d] (the .view(B, C, H, W) of :4286).  The pe(v) depthwise conv and the residual are
// separate fused ops (the planner adds pe as a depthwise conv with this output as its residual).
//
// Flash-style: keys are streamed through shared memory in tiles with an online softmax, so the N x N score matrix
// (61 MB per image for yolo11x@1280) is never materialised.  fp32 math on CUDA cores.
#include "common.cuh"

namespace {
constexpr int QT = 32;   // queries per block
constexpr int KT = 64;   // keys per tile
constexpr int NT = 256;  // threads

template <int KD, int HD>
__global__ void __launch_bounds__(NT) attention_kernel(const bf16 *__restrict__ qkv, int cs, bf16 *__restrict__ out,
                                                       int out_cs, int N, float scale_log2e) {
  __shared__ float Qs[QT][KD + 1];
  __shared__ float Ks[KT][KD + 1];
  __shared__ float Vs[KT][HD];
  __shared__ float S[QT][KT + 1];
  __shared__ float row_m[QT], row_l[QT], row_alpha[QT];

  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bf16 *base = qkv + (size_t)b * N * cs + h * (2 * KD + HD);

  for (int i = tid; i < QT * KD; i += NT) {
    const int q = i / KD, d = i % KD;
    Qs[q][d] = (q0 + q < N) ? __bfloat162float(base[(size_t)(q0 + q) * cs + d]) * scale_log2e : 0.f;
  }
  if (tid < QT) {
    row_m[tid] = -INFINITY;
    row_l[tid] = 0.f;
  }
  const int oq = tid / 8, od = (tid % 8) * (HD / 8);
  float acc[HD / 8];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) acc[i] = 0.f;
  __syncthreads();

  for (int k0 = 0; k0 < N; k0 += KT) {
    for (int i = tid; i < KT * (KD / 8); i += NT) {
      const int j = i / (KD / 8), d = (i % (KD / 8)) * 8;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (k0 + j < N) unpack8(*reinterpret_cast<const uint4 *>(base + (size_t)(k0 + j) * cs + KD + d), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) Ks[j][d + t] = f[t];
    }
    for (int i = tid; i < KT * (HD / 8); i += NT) {
      const int j = i / (HD / 8), d = (i % (HD / 8)) * 8;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (k0 + j < N) unpack8(*reinterpret_cast<const uint4 *>(base + (size_t)(k0 + j) * cs + 2 * KD + d), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) Vs[j][d + t] = f[t];
    }
    __syncthreads();
    // scores (already in log2 units)
    {
      const int q = tid / 8;
#pragma unroll
      for (int i = 0; i < KT / 8; ++i) {
        const int j = (tid % 8) + 8 * i;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < KD; ++d) s = fmaf(Qs[q][d], Ks[j][d], s);
        S[q][j] = (k0 + j < N) ? s : -INFINITY;
      }
    }
    __syncthreads();
    // online softmax: warp w owns rows 4w..4w+3
#pragma unroll
    for (int r = 0; r < QT / (NT / 32); ++r) {
      const int q = warp * (QT / (NT / 32)) + r;
      const float s0 = S[q][lane], s1 = S[q][lane + 32];
      float m = fmaxf(s0, s1);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float m_old = row_m[q];
      const float m_new = fmaxf(m_old, m);
      const float p0 = exp2f(s0 - m_new), p1 = exp2f(s1 - m_new);
      float t = p0 + p1;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      S[q][lane] = p0;
      S[q][lane + 32] = p1;
      if (lane == 0) {
        const float alpha = exp2f(m_old - m_new);
        row_alpha[q] = alpha;
        row_l[q] = row_l[q] * alpha + t;
        row_m[q] = m_new;
      }
    }
    __syncthreads();
    {
      const float alpha = row_alpha[oq];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) acc[i] *= alpha;
      for (int j = 0; j < KT; ++j) {
        const float pj = S[oq][j];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) acc[i] = fmaf(pj, Vs[j][od + i], acc[i]);
      }
    }
    __syncthreads();
  }
  if (q0 + oq < N) {
    const float inv = 1.f / row_l[oq];
    bf16 *o = out + ((size_t)b * N + q0 + oq) * out_cs + h * HD + od;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i] = __float2bfloat16_rn(acc[i] * inv);
  }
}
}  // namespace

int launch_attention(const bf16 *qkv, int qkv_cs, bf16 *out, int out_cs, int B, int N, int nh, int kd, int hd,
                     cudaStream_t st) {
  B2Y_REQUIRE(kd == 32 && hd == 64, B2Y_ERR_UNSUPPORTED, "attention: only key_dim 32 / head_dim 64 (got %d/%d)", kd, hd);
  B2Y_REQUIRE(qkv_cs % 8 == 0 && ((uintptr_t)qkv & 15) == 0, B2Y_ERR_UNSUPPORTED, "attention: qkv must be 16B aligned");
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)kd);
  dim3 grid(ceil_div(N, QT), nh, B);
  attention_kernel<32, 64><<<grid, NT, 0, st>>>(qkv, qkv_cs, out, out_cs, N, scale_log2e);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

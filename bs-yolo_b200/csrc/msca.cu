// MSCAAttention up to its four branch maps in ONE kernel (ultralytics/nn/Addmodules/MSCA.py:52-74):
//
//     attn     = dw5x5(x)                                           conv0               (:54)
//     branch_j = dw(k_j x 1)(dw(1 x k_j)(attn)),  k = 5, 7, 11, 21  conv{j}_1, conv{j}_2 (:56-69; the shared per-channel
//                                                                   1x1 `dilconv` of branches 0..2 is folded into the
//                                                                   second strip's weights and bias by the planner)
//     gap_j    = mean over the map of branch_j                      self.gap            (:71-74)
//
// Everything here is depthwise, i.e. independent per channel, and the map of one image is small (20 x 20 at 640 px,
// 40 x 40 at 1280 px): one CTA owns the WHOLE map of one image for 8 channels, keeps x, attn and the strip intermediate
// in shared memory, and writes only the four branch maps (bf16) and their means.  Unfused this was ten launches (conv0,
// eight strip convs, the pooling) with nine round trips through HBM at 6-11 % of the roof; fused it reads x once and is
// bound by the fp32 FMA rate (113 MAC per pixel and channel), not by HBM.
// Work item = one channel pair x 10 consecutive outputs along the axis being convolved, the taps in registers
// (dw_regs.cuh): a 20 x 20 map is exactly 160 items per phase = one item per thread.  Lane order runs over (pair, the
// axis NOT blocked) and the plane pitches are padded to 4 mod 32 words, so every shared-memory access is conflict-free.
// Intermediates are rounded to bf16 exactly where the unfused path stored them; the means are summed in a fixed order.
#include <stdlib.h>

#include "dw_regs.cuh"

namespace {
constexpr int NT = 160;
constexpr int RO = 10;
constexpr int CP = 4;     // channel pairs (8 channels) per CTA
constexpr int HALO = 10;  // largest strip: 21 taps

struct MscaK {
  const bf16 *in; int in_cs;
  bf16 *out; int out_cs;
  float *gap;
  const float *params;
  int B, H, W, C, Hp, Wp, pitchX, pitchA, pitchU;
};

// taps of one channel pair from the CTA's parameter slice in shared memory ([122 rows][2*CP floats]); `row0` = first row
template <int K>
__device__ __forceinline__ void load_taps(const float *row0, float2 (&w)[K], float2 &bias) {
#pragma unroll
  for (int t = 0; t < K; ++t) w[t] = *reinterpret_cast<const float2 *>(row0 + t * 2 * CP);
  bias = *reinterpret_cast<const float2 *>(row0 + K * 2 * CP);
}

// one strip branch: 1 x K along W (attn -> U), K x 1 along H (U -> global + partial sums), mean.
// prm: first parameter row of the branch ([wh K | bh 1 | wv K | bv 1] rows of C floats), already offset to channel c0.
template <int K>
__device__ __forceinline__ void branch(const MscaK &p, int j, const float *prm, uint32_t A_s, uint32_t U_s, float2 *part,
                                       int b, int c0) {
  constexpr int PAD = K / 2;
  const int H = p.H, W = p.W;
  float2 w[K], bias;
  // ---- 1 x K along W: lanes over (pair, row), 10 output columns per item
  {
    const int ncb = p.Wp / RO;
    int cur = -1;
    for (int it = threadIdx.x; it < CP * H * ncb; it += NT) {
      const int cp = it % CP, row = (it / CP) % H, cb = it / (CP * H);
      if (cp != cur) { cur = cp; load_taps<K>(prm + cp * 2, w, bias); }
      float2 acc[RO];
#pragma unroll
      for (int o = 0; o < RO; ++o) acc[o] = bias;
      Dw1dImm<K, RO, CP * 4>::run(A_s + (uint32_t)(row * p.pitchA + (HALO + cb * RO - PAD) * CP + cp) * 4, w, acc);
      const uint32_t ua = U_s + (uint32_t)((HALO + row) * p.pitchU + cb * RO * CP + cp) * 4;
#pragma unroll
      for (int o = 0; o < RO; ++o)
        if (cb * RO + o < W) sts32(ua + o * CP * 4, f2_to_bf2(acc[o].x, acc[o].y));
    }
  }
  __syncthreads();
  // ---- K x 1 along H: lanes over (pair, column), 10 output rows per item
  const int nrb = p.Hp / RO;
  const int items = CP * W * nrb;
  {
    int cur = -1;
    for (int it = threadIdx.x; it < items; it += NT) {
      const int cp = it % CP, col = (it / CP) % W, rb = it / (CP * W);
      if (cp != cur) { cur = cp; load_taps<K>(prm + (K + 1) * 2 * CP + cp * 2, w, bias); }
      float2 acc[RO];
      dw1d<K, RO>(U_s + (uint32_t)((HALO + rb * RO - PAD) * p.pitchU + col * CP + cp) * 4, p.pitchU * 4, w, bias, acc);
      float2 s = make_float2(0.f, 0.f);
      bf16 *dst = p.out + ((size_t)(b * H + rb * RO) * W + col) * p.out_cs + (size_t)j * p.C + c0 + cp * 2;
      const size_t orow = (size_t)W * p.out_cs;
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        if (rb * RO + o < H) {
          const uint32_t pk = f2_to_bf2(acc[o].x, acc[o].y);
          const float2 r = bf2_to_f2(pk);  // the mean is taken over the STORED (rounded) map, like the pooling kernel did
          s.x += r.x; s.y += r.y;
          *reinterpret_cast<uint32_t *>(dst) = pk;
        }
        dst += orow;
      }
      part[it] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < CP * 2) {  // fixed summation order per channel
    const int cp = threadIdx.x >> 1, comp = threadIdx.x & 1;
    float s = 0.f;
    for (int i = cp; i < items; i += CP) s += reinterpret_cast<const float *>(part + i)[comp];
    p.gap[((size_t)b * 4 + j) * p.C + c0 + threadIdx.x] = s / (float)(H * W);
  }
  __syncthreads();
}

template <int MINB>
__global__ void __launch_bounds__(NT, MINB) msca_branch_kernel(MscaK p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  const int H = p.H, W = p.W, Hp = p.Hp;
  const int groups = p.C / (2 * CP);
  const int b = blockIdx.x / groups, c0 = (blockIdx.x % groups) * 2 * CP;
  // shared memory: X (zero halo 2) | A (attn, zero halo 10 left/right) | U (strip intermediate, zero halo 10 top/bottom) | sums
  uint32_t *X = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *A = X + (size_t)(Hp + 4) * p.pitchX;
  uint32_t *U = A + (size_t)H * p.pitchA;
  float2 *part = reinterpret_cast<float2 *>(U + (size_t)(Hp + 2 * HALO) * p.pitchU);
  float *wsm = reinterpret_cast<float *>(part + (size_t)CP * W * (Hp / RO));  // this CTA's 8 channels of all 122 parameter rows
  {  // parameters are constants: staged before the dependency wait, 32 bytes per row
    const uint32_t ws = (uint32_t)__cvta_generic_to_shared(wsm);
    for (int i = threadIdx.x; i < 122 * 2; i += NT)
      cp_async_16(ws + i * 16, p.params + (size_t)(i >> 1) * p.C + c0 + (i & 1) * 4, true);
  }
  {  // zero A and U once: the strip halos stay zero for all four branches (the convs' zero padding)
    const int n = H * p.pitchA + (Hp + 2 * HALO) * p.pitchU;
    for (int i = threadIdx.x; i < n; i += NT) A[i] = 0u;
  }
  pdl_wait();
  {  // stage x: 16 bytes (8 channels) per pixel, zero outside the map
    const uint32_t xs = (uint32_t)__cvta_generic_to_shared(X);
    const bf16 *img = p.in + (size_t)b * H * W * p.in_cs + c0;
    const int cols = W + 4;
    int row = threadIdx.x / cols, col = threadIdx.x % cols;
    const int drow = NT / cols, dcol = NT % cols;
    while (row < Hp + 4) {
      const int gy = row - 2, gx = col - 2;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      cp_async_16(xs + (uint32_t)(row * p.pitchX + col * CP) * 4, img + ((size_t)(ok ? gy : 0) * W + (ok ? gx : 0)) * p.in_cs, ok);
      row += drow;
      col += dcol;
      if (col >= cols) { col -= cols; ++row; }
    }
    cp_async_commit_wait();
  }
  __syncthreads();
  const float *prm = wsm;
  const uint32_t X_s = (uint32_t)__cvta_generic_to_shared(X), A_s = (uint32_t)__cvta_generic_to_shared(A);
  const uint32_t U_s = (uint32_t)__cvta_generic_to_shared(U);
  // ---- conv0: 5x5 -> attn (A, bf16) at column offset HALO; lanes over (pair, column), 10 output rows per item
  {
    float2 w[25], bias = make_float2(0.f, 0.f);
    int cur = -1;
    for (int it = threadIdx.x; it < CP * W * (Hp / RO); it += NT) {
      const int cp = it % CP, col = (it / CP) % W, rb = it / (CP * W);
      if (cp != cur) { cur = cp; load_taps<25>(prm + cp * 2, w, bias); }
      float2 acc[RO];
#pragma unroll
      for (int o = 0; o < RO; ++o) acc[o] = bias;
      dw2d_accum<5, 0, 5, RO, CP * 4>(X_s + (uint32_t)(rb * RO * p.pitchX + col * CP + cp) * 4, p.pitchX * 4, w, acc);
      uint32_t aa = A_s + (uint32_t)(rb * RO * p.pitchA + (HALO + col) * CP + cp) * 4;
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        if (rb * RO + o < H) sts32(aa, f2_to_bf2(acc[o].x, acc[o].y));
        aa += p.pitchA * 4;
      }
    }
  }
  __syncthreads();
  // ---- the four strip branches; packed rows: [w0 25 | b0 1 | per branch: wh K | bh 1 | wv K | bv 1]
  prm += 26 * 2 * CP;
  branch<5>(p, 0, prm, A_s, U_s, part, b, c0);
  prm += 12 * 2 * CP;
  branch<7>(p, 1, prm, A_s, U_s, part, b, c0);
  prm += 16 * 2 * CP;
  branch<11>(p, 2, prm, A_s, U_s, part, b, c0);
  prm += 24 * 2 * CP;
  branch<21>(p, 3, prm, A_s, U_s, part, b, c0);
}

inline int pad_pitch(int words) {  // smallest pitch >= words with pitch % 32 == CP: row r starts CP banks after row r-1
  int p = words;
  while (p % 32 != CP) ++p;
  return p;
}
}  // namespace

int launch_msca_branch(const MscaParams &q, cudaStream_t st) {
  B2Y_REQUIRE(q.C % 8 == 0 && q.in_cs % 8 == 0 && q.out_cs % 2 == 0 && (reinterpret_cast<uintptr_t>(q.in) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(q.out) & 3) == 0 && (reinterpret_cast<uintptr_t>(q.params) & 15) == 0,
              B2Y_ERR_UNSUPPORTED, "msca_branch: C must be a multiple of 8 and the input view 16-byte aligned");
  MscaK p;
  p.in = q.in; p.in_cs = q.in_cs; p.out = q.out; p.out_cs = q.out_cs; p.gap = q.gap; p.params = q.params;
  p.B = q.B; p.H = q.H; p.W = q.W; p.C = q.C;
  p.Hp = (q.H + RO - 1) / RO * RO;
  p.Wp = (q.W + RO - 1) / RO * RO;
  p.pitchX = (q.W + 4) * CP;
  p.pitchA = pad_pitch((p.Wp + 2 * HALO) * CP);
  p.pitchU = pad_pitch(q.W * CP);
  const int items = CP * q.W * (p.Hp / RO);
  const size_t smem = ((size_t)(p.Hp + 4) * p.pitchX + (size_t)q.H * p.pitchA + (size_t)(p.Hp + 2 * HALO) * p.pitchU) * 4 +
                      (size_t)items * 8 + 122 * 2 * CP * 4;
  B2Y_REQUIRE(smem <= 200 * 1024, B2Y_ERR_UNSUPPORTED, "msca_branch: a %dx%d map does not fit shared memory", q.H, q.W);
  static const int minb = [] { const char *e = getenv("B2Y_MSCA_MINB"); return e ? atoi(e) : 4; }();  // A/B knob: 3 (128 regs) | 4 (96 regs, measured 7 % faster)
  auto go = [&](auto kern) -> int {
    B2Y_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b2y_launch(kern, dim3((unsigned)(q.B * (q.C / (2 * CP)))), dim3(NT), smem, st, p);
    B2Y_LAUNCH_CHECK();
    return B2Y_OK;
  };
  return minb == 4 ? go(msca_branch_kernel<4>) : go(msca_branch_kernel<3>);
}

// MSCAAttention up to its four branch maps in ONE kernel (ultralytics/nn/Addmodules/MSCA.py:52-74):
//
//     attn     = dw5x5(x)                                           conv0               (:54)
//     branch_j = dw(k_j x 1)(dw(1 x k_j)(attn)),  k = 5, 7, 11, 21  conv{j}_1, conv{j}_2 (:56-69; the shared per-channel
//                                                                   1x1 `dilconv` of branches 0..2 is folded into the
//                                                                   second strip's weights and bias by the planner)
//     gap_j    = mean over the map of branch_j                      self.gap            (:71-74)
//
// Everything here is depthwise, i.e. independent per channel, and the map of one image is small (20 x 20 at 640 px,
// 40 x 40 at 1280 px): one CTA owns the WHOLE map of one image for a group of 4*CQ channels, keeps x, attn and the strip
// intermediate in shared memory, and writes only the four branch maps (bf16) and their means.  Unfused this was ten
// launches (conv0, eight strip convs, the pooling) with nine round trips through HBM at 6-11 % of the roof; fused it reads x
// once and is bound by the fp32 FMA rate (113 MAC per pixel and channel), not by HBM.
// Work item = one channel quad x 4 outputs along the axis being convolved, so a loaded input feeds 4 outputs and a weight
// is fetched once per input position (float4 broadcast).  Lane order always runs over (quad, the axis NOT blocked), and the
// plane pitches are padded to CQ mod 16 quads, so every shared-memory access of a warp is conflict-free.  Intermediates
// are rounded to bf16 exactly where the unfused path stored them; the means are summed in a fixed order (deterministic).
#include "common.cuh"

namespace {
constexpr int NT = 256;
constexpr int RO = 4;
constexpr int HALO = 10;  // largest strip: 21 taps

struct MscaK {
  const bf16 *in; int in_cs;
  bf16 *out; int out_cs;
  float *gap;
  const float *params;
  int B, H, W, C, CQ, Hp, pitchX, pitchA, pitchU;
};

__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// K-tap 1-D convolution of RO consecutive outputs along an axis whose positions are `stride` quads apart; src points at
// the first tap of output 0.  w: K float4 weights `nq` float4s apart.
template <int K>
__device__ __forceinline__ void conv1d_block(const uint2 *src, int stride, const float4 *w, int nq, float4 bias,
                                             float4 (&acc)[RO]) {
#pragma unroll
  for (int o = 0; o < RO; ++o) acc[o] = bias;
#pragma unroll
  for (int j = 0; j < K + RO - 1; ++j) {
    float x[4];
    unpack4(src[(size_t)j * stride], x);
#pragma unroll
    for (int o = 0; o < RO; ++o) {
      const int r = j - o;
      if (r >= 0 && r < K) {
        const float4 ww = w[r * nq];
        acc[o].x = fmaf(x[0], ww.x, acc[o].x);
        acc[o].y = fmaf(x[1], ww.y, acc[o].y);
        acc[o].z = fmaf(x[2], ww.z, acc[o].z);
        acc[o].w = fmaf(x[3], ww.w, acc[o].w);
      }
    }
  }
}

template <int K>
__device__ __forceinline__ void branch(const MscaK &p, int j, const float4 *wh, const float4 *wv, uint2 *A, uint2 *U,
                                       float4 *part, int b, int c0) {
  constexpr int PAD = K / 2;
  const int CQ = p.CQ, H = p.H, W = p.W;
  const float4 *bh = wh + K * CQ, *bv = wv + K * CQ;
  // ---- 1 x K along W: attn -> U (bf16), rows [HALO, HALO + H)
  const int ncb = (W + RO - 1) / RO;
  for (int it = threadIdx.x; it < CQ * H * ncb; it += NT) {
    const int q = it % CQ, row = (it / CQ) % H, cb = it / (CQ * H);
    float4 acc[RO];
    conv1d_block<K>(A + (size_t)row * p.pitchA + (HALO + cb * RO - PAD) * CQ + q, CQ, wh + q, CQ, bh[q], acc);
#pragma unroll
    for (int o = 0; o < RO; ++o) {
      const int col = cb * RO + o;
      if (col < W) {
        const float v[4] = {acc[o].x, acc[o].y, acc[o].z, acc[o].w};
        U[(size_t)(HALO + row) * p.pitchU + col * CQ + q] = pack4(v);
      }
    }
  }
  __syncthreads();
  // ---- K x 1 along H: U -> branch map (global, bf16) + partial sums of the rounded outputs
  const int nrb = p.Hp / RO;
  const int items = CQ * W * nrb;
  for (int it = threadIdx.x; it < items; it += NT) {
    const int q = it % CQ, w = (it / CQ) % W, rb = it / (CQ * W);
    float4 acc[RO];
    conv1d_block<K>(U + (size_t)(HALO + rb * RO - PAD) * p.pitchU + w * CQ + q, p.pitchU, wv + q, CQ, bv[q], acc);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int o = 0; o < RO; ++o) {
      const int row = rb * RO + o;
      if (row < H) {
        const float v[4] = {acc[o].x, acc[o].y, acc[o].z, acc[o].w};
        const uint2 pk = pack4(v);
        float r[4];
        unpack4(pk, r);
        s.x += r[0]; s.y += r[1]; s.z += r[2]; s.w += r[3];
        *reinterpret_cast<uint2 *>(p.out + ((size_t)(b * H + row) * W + w) * p.out_cs + (size_t)j * p.C + c0 + q * 4) = pk;
      }
    }
    part[it] = s;
  }
  __syncthreads();
  // ---- mean of the branch map: fixed summation order per channel
  if (threadIdx.x < CQ * 4) {
    const int q = threadIdx.x >> 2, comp = threadIdx.x & 3;
    float s = 0.f;
    for (int i = q; i < items; i += CQ) s += reinterpret_cast<const float *>(part + i)[comp];
    p.gap[((size_t)b * 4 + j) * p.C + c0 + threadIdx.x] = s / (float)(H * W);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT) msca_branch_kernel(MscaK p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  const int CQ = p.CQ, H = p.H, W = p.W, Hp = p.Hp;
  const int groups = p.C / (4 * CQ);
  const int b = blockIdx.x / groups, c0 = (blockIdx.x % groups) * 4 * CQ;
  // shared memory: weights (122 float4 per quad) | X | A | U | partial sums
  float4 *wsm = reinterpret_cast<float4 *>(smem_raw);
  uint2 *X = reinterpret_cast<uint2 *>(wsm + 122 * CQ);
  uint2 *A = X + (size_t)(Hp + 4) * p.pitchX;
  uint2 *U = A + (size_t)Hp * p.pitchA;
  float4 *part = reinterpret_cast<float4 *>(U + (size_t)(Hp + 2 * HALO) * p.pitchU);
  // parameters: row r of the packed blob is C floats; this CTA takes its 4*CQ channels of every row
  for (int i = threadIdx.x; i < 122 * CQ; i += NT) {
    const int row = i / CQ, q = i % CQ;
    wsm[i] = *reinterpret_cast<const float4 *>(p.params + (size_t)row * p.C + c0 + q * 4);
  }
  // zero A and U once: the strip halos stay zero for all four branches (the convs' zero padding)
  {
    uint2 *z = A;
    const int n = Hp * p.pitchA + (Hp + 2 * HALO) * p.pitchU;
    for (int i = threadIdx.x; i < n; i += NT) z[i] = make_uint2(0u, 0u);
  }
  pdl_wait();
  {
    const uint32_t xs = (uint32_t)__cvta_generic_to_shared(X);
    const bf16 *img = p.in + (size_t)b * H * W * p.in_cs + c0;
    const int n = (Hp + 4) * p.pitchX;
    for (int i = threadIdx.x; i < n; i += NT) {
      const int q = i % CQ, col = (i / CQ) % (W + 4), row = i / p.pitchX;
      const int gy = row - 2, gx = col - 2;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      cp_async8(xs + i * 8, img + ((size_t)(ok ? gy : 0) * W + (ok ? gx : 0)) * p.in_cs + q * 4, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  // ---- conv0: 5x5 -> attn (A, bf16) at column offset HALO
  {
    const float4 *w0 = wsm, *b0 = wsm + 25 * CQ;
    for (int it = threadIdx.x; it < CQ * W * (Hp / RO); it += NT) {
      const int q = it % CQ, w = (it / CQ) % W, rb = it / (CQ * W);
      float4 acc[RO];
#pragma unroll
      for (int o = 0; o < RO; ++o) acc[o] = b0[q];
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        float x[RO + 4][4];
#pragma unroll
        for (int i = 0; i < RO + 4; ++i) unpack4(X[(size_t)(rb * RO + i) * p.pitchX + (w + s) * CQ + q], x[i]);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const float4 ww = w0[(r * 5 + s) * CQ + q];
#pragma unroll
          for (int o = 0; o < RO; ++o) {
            acc[o].x = fmaf(x[o + r][0], ww.x, acc[o].x);
            acc[o].y = fmaf(x[o + r][1], ww.y, acc[o].y);
            acc[o].z = fmaf(x[o + r][2], ww.z, acc[o].z);
            acc[o].w = fmaf(x[o + r][3], ww.w, acc[o].w);
          }
        }
      }
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        const int row = rb * RO + o;
        if (row < H) {
          const float v[4] = {acc[o].x, acc[o].y, acc[o].z, acc[o].w};
          A[(size_t)row * p.pitchA + (HALO + w) * CQ + q] = pack4(v);
        }
      }
    }
  }
  __syncthreads();
  // ---- the four strip branches; packed rows: [w0 25 | b0 1 | per branch: wh K | bh 1 | wv K | bv 1]
  const float4 *pw = wsm + 26 * CQ;
  branch<5>(p, 0, pw, pw + 6 * CQ, A, U, part, b, c0);
  pw += 12 * CQ;
  branch<7>(p, 1, pw, pw + 8 * CQ, A, U, part, b, c0);
  pw += 16 * CQ;
  branch<11>(p, 2, pw, pw + 12 * CQ, A, U, part, b, c0);
  pw += 24 * CQ;
  branch<21>(p, 3, pw, pw + 22 * CQ, A, U, part, b, c0);
}

inline int pad_pitch(int quads, int CQ) {  // smallest pitch >= quads with pitch % 16 == CQ % 16
  int p = quads;
  while (p % 16 != CQ % 16) ++p;
  return p;
}
}  // namespace

int launch_msca_branch(const MscaParams &q, cudaStream_t st) {
  B2Y_REQUIRE(q.C % 8 == 0 && q.in_cs % 4 == 0 && q.out_cs % 4 == 0 && (reinterpret_cast<uintptr_t>(q.in) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(q.out) & 7) == 0 && (reinterpret_cast<uintptr_t>(q.params) & 15) == 0,
              B2Y_ERR_UNSUPPORTED, "msca_branch: C must be a multiple of 8 and the views 8-byte aligned");
  MscaK p;
  p.in = q.in; p.in_cs = q.in_cs; p.out = q.out; p.out_cs = q.out_cs; p.gap = q.gap; p.params = q.params;
  p.B = q.B; p.H = q.H; p.W = q.W; p.C = q.C;
  p.CQ = 2;
  p.Hp = (q.H + RO - 1) / RO * RO;
  p.pitchX = (q.W + 4) * p.CQ;
  p.pitchA = pad_pitch((q.W + 2 * HALO + RO) * p.CQ, p.CQ);
  p.pitchU = pad_pitch(q.W * p.CQ, p.CQ);
  const int items = p.CQ * q.W * (p.Hp / RO);
  const size_t smem = (size_t)122 * p.CQ * 16 +
                      ((size_t)(p.Hp + 4) * p.pitchX + (size_t)p.Hp * p.pitchA + (size_t)(p.Hp + 2 * HALO) * p.pitchU) * 8 +
                      (size_t)items * 16;
  B2Y_REQUIRE(smem <= 200 * 1024, B2Y_ERR_UNSUPPORTED, "msca_branch: a %dx%d map does not fit shared memory", q.H, q.W);
  B2Y_CUDA_CHECK(cudaFuncSetAttribute(msca_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  b2y_launch(msca_branch_kernel, dim3((unsigned)(q.B * (q.C / (4 * p.CQ)))), dim3(NT), smem, st, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

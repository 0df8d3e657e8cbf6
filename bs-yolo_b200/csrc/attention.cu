// C2PSA attention core (ultralytics/nn/modules/block.py:4280-4286):
//   per image b and head h:  S = (Q^T K) * kd^-0.5,  P = softmax(S, keys),  O = V P^T
// qkv is the output of the 1x1 `qkv` conv in NHWC: per pixel, head h owns channels
//   [h*(2kd+hd), +kd) = q, [.. +kd, +kd) = k, [.. +2kd, +hd) = v     (the .view(B, nh, 2kd+hd, N).split of :4280-4282)
// out channel h*hd + d = O[d] (the .view(B, C, H, W) of :4286).  The pe(v) depthwise conv and the residual are
// separate fused ops (the planner adds pe as a depthwise conv with this output as its residual).
//
// Flash-style: one CTA = 64 queries of one (image, head); keys/values stream through shared memory in 64-key tiles
// (cp.async double buffering) with an online softmax, so the N x N score matrix (61 MB per image for yolo11x@1280) is
// never materialised.  Both GEMMs run on the tensor cores through warp-level mma.sync (m16n8k16 bf16, fp32 accumulate):
// at N = 400..1600 keys and d = 32/64 this op is 1 % of the network's FLOPs, far below what would amortise a
// tcgen05/TMEM pipeline; softmax statistics stay in fp32 registers.
#include "common.cuh"

namespace {
constexpr int KD = 32, HD = 64;
constexpr int QT = 64;              // queries per CTA (4 warps x 16 rows)
constexpr int KT = 64;              // keys per tile
constexpr int NT = 128;
constexpr int KPITCH = 80;          // bytes per K row in smem (64 + 16 pad: conflict-free fragment loads)
constexpr int VPITCH = 144;         // bytes per V row in smem (128 + 16 pad: conflict-free ldmatrix)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void mma_bf16(float *c, const uint32_t *a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

__global__ void __launch_bounds__(NT) attention_kernel(const bf16 *__restrict__ qkv, int cs, bf16 *__restrict__ out,
                                                       int out_cs, int N, float scale_log2e) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) uint8_t Ksm[2][KT * KPITCH];
  __shared__ __align__(16) uint8_t Vsm[2][KT * VPITCH];
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const bf16 *base = qkv + (size_t)b * N * cs + h * (2 * KD + HD);

  auto load_tile = [&](int buf, int k0) {
    const uint32_t ks = (uint32_t)__cvta_generic_to_shared(Ksm[buf]), vs = (uint32_t)__cvta_generic_to_shared(Vsm[buf]);
    for (int i = tid; i < KT * 4; i += NT) {  // K: 4 x 16 B per key
      const int j = i >> 2, c = i & 3;
      const bool ok = k0 + j < N;
      cp_async16(ks + j * KPITCH + c * 16, base + (size_t)(ok ? k0 + j : 0) * cs + KD + c * 8, ok);
    }
    for (int i = tid; i < KT * 8; i += NT) {  // V: 8 x 16 B per key
      const int j = i >> 3, c = i & 7;
      const bool ok = k0 + j < N;
      cp_async16(vs + j * VPITCH + c * 16, base + (size_t)(ok ? k0 + j : 0) * cs + 2 * KD + c * 8, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // Q fragments (A operand, 16 x 32 per warp): rows g and g+8 of the warp's 16 queries
  uint32_t qf[2][4];
  {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = ks * 16 + 2 * t;
      qf[ks][0] = r0 < N ? *reinterpret_cast<const uint32_t *>(base + (size_t)r0 * cs + c) : 0u;
      qf[ks][1] = r1 < N ? *reinterpret_cast<const uint32_t *>(base + (size_t)r1 * cs + c) : 0u;
      qf[ks][2] = r0 < N ? *reinterpret_cast<const uint32_t *>(base + (size_t)r0 * cs + c + 8) : 0u;
      qf[ks][3] = r1 < N ? *reinterpret_cast<const uint32_t *>(base + (size_t)r1 * cs + c + 8) : 0u;
    }
  }
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int ntiles = (N + KT - 1) / KT;
  load_tile(0, 0);
  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1, k0 = it * KT;
    if (it + 1 < ntiles) {
      load_tile(buf ^ 1, k0 + KT);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint8_t *Kb = Ksm[buf];
    // ---- S = Q K^T (16 x 64 per warp), in log2 units
    float s[KT / 8][4];
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint8_t *kp = Kb + (nt * 8 + g) * KPITCH + (ks * 16 + 2 * t) * 2;
        mma_bf16(s[nt], qf[ks], *reinterpret_cast<const uint32_t *>(kp), *reinterpret_cast<const uint32_t *>(kp + 16));
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      const int key = k0 + nt * 8 + 2 * t;
      s[nt][0] = key < N ? s[nt][0] * scale_log2e : -INFINITY;
      s[nt][1] = key + 1 < N ? s[nt][1] * scale_log2e : -INFINITY;
      s[nt][2] = key < N ? s[nt][2] * scale_log2e : -INFINITY;
      s[nt][3] = key + 1 < N ? s[nt][3] * scale_log2e : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile has at least one valid key
    const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[KT / 16][4];  // P as A operand fragments (16 x 64)
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - mn0), p1 = exp2f(s[nt][1] - mn0);
      const float p2 = exp2f(s[nt][2] - mn1), p3 = exp2f(s[nt][3] - mn1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1;
    }
    // ---- O += P V : V^T fragments via ldmatrix.trans on the [key][d] tile
    const uint32_t vb = (uint32_t)__cvta_generic_to_shared(Vsm[buf]);
#pragma unroll
    for (int kk = 0; kk < KT / 16; ++kk) {
#pragma unroll
      for (int dn = 0; dn < HD / 16; ++dn) {  // two 8-wide d tiles per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = vb + (uint32_t)((kk * 16 + (lane & 15)) * VPITCH + (dn * 16 + (lane >> 4) * 8) * 2);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(addr));
        mma_bf16(o[dn * 2], pf[kk], b0, b1);
        mma_bf16(o[dn * 2 + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // the next iteration's prefetch overwrites the other buffer only after everyone left this one
  }
  // row sums live distributed over the 4 lanes of a quad
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const int d = h * HD + i * 8 + 2 * t;
    if (r0 < N) *reinterpret_cast<uint32_t *>(out + ((size_t)b * N + r0) * out_cs + d) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
    if (r1 < N) *reinterpret_cast<uint32_t *>(out + ((size_t)b * N + r1) * out_cs + d) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
  }
}
}  // namespace

int launch_attention(const bf16 *qkv, int qkv_cs, bf16 *out, int out_cs, int B, int N, int nh, int kd, int hd,
                     cudaStream_t st) {
  B2Y_REQUIRE(kd == KD && hd == HD, B2Y_ERR_UNSUPPORTED, "attention: only key_dim 32 / head_dim 64 (got %d/%d)", kd, hd);
  B2Y_REQUIRE(qkv_cs % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && out_cs % 2 == 0 && ((uintptr_t)out & 3) == 0,
              B2Y_ERR_UNSUPPORTED, "attention: qkv must be 16-byte aligned, out 4-byte aligned");
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)kd);
  dim3 grid(ceil_div(N, QT), nh, B);
  b2y_launch(attention_kernel, dim3(grid), dim3(NT), 0, st, qkv, qkv_cs, out, out_cs, N, scale_log2e);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

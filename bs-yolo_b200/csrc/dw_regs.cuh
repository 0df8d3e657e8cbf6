// Register-blocked depthwise convolution from shared memory (pmsfa_dw.cu, msca.cu).
//
// A thread owns ONE channel pair (a 32-bit bf16x2 word per pixel) and RO consecutive outputs along one axis.  Its taps'
// weights sit in registers for the whole phase, so the only shared-memory traffic is the input words, each feeding up to
// min(RO, K) outputs: <= 0.2 LDS per FMA.  (The first version kept 4 channels per thread and fetched the weights as float4
// broadcasts from shared memory: ncu showed 25 % of the issued instructions were FMAs and the FMA pipe at 27 %,
// profiles/r02_ncu_bsyolo_simt.md.)  Lanes run over (channel pair, the axis that is NOT blocked), so a warp's words are
// contiguous in the [row][col][pair] planes and every access is conflict-free.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// Shared memory is addressed with 32-bit shared-space byte addresses and ld.shared with immediate offsets: with generic
// pointers and runtime pitches the compiler spent six integer instructions per load on 64-bit address arithmetic (ncu, v3:
// 35 % FFMA, 27 % IADD3/LEA/IMAD).
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds32_off(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// K x K taps, kernel rows [R0, R1), RO vertically adjacent outputs, input-stationary: every input word is loaded once and
// immediately applied to all the outputs it contributes to (no input array).  `saddr`: shared byte address of plane word
// (row r0, col w, pair cp), (r0, w) = the top-left tap of output 0; `rs`: bytes per plane row; CSB: bytes between
// columns (compile time, so the K loads of a row use immediate offsets).  wreg[(r - R0) * K + s]: the taps (fp32 pairs).
template <int K, int R0, int R1, int RO, int CSB>
__device__ __forceinline__ void dw2d_accum(uint32_t saddr, int rs, const float2 *wreg, float2 (&acc)[RO]) {
  constexpr int KR = R1 - R0;
  uint32_t ra = saddr + R0 * rs;
#pragma unroll
  for (int i = 0; i < RO + KR - 1; ++i) {
    uint32_t xs[K];
    xs[0] = lds32(ra);
    if (K > 1) xs[1 % K] = lds32_off<1 * CSB>(ra);
    if (K > 2) xs[2 % K] = lds32_off<2 * CSB>(ra);
    if (K > 3) xs[3 % K] = lds32_off<3 * CSB>(ra);
    if (K > 4) xs[4 % K] = lds32_off<4 * CSB>(ra);
    if (K > 5) xs[5 % K] = lds32_off<5 * CSB>(ra);
    if (K > 6) xs[6 % K] = lds32_off<6 * CSB>(ra);
    static_assert(K <= 7, "dw2d_accum: K <= 7");
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const float2 x = bf2_to_f2(xs[s]);
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        const int r = i - o;
        if (r >= 0 && r < KR) {
          acc[o].x = fmaf(x.x, wreg[r * K + s].x, acc[o].x);
          acc[o].y = fmaf(x.y, wreg[r * K + s].y, acc[o].y);
        }
      }
    }
    ra += rs;
  }
}

// K taps along one axis, RO consecutive outputs; `saddr` = shared byte address of the first tap of output 0, `st` bytes
// between successive positions (runtime: one add per load).  Inputs are consumed as they arrive; K tap pairs in registers.
template <int K, int RO>
__device__ __forceinline__ void dw1d(uint32_t saddr, int st, const float2 *wreg, float2 bias, float2 (&acc)[RO]) {
#pragma unroll
  for (int o = 0; o < RO; ++o) acc[o] = bias;
#pragma unroll
  for (int j = 0; j < K + RO - 1; ++j) {
    const float2 x = bf2_to_f2(lds32(saddr));
    saddr += st;
#pragma unroll
    for (int o = 0; o < RO; ++o) {
      const int r = j - o;
      if (r >= 0 && r < K) {
        acc[o].x = fmaf(x.x, wreg[r].x, acc[o].x);
        acc[o].y = fmaf(x.y, wreg[r].y, acc[o].y);
      }
    }
  }
}
// same with a compile-time stride STB (bytes): immediate offsets, no address arithmetic at all
template <int K, int RO, int STB, int J = 0>
struct Dw1dImm {
  static __device__ __forceinline__ void run(uint32_t saddr, const float2 *wreg, float2 (&acc)[RO]) {
    const float2 x = bf2_to_f2(lds32_off<J * STB>(saddr));
#pragma unroll
    for (int o = 0; o < RO; ++o) {
      const int r = J - o;
      if (r >= 0 && r < K) {
        acc[o].x = fmaf(x.x, wreg[r].x, acc[o].x);
        acc[o].y = fmaf(x.y, wreg[r].y, acc[o].y);
      }
    }
    Dw1dImm<K, RO, STB, J + 1>::run(saddr, wreg, acc);
  }
};
template <int K, int RO, int STB>
struct Dw1dImm<K, RO, STB, K + RO - 1> {
  static __device__ __forceinline__ void run(uint32_t, const float2 *, float2 (&)[RO]) {}
};

__device__ __forceinline__ float silu_tanh_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ void cp_async_8(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 8 : 0;  // src-size 0: the 8 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

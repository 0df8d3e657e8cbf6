// Fused depthwise tail of a PMSFA block (ultralytics/nn/modules/block.py:3035-3053), one kernel:
//
//     conv1_out = [a | b | h]        (c channels: quarters a, b of c/4 and the half h of c/2; written by conv1, tensor cores)
//     t   = SiLU(dw5x5([a | b]))      conv2 on the first half              (block.py:3047)
//     out = [ SiLU(dw7x7(t_a)) | t_b | h ]                                  (block.py:3049-3051: cat([conv3_out, conv2_out_2,
//                                                                            conv1_out_2]) -- the input of conv4)
//
// Unfused this is 3 launches and 2 round trips through HBM (t, the cat copy) at 6-11 % of the HBM roof.
// Here one CTA owns a TH x TW pixel tile:
//   * cp.async stages quarter a with a 5-pixel halo and quarter b with a 2-pixel halo into shared memory as two dense
//     planes [row][col][channel pair] (zero-filled outside the map = the convs' zero padding);
//   * conv2 runs from shared memory: on a over the tile + 3-pixel halo (what conv3 needs) -> bf16 plane t_a in shared
//     memory (zero outside the map: conv3 pads its INPUT with zeros), on b over the tile -> global;
//   * conv3 runs over t_a -> global; the half h is passed through with 16-byte copies.
// Work item = one channel pair x 1 column x 8 output rows with the taps in registers (dw_regs.cuh); the 7x7 taps are
// applied in two row passes (4 + 3 rows) so that at most 56 weight registers are live.
// The layer is CUDA-core bound (396 MAC per pixel at c = 16 against 64 B of traffic), not HBM bound: the bound quoted
// in DESIGN.md is the fp32 FMA rate.  t is rounded to bf16 exactly where the unfused path stored it.
#include <stdlib.h>

#include "dw_regs.cuh"

namespace {
constexpr int NT = 128;
constexpr int RO = 8;  // output rows per work item

struct PmsfaK {
  const bf16 *in; int in_cs;
  bf16 *out; int out_cs;
  const float *w2, *b2, *w3, *b3;  // [25][c/2], [c/2], [49][c/4], [c/4]
  int B, H, W, c;
  int TH, TW, tiles_x, tiles_y;
  int n8_log2;  // log2 of the 8-byte pieces per pixel of one quarter (c / 16)
};

// stage a (rows x cols) pixel window of `n8` 8-byte pieces per pixel, top-left at (gy0, gx0), zero outside the map
__device__ __forceinline__ void stage_plane(uint32_t dst_s, const bf16 *img, int in_cs, int ch0, int rows, int cols, int valid_rows,
                                            int gy0, int gx0, int H, int W, int n8_log2) {
  const int epr = cols << n8_log2;  // 8-byte elements per plane row
  int row = threadIdx.x / epr, rem = threadIdx.x % epr;
  const int drow = NT / epr, drem = NT % epr;
  const int n8m = (1 << n8_log2) - 1;
  while (row < rows) {
    const int col = rem >> n8_log2, piece = rem & n8m;
    const int gy = gy0 + row, gx = gx0 + col;
    const bool ok = row < valid_rows && gy >= 0 && gy < H && gx >= 0 && gx < W;
    cp_async_8(dst_s + (uint32_t)(row * epr + rem) * 8, img + ((size_t)(ok ? gy : 0) * W + (ok ? gx : 0)) * in_cs + ch0 + piece * 4, ok);
    row += drow;
    rem += drem;
    if (rem >= epr) { rem -= epr; ++row; }
  }
}

template <int CPQ, int MINB>  // channel pairs per quarter (c / 8); CTAs per SM the register budget is sized for
__global__ void __launch_bounds__(NT, MINB) pmsfa_dw_kernel(PmsfaK p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  constexpr int cpq = CPQ, Q = 2 * CPQ;        // channel pairs / channels per quarter
  const int TH = p.TH, TW = p.TW;
  const int AH = TH + 10, AW = TW + 10;        // plane a: tile + 5-pixel halo (2 for conv2 + 3 for conv3)
  const int BH = TH + 4, BW = TW + 4;          // plane b: tile + 2-pixel halo
  const int MW = TW + 6;                       // t_a: tile + 3-pixel halo
  const int MHp = (TH + 6 + RO - 1) / RO * RO; // ... rows padded to whole row blocks
  uint32_t *pa = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *pb = pa + (size_t)(MHp + 4) * AW * cpq;
  uint32_t *tm = pb + (size_t)BH * BW * cpq;

  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, b = tile / (p.tiles_x * p.tiles_y);
  const int y0 = ty * TH, x0 = tx * TW;
  const bf16 *img = p.in + (size_t)b * p.H * p.W * p.in_cs;
  bf16 *oimg = p.out + (size_t)b * p.H * p.W * p.out_cs;
  pdl_wait();  // the activations come from the preceding kernel (conv1)

  stage_plane((uint32_t)__cvta_generic_to_shared(pa), img, p.in_cs, 0, MHp + 4, AW, AH, y0 - 5, x0 - 5, p.H, p.W, p.n8_log2);
  stage_plane((uint32_t)__cvta_generic_to_shared(pb), img, p.in_cs, Q, BH, BW, BH, y0 - 2, x0 - 2, p.H, p.W, p.n8_log2);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // ---- pass the half h through while the planes land: 16-byte vectors, c/2 channels = c bytes per pixel; four loads are
  //      issued before their stores so that a warp waits for memory once per four vectors, not once per vector
  {
    const int vl2 = p.n8_log2;  // c / 16 vectors per pixel
    const int epr = TW << vl2;
    int row = threadIdx.x / epr, rem = threadIdx.x % epr;
    const int drow = NT / epr, drem = NT % epr;
    while (row < TH) {
      uint4 val[4];
      size_t dst[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int col = rem >> vl2, v = rem & ((1 << vl2) - 1);
        const int gy = y0 + row, gx = x0 + col;
        ok[u] = row < TH && gy < p.H && gx < p.W;
        const size_t px = ok[u] ? (size_t)gy * p.W + gx : 0;
        dst[u] = px * p.out_cs + 2 * Q + v * 8;
        if (ok[u]) val[u] = __ldg(reinterpret_cast<const uint4 *>(img + px * p.in_cs + 2 * Q) + v);
        row += drow;
        rem += drem;
        if (rem >= epr) { rem -= epr; ++row; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (ok[u]) *reinterpret_cast<uint4 *>(oimg + dst[u]) = val[u];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- conv2 (5x5 + SiLU): quarter a over the tile + 3-pixel halo -> t_a (shared, bf16, zero outside the map);
  //      quarter b over the tile -> global channels [c/4, c/2).  Items run over (pair fastest, column, row block); NT is
  //      a multiple of CPQ, so a thread keeps its pair and reloads its 25 taps only when it moves from a to b.
  const uint32_t pa_s = (uint32_t)__cvta_generic_to_shared(pa), pb_s = (uint32_t)__cvta_generic_to_shared(pb);
  const uint32_t tm_s = (uint32_t)__cvta_generic_to_shared(tm);
  const int items_a = (MHp / RO) * MW * cpq;
  const int items_b = (TH / RO) * TW * cpq;
  const size_t orow = (size_t)p.W * p.out_cs;  // elements per output row
  {
    float2 wreg[25];
    float2 bias = make_float2(0.f, 0.f);
    int cur = -1;
    for (int it = threadIdx.x; it < items_a + items_b; it += NT) {
      const bool is_a = it < items_a;
      const int j = is_a ? it : it - items_a;
      const int wdt = is_a ? MW : TW;
      const int cp = j % cpq, t = j / cpq;
      const int w = t % wdt, rb = t / wdt;
      const int ch = (is_a ? 0 : Q) + cp * 2;  // channel of the pair within conv2's c/2 channels
      if (ch != cur) {
        cur = ch;
#pragma unroll
        for (int k = 0; k < 25; ++k) wreg[k] = __ldg(reinterpret_cast<const float2 *>(p.w2 + (size_t)k * (2 * Q) + ch));
        bias = __ldg(reinterpret_cast<const float2 *>(p.b2 + ch));
      }
      float2 acc[RO];
#pragma unroll
      for (int o = 0; o < RO; ++o) acc[o] = bias;
      if (is_a) {
        dw2d_accum<5, 0, 5, RO, cpq * 4>(pa_s + (uint32_t)((rb * RO * AW + w) * cpq + cp) * 4, AW * cpq * 4, wreg, acc);
        const int gx = x0 - 3 + w;
        const bool col_in = gx >= 0 && gx < p.W;
        uint32_t ta = tm_s + (uint32_t)((rb * RO * MW + w) * cpq + cp) * 4;
#pragma unroll
        for (int o = 0; o < RO; ++o) {
          const int gy = y0 - 3 + rb * RO + o;
          const bool in = col_in && gy >= 0 && gy < p.H;
          sts32(ta, in ? f2_to_bf2(silu_tanh_f(acc[o].x), silu_tanh_f(acc[o].y)) : 0u);
          ta += MW * cpq * 4;
        }
      } else {
        dw2d_accum<5, 0, 5, RO, cpq * 4>(pb_s + (uint32_t)((rb * RO * BW + w) * cpq + cp) * 4, BW * cpq * 4, wreg, acc);
        const int gy0 = y0 + rb * RO, gx = x0 + w;
        bf16 *dst = oimg + ((size_t)gy0 * p.W + gx) * p.out_cs + Q + cp * 2;
#pragma unroll
        for (int o = 0; o < RO; ++o) {
          if (gy0 + o < p.H && gx < p.W)
            *reinterpret_cast<uint32_t *>(dst) = f2_to_bf2(silu_tanh_f(acc[o].x), silu_tanh_f(acc[o].y));
          dst += orow;
        }
      }
    }
  }
  __syncthreads();
  // ---- conv3 (7x7 + SiLU) over t_a -> global channels [0, c/4): two row passes (taps rows 0..3, then 4..6)
  {
    for (int it = threadIdx.x; it < items_b; it += NT) {
      const int cp = it % cpq, t = it / cpq;
      const int w = t % TW, rb = t / TW;
      const uint32_t src = tm_s + (uint32_t)((rb * RO * MW + w) * cpq + cp) * 4;
      float2 acc[RO];
      const float2 bias = __ldg(reinterpret_cast<const float2 *>(p.b3 + cp * 2));
#pragma unroll
      for (int o = 0; o < RO; ++o) acc[o] = bias;
      {
        float2 wreg[28];
#pragma unroll
        for (int k = 0; k < 28; ++k) wreg[k] = __ldg(reinterpret_cast<const float2 *>(p.w3 + (size_t)k * Q + cp * 2));
        dw2d_accum<7, 0, 4, RO, cpq * 4>(src, MW * cpq * 4, wreg, acc);
      }
      {
        float2 wreg[21];
#pragma unroll
        for (int k = 0; k < 21; ++k) wreg[k] = __ldg(reinterpret_cast<const float2 *>(p.w3 + (size_t)(28 + k) * Q + cp * 2));
        dw2d_accum<7, 4, 7, RO, cpq * 4>(src, MW * cpq * 4, wreg, acc);
      }
      const int gy0 = y0 + rb * RO, gx = x0 + w;
      bf16 *dst = oimg + ((size_t)gy0 * p.W + gx) * p.out_cs + cp * 2;
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        if (gy0 + o < p.H && gx < p.W)
          *reinterpret_cast<uint32_t *>(dst) = f2_to_bf2(silu_tanh_f(acc[o].x), silu_tanh_f(acc[o].y));
        dst += orow;
      }
    }
  }
}

size_t pmsfa_smem(int c, int TH, int TW) {
  const int cpq = c >> 3;
  const int MHp = (TH + 6 + RO - 1) / RO * RO;
  const int tm_rows = MHp > TH + 6 ? MHp : TH + 6;
  return ((size_t)(MHp + 4) * (TW + 10) + (size_t)(TH + 4) * (TW + 4) + (size_t)tm_rows * (TW + 6)) * cpq * 4;
}
}  // namespace

int launch_pmsfa_dw(const PmsfaParams &q, cudaStream_t st) {
  B2Y_REQUIRE(q.c == 16 || q.c == 32 || q.c == 64 || q.c == 128, B2Y_ERR_UNSUPPORTED, "pmsfa_dw: c = %d not in {16,32,64,128}", q.c);
  B2Y_REQUIRE(q.in_cs % 8 == 0 && q.out_cs % 8 == 0 && (reinterpret_cast<uintptr_t>(q.in) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(q.out) & 15) == 0,
              B2Y_ERR_UNSUPPORTED, "pmsfa_dw: views must be 16-byte aligned");
  PmsfaK p;
  p.in = q.in; p.in_cs = q.in_cs; p.out = q.out; p.out_cs = q.out_cs;
  p.w2 = q.w2; p.b2 = q.b2; p.w3 = q.w3; p.b3 = q.b3;
  p.B = q.B; p.H = q.H; p.W = q.W; p.c = q.c;
  p.n8_log2 = q.c == 16 ? 0 : q.c == 32 ? 1 : q.c == 64 ? 2 : 3;
  // tile: wide maps take 16 x 32 tiles; small maps (40x40, 20x20) are split so that the launch still has >= ~4 CTAs per SM
  int TH = q.H < 16 ? q.H : 16, TW = q.W < 32 ? q.W : 32;
  auto ntiles = [&](int th, int tw) { return (long long)q.B * ceil_div(q.H, th) * ceil_div(q.W, tw); };
  while (ntiles(TH, TW) < 148 * 4 && (TH > 8 || TW > 8)) {
    if (TW >= 2 * TH && TW > 8) TW = (TW + 1) / 2;
    else if (TH > 8) TH = (TH + 1) / 2;
    else TW = (TW + 1) / 2;
  }
  TH = (TH + RO - 1) / RO * RO;  // whole row blocks
  p.TH = TH; p.TW = TW;
  p.tiles_x = ceil_div(q.W, TW); p.tiles_y = ceil_div(q.H, TH);
  const size_t smem = pmsfa_smem(q.c, TH, TW);
  B2Y_REQUIRE(smem <= 200 * 1024, B2Y_ERR_UNSUPPORTED, "pmsfa_dw: tile does not fit shared memory");
  auto go = [&](auto kern) -> int {
    B2Y_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b2y_launch(kern, dim3((unsigned)ntiles(TH, TW)), dim3(NT), smem, st, p);
    B2Y_LAUNCH_CHECK();
    return B2Y_OK;
  };
  static const int minb = [] { const char *e = getenv("B2Y_PMSFA_MINB"); return e ? atoi(e) : 3; }();  // A/B knob: 3 (159 regs) | 4
  if (minb == 4)
    return q.c == 16 ? go(pmsfa_dw_kernel<2, 4>) : q.c == 32 ? go(pmsfa_dw_kernel<4, 4>) : q.c == 64 ? go(pmsfa_dw_kernel<8, 4>)
                                                                                                  : go(pmsfa_dw_kernel<16, 4>);
  return q.c == 16 ? go(pmsfa_dw_kernel<2, 3>) : q.c == 32 ? go(pmsfa_dw_kernel<4, 3>) : q.c == 64 ? go(pmsfa_dw_kernel<8, 3>)
                                                                                                : go(pmsfa_dw_kernel<16, 3>);
}

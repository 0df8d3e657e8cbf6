// Fused depthwise tail of a PMSFA block (ultralytics/nn/modules/block.py:3035-3053), one kernel:
//
//     conv1_out = [a | b | h]        (c channels: quarters a, b of c/4 and the half h of c/2; written by conv1, tensor cores)
//     t   = SiLU(dw5x5([a | b]))      conv2 on the first half              (block.py:3047)
//     out = [ SiLU(dw7x7(t_a)) | t_b | h ]                                  (block.py:3049-3051: cat([conv3_out, conv2_out_2,
//                                                                            conv1_out_2]) -- the input of conv4)
//
// Unfused this is 3 launches and 2 round trips through HBM (t, the cat copy) at 6-11 % of the HBM roof:
// the per-thread row-strip kernels are latency-bound at 8 warps per SM.  Here one CTA owns a TH x TW pixel tile:
//   * cp.async stages quarter a with a 5-pixel halo and quarter b with a 2-pixel halo into shared memory as two dense
//     planes [row][col][c/4] (zero-filled outside the map = the convs' zero padding);
//   * conv2 runs from shared memory: on a over the tile + 3-pixel halo (what conv3 needs) -> bf16 plane t_a in shared
//     memory (zero outside the map: conv3 pads its INPUT with zeros), on b over the tile -> global;
//   * conv3 runs over t_a -> global; the half h is passed through with 16-byte copies.
// Work item = 4 channels x 1 column x 4 output rows (vertical register block): consecutive lanes take consecutive
// 8-byte channel quads of consecutive columns, so every shared-memory access of a warp is one contiguous, conflict-free
// segment; a loaded input column (4 + k - 1 pixels) feeds 4 x k FMAs per channel, weights are float4 broadcasts.
// The layer is CUDA-core bound (396 MAC per pixel at c = 16 against 64 B of traffic), not HBM bound: the bound quoted
// in DESIGN.md is the fp32 FMA rate.  t is rounded to bf16 exactly where the unfused path stored it.
#include "common.cuh"

namespace {
constexpr int NT = 256;
constexpr int RO = 4;  // output rows per work item

struct PmsfaK {
  const bf16 *in; int in_cs;
  bf16 *out; int out_cs;
  const float *w2, *b2, *w3, *b3;  // [25][c/2], [c/2], [49][c/4], [c/4]
  int B, H, W, c;
  int TH, TW, tiles_x, tiles_y;
};

__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 8 : 0;  // src-size 0: the 8 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ float silu_t(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// K x K depthwise conv of RO vertically adjacent outputs (rows r0.., column w) for channel quad qd of a plane
// [rows][pitch][q4 quads]; plane row r0 / column w is the top-left tap of output row 0.  wsm: [K*K][nq] float4.
template <int K>
__device__ __forceinline__ void dw_block(const uint2 *plane, int pitch, int q4, int r0, int w, int qd, const float4 *wsm,
                                         int nq, int wq, float4 bias, float4 (&acc)[RO]) {
#pragma unroll
  for (int o = 0; o < RO; ++o) acc[o] = bias;
#pragma unroll
  for (int s = 0; s < K; ++s) {
    float x[RO + K - 1][4];
#pragma unroll
    for (int i = 0; i < RO + K - 1; ++i) unpack4(plane[((size_t)(r0 + i) * pitch + (w + s)) * q4 + qd], x[i]);
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const float4 ww = wsm[(r * K + s) * nq + wq];
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        acc[o].x = fmaf(x[o + r][0], ww.x, acc[o].x);
        acc[o].y = fmaf(x[o + r][1], ww.y, acc[o].y);
        acc[o].z = fmaf(x[o + r][2], ww.z, acc[o].z);
        acc[o].w = fmaf(x[o + r][3], ww.w, acc[o].w);
      }
    }
  }
}

__global__ void __launch_bounds__(NT) pmsfa_dw_kernel(PmsfaK p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  const int Q = p.c >> 2, q4 = Q >> 2, nq2 = 2 * q4;
  const int TH = p.TH, TW = p.TW;
  const int AH = TH + 10, AW = TW + 10;      // plane a: tile + 5-pixel halo (2 for conv2 + 3 for conv3)
  const int BH = TH + 4, BW = TW + 4;        // plane b: tile + 2-pixel halo
  const int MH = TH + 6, MW = TW + 6;        // t_a: tile + 3-pixel halo; rows padded up to a multiple of RO below
  const int MHp = (MH + RO - 1) / RO * RO;
  // shared memory: w2 | w3 | plane a (+ RO-1 slack rows for the padded last row block) | plane b | t_a
  float4 *w2s = reinterpret_cast<float4 *>(smem_raw);
  float4 *w3s = w2s + 25 * nq2;
  uint2 *pa = reinterpret_cast<uint2 *>(w3s + 49 * q4);
  uint2 *pb = pa + (size_t)(MHp + 4) * AW * q4;
  uint2 *tm = pb + (size_t)BH * BW * q4;
  for (int i = threadIdx.x; i < 25 * nq2; i += NT) w2s[i] = reinterpret_cast<const float4 *>(p.w2)[i];
  for (int i = threadIdx.x; i < 49 * q4; i += NT) w3s[i] = reinterpret_cast<const float4 *>(p.w3)[i];

  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, b = tile / (p.tiles_x * p.tiles_y);
  const int y0 = ty * TH, x0 = tx * TW;
  const bf16 *img = p.in + (size_t)b * p.H * p.W * p.in_cs;
  bf16 *oimg = p.out + (size_t)b * p.H * p.W * p.out_cs;
  pdl_wait();  // weights are constants; the activations come from the preceding kernel

  // ---- stage the two planes (8-byte quads; zero outside the map)
  {
    const uint32_t pa_s = (uint32_t)__cvta_generic_to_shared(pa), pb_s = (uint32_t)__cvta_generic_to_shared(pb);
    const int na = (MHp + 4) * AW * q4;
    for (int i = threadIdx.x; i < na; i += NT) {
      const int qd = i % q4, col = (i / q4) % AW, row = i / (q4 * AW);
      const int gy = y0 - 5 + row, gx = x0 - 5 + col;
      const bool ok = row < AH && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      cp_async8(pa_s + i * 8, img + ((size_t)(ok ? gy : 0) * p.W + (ok ? gx : 0)) * p.in_cs + qd * 4, ok);
    }
    const int nb = BH * BW * q4;
    for (int i = threadIdx.x; i < nb; i += NT) {
      const int qd = i % q4, col = (i / q4) % BW, row = i / (q4 * BW);
      const int gy = y0 - 2 + row, gx = x0 - 2 + col;
      const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      cp_async8(pb_s + i * 8, img + ((size_t)(ok ? gy : 0) * p.W + (ok ? gx : 0)) * p.in_cs + Q + qd * 4, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // ---- pass the half h through while the planes land: 16-byte vectors, c/2 channels = c bytes per pixel
  {
    const int vpp = p.c >> 4;  // 16-byte vectors per pixel
    const int nv = TH * TW * vpp;
    for (int i = threadIdx.x; i < nv; i += NT) {
      const int v = i % vpp, col = (i / vpp) % TW, row = i / (vpp * TW);
      const int gy = y0 + row, gx = x0 + col;
      if (gy < p.H && gx < p.W) {
        const size_t px = (size_t)gy * p.W + gx;
        const uint4 val = __ldg(reinterpret_cast<const uint4 *>(img + px * p.in_cs + 2 * Q) + v);
        *(reinterpret_cast<uint4 *>(oimg + px * p.out_cs + 2 * Q) + v) = val;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- conv2 (5x5 + SiLU): quarter a over the tile + 3-pixel halo -> t_a (shared, bf16, zero outside the map);
  //      quarter b over the tile -> global channels [c/4, c/2)
  const int items_a = (MHp / RO) * MW * q4;
  const int items_b = ((TH + RO - 1) / RO) * TW * q4;
  for (int it = threadIdx.x; it < items_a + items_b; it += NT) {
    float4 acc[RO];
    if (it < items_a) {
      const int qd = it % q4, w = (it / q4) % MW, rb = it / (q4 * MW);
      const float4 bias = *reinterpret_cast<const float4 *>(p.b2 + qd * 4);
      dw_block<5>(pa, AW, q4, rb * RO, w, qd, w2s, nq2, qd, bias, acc);
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        const int r = rb * RO + o;
        const int gy = y0 - 3 + r, gx = x0 - 3 + w;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
          v[0] = silu_t(acc[o].x); v[1] = silu_t(acc[o].y); v[2] = silu_t(acc[o].z); v[3] = silu_t(acc[o].w);
        }
        tm[((size_t)r * MW + w) * q4 + qd] = pack4(v);
      }
    } else {
      const int j = it - items_a;
      const int qd = j % q4, w = (j / q4) % TW, rb = j / (q4 * TW);
      const float4 bias = *reinterpret_cast<const float4 *>(p.b2 + Q + qd * 4);
      dw_block<5>(pb, BW, q4, rb * RO, w, qd, w2s, nq2, q4 + qd, bias, acc);
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        const int gy = y0 + rb * RO + o, gx = x0 + w;
        if (rb * RO + o < TH && gy < p.H && gx < p.W) {
          const float v[4] = {silu_t(acc[o].x), silu_t(acc[o].y), silu_t(acc[o].z), silu_t(acc[o].w)};
          *reinterpret_cast<uint2 *>(oimg + ((size_t)gy * p.W + gx) * p.out_cs + Q + qd * 4) = pack4(v);
        }
      }
    }
  }
  __syncthreads();
  // ---- conv3 (7x7 + SiLU) over t_a -> global channels [0, c/4)
  for (int it = threadIdx.x; it < items_b; it += NT) {
    const int qd = it % q4, w = (it / q4) % TW, rb = it / (q4 * TW);
    float4 acc[RO];
    const float4 bias = *reinterpret_cast<const float4 *>(p.b3 + qd * 4);
    dw_block<7>(tm, MW, q4, rb * RO, w, qd, w3s, q4, qd, bias, acc);
#pragma unroll
    for (int o = 0; o < RO; ++o) {
      const int gy = y0 + rb * RO + o, gx = x0 + w;
      if (rb * RO + o < TH && gy < p.H && gx < p.W) {
        const float v[4] = {silu_t(acc[o].x), silu_t(acc[o].y), silu_t(acc[o].z), silu_t(acc[o].w)};
        *reinterpret_cast<uint2 *>(oimg + ((size_t)gy * p.W + gx) * p.out_cs + qd * 4) = pack4(v);
      }
    }
  }
}

size_t pmsfa_smem(int c, int TH, int TW) {
  const int q4 = c >> 4;
  const int MHp = (TH + 6 + RO - 1) / RO * RO;
  const int THp = (TH + RO - 1) / RO * RO;
  // t_a needs rows up to THp - 1 + 6 for the padded last row block of conv3
  const int tm_rows = (MHp > THp + 6) ? MHp : THp + 6;
  return (size_t)(25 * 2 * q4 + 49 * q4) * 16 +
         ((size_t)(MHp + 4) * (TW + 10) + (size_t)(THp + 4) * (TW + 4) + (size_t)tm_rows * (TW + 6)) * q4 * 8;
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
  B2Y_CUDA_CHECK(cudaFuncSetAttribute(pmsfa_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  b2y_launch(pmsfa_dw_kernel, dim3((unsigned)ntiles(TH, TW)), dim3(NT), smem, st, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

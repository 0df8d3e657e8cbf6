// Depthwise 3x3 convolution, stride 1 or 2, pad 1, + bias (+ SiLU) (+ residual add / mul): DWConv (ultralytics/nn/modules/
// conv.py:224-229) as used by the Detect class branch (head.py:50-55), SCDown.cv2 (block.py:4503-4535) and Attention.pe
// (block.py:4286).  9 MAC per output element: HBM-bound, 2*C B in (stride 1) / 8*C B in (stride 2) + 2*C B out per output pixel.
//
// One CTA = one image x one TH x TW output tile x one chunk of up to 64 channels.  cp.async stages the input tile with its
// 1-pixel halo ((TH-1)*S+3 rows x (TW-1)*S+3 columns, zero-filled outside the map) as [row][col][channel pair]; a thread owns
// one channel pair and one output column and walks down the rows with its 9 taps in registers (input-stationary: a loaded
// row of 3 words feeds every output row it touches).  Lanes run over channel pairs first, so shared-memory reads and global
// stores of a warp are contiguous 64-128 byte runs.
// The per-thread row-strip kernels in conv_simt.cu (no staging, weights re-read through L1) reach 0.1-0.3 of the HBM roof
// on these layers, the grouped tensor-core lowering (conv_tc.cu, C <= 128) ~0.45: both are latency-bound, not byte-bound.
#include "dw_regs.cuh"

namespace {
constexpr int NT = 256;
constexpr int RO = 4;  // output rows per work item

struct Dw3K {
  const bf16 *in; int in_cs;
  bf16 *out; int out_cs;
  const bf16 *res; int res_cs; int res_mode;
  const float *w, *bias;  // [9][C], [C]
  int B, H, W, Ho, Wo, C, act;
  int TH, TW, tiles_x, tiles_y, CC, nchunks;  // CC: channels per chunk (multiple of 8)
  int in_stride;                              // input channels between the chunks (== CC unless the input is a chunked gather)
};

template <int S>
__global__ void __launch_bounds__(NT, 2) dw3_tile_kernel(Dw3K p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  const int TH = p.TH, TW = p.TW;
  const int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
  int t = blockIdx.x;
  const int chunk = t % p.nchunks; t /= p.nchunks;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y, b = t / p.tiles_y;
  const int c0 = chunk * p.CC;
  const int cc = min(p.CC, p.C - c0);  // channels of this chunk (multiple of 8)
  const int npairs = cc >> 1, nvec = cc >> 3;
  const int y0 = ty * TH, x0 = tx * TW;
  const bf16 *img = p.in + (size_t)b * p.H * p.W * p.in_cs + chunk * p.in_stride;
  pdl_wait();
  {  // stage the input tile: 16-byte vectors, zero outside the map
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const int epr = IW * nvec;
    int row = threadIdx.x / epr, rem = threadIdx.x % epr;
    const int drow = NT / epr, drem = NT % epr;
    const int gy0 = y0 * S - 1, gx0 = x0 * S - 1;
    while (row < IH) {
      const int col = rem / nvec, v = rem - col * nvec;
      const int gy = gy0 + row, gx = gx0 + col;
      const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      cp_async_16(base + (uint32_t)((row * IW + col) * npairs * 4 + v * 16), img + ((size_t)(ok ? gy : 0) * p.W + (ok ? gx : 0)) * p.in_cs + v * 8, ok);
      row += drow;
      rem += drem;
      if (rem >= epr) { rem -= epr; ++row; }
    }
    cp_async_commit_wait();
  }
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const int csb = npairs * 4, rsb = IW * csb;  // bytes between columns / rows of the staged tile
  const int items = npairs * TW * (TH / RO);
  float2 w[9], bias = make_float2(0.f, 0.f);
  int cur = -1;
  for (int it = threadIdx.x; it < items; it += NT) {
    const int cp = it % npairs, q = it / npairs;
    const int x = q % TW, rb = q / TW;
    if (cp != cur) {
      cur = cp;
#pragma unroll
      for (int k = 0; k < 9; ++k) w[k] = __ldg(reinterpret_cast<const float2 *>(p.w + (size_t)k * p.C + c0 + cp * 2));
      bias = __ldg(reinterpret_cast<const float2 *>(p.bias + c0 + cp * 2));
    }
    float2 acc[RO];
#pragma unroll
    for (int o = 0; o < RO; ++o) acc[o] = bias;
    uint32_t ra = sbase + (uint32_t)(rb * RO * S) * rsb + (uint32_t)(x * S) * csb + cp * 4;
#pragma unroll
    for (int i = 0; i < (RO - 1) * S + 3; ++i) {
      const float2 a0 = bf2_to_f2(lds32(ra)), a1 = bf2_to_f2(lds32(ra + csb)), a2 = bf2_to_f2(lds32(ra + 2 * csb));
      ra += rsb;
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        const int r = i - o * S;
        if (r >= 0 && r < 3) {
          acc[o].x = fmaf(a0.x, w[r * 3].x, acc[o].x); acc[o].y = fmaf(a0.y, w[r * 3].y, acc[o].y);
          acc[o].x = fmaf(a1.x, w[r * 3 + 1].x, acc[o].x); acc[o].y = fmaf(a1.y, w[r * 3 + 1].y, acc[o].y);
          acc[o].x = fmaf(a2.x, w[r * 3 + 2].x, acc[o].x); acc[o].y = fmaf(a2.y, w[r * 3 + 2].y, acc[o].y);
        }
      }
    }
    const int gx = x0 + x, gy0 = y0 + rb * RO;
    if (gx < p.Wo) {
      const size_t px0 = ((size_t)b * p.Ho + gy0) * p.Wo + gx;
#pragma unroll
      for (int o = 0; o < RO; ++o) {
        if (gy0 + o >= p.Ho) break;
        float vx = acc[o].x, vy = acc[o].y;
        if (p.act) { vx = silu_tanh_f(vx); vy = silu_tanh_f(vy); }
        const size_t px = px0 + (size_t)o * p.Wo;
        if (p.res_mode) {
          const float2 r = bf2_to_f2(__ldg(reinterpret_cast<const uint32_t *>(p.res + px * p.res_cs + c0 + cp * 2)));
          if (p.res_mode == B2Y_RES_ADD) { vx += r.x; vy += r.y; }
          else { vx *= r.x; vy *= r.y; }
        }
        *reinterpret_cast<uint32_t *>(p.out + px * p.out_cs + c0 + cp * 2) = f2_to_bf2(vx, vy);
      }
    }
  }
}
}  // namespace

bool dw3_tile_supported(const ConvParams &p) {
  if (!(p.kh == 3 && p.kw == 3 && p.ph == 1 && p.pw == 1 && p.sh == p.sw && (p.sh == 1 || p.sh == 2))) return false;
  if (p.cin != p.cout || p.out_f32 || p.cout % 8 || p.in_cs % 8 || p.out_cs % 2) return false;
  if (p.dw_chunk && (p.dw_chunk % 8 || p.dw_chunk > 64 || p.dw_in_stride % 8)) return false;
  if ((reinterpret_cast<uintptr_t>(p.in) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 3)) return false;
  if (p.res_mode == B2Y_RES_PREADD_UP2) return false;
  if (p.res_mode && ((p.res_cs % 2) || (reinterpret_cast<uintptr_t>(p.res) & 3))) return false;
  return p.Ho == (p.H + 2 - 3) / p.sh + 1 && p.Wo == (p.W + 2 - 3) / p.sw + 1;
}

int launch_dw3_tile(const ConvParams &c, cudaStream_t st) {
  Dw3K p;
  p.in = c.in; p.in_cs = c.in_cs; p.out = reinterpret_cast<bf16 *>(c.out); p.out_cs = c.out_cs;
  p.res = c.res; p.res_cs = c.res_cs; p.res_mode = c.res_mode;
  p.w = reinterpret_cast<const float *>(c.w); p.bias = c.bias;
  p.B = c.B; p.H = c.H; p.W = c.W; p.Ho = c.Ho; p.Wo = c.Wo; p.C = c.cout; p.act = c.act;
  const int S = c.sh;
  p.nchunks = ceil_div(p.C, 64);
  p.CC = (ceil_div(p.C, p.nchunks) + 7) / 8 * 8;
  p.nchunks = ceil_div(p.C, p.CC);
  p.in_stride = p.CC;
  if (c.dw_chunk) {  // the input is a gather of equal chunks (Attention.pe: the v slice of every head of qkv)
    p.CC = c.dw_chunk;
    p.nchunks = p.C / c.dw_chunk;
    p.in_stride = c.dw_in_stride;
  }
  // tile: split the map evenly into tiles of at most 16 x 32 (stride 1) / 8 x 16 (stride 2) outputs, rows in blocks of RO
  const int maxh = S == 1 ? 16 : 8, maxw = S == 1 ? 32 : 16;
  int TH = ceil_div(p.Ho, ceil_div(p.Ho, maxh)), TW = ceil_div(p.Wo, ceil_div(p.Wo, maxw));
  TH = (TH + RO - 1) / RO * RO;
  // small launches: shrink the tile until there are ~2 CTAs per SM worth of tiles
  auto ntiles = [&](int th, int tw) { return (long long)p.B * ceil_div(p.Ho, th) * ceil_div(p.Wo, tw) * p.nchunks; };
  while (ntiles(TH, TW) < 148 * 2 && (TH > RO || TW > 8)) {
    if (TW > 8 && TW >= TH) TW = (TW + 1) / 2;
    else if (TH > RO) TH = ((TH / 2) + RO - 1) / RO * RO;
    else break;
  }
  p.TH = TH; p.TW = TW;
  p.tiles_y = ceil_div(p.Ho, TH); p.tiles_x = ceil_div(p.Wo, TW);
  const size_t smem = (size_t)((TH - 1) * S + 3) * ((TW - 1) * S + 3) * p.CC * 2;
  B2Y_REQUIRE(smem <= 100 * 1024, B2Y_ERR_UNSUPPORTED, "dw3_tile: tile does not fit shared memory");
  const long long grid = ntiles(TH, TW);
  if (S == 1) {
    B2Y_CUDA_CHECK(cudaFuncSetAttribute(dw3_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b2y_launch(dw3_tile_kernel<1>, dim3((unsigned)grid), dim3(NT), smem, st, p);
  } else {
    B2Y_CUDA_CHECK(cudaFuncSetAttribute(dw3_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b2y_launch(dw3_tile_kernel<2>, dim3((unsigned)grid), dim3(NT), smem, st, p);
  }
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

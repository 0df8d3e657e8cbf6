// CUDA-core convolution kernels: generic dense conv (validation twin of the tcgen05 kernel and the path for
// shapes that kernel does not take) and depthwise kh x kw.
// Reference semantics: ultralytics/nn/modules/conv.py:145-151 (act(conv(x)+b) with BN folded), DWConv :224-229.
#include <stdlib.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------ dense, generic
// thread = 1 output pixel x 8 output channels; weights [cout_pad][kh*kw*cin] bf16 are warp-uniform (broadcast loads).
template <bool VEC>
__global__ void __launch_bounds__(128) conv_simt_kernel(ConvParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long npix = (long long)p.B * p.Ho * p.Wo;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int co0 = blockIdx.y * 8;
  const int wo = (int)(pix % p.Wo);
  const int ho = (int)((pix / p.Wo) % p.Ho);
  const int b = (int)(pix / ((long long)p.Wo * p.Ho));
  const int K = p.kh * p.kw * p.cin;
  const bf16 *wbase = reinterpret_cast<const bf16 *>(p.w) + (size_t)co0 * K;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;

  for (int r = 0; r < p.kh; ++r) {
    const int hi = ho * p.sh - p.ph + r;
    if (hi < 0 || hi >= p.H) continue;
    for (int s = 0; s < p.kw; ++s) {
      const int wi = wo * p.sw - p.pw + s;
      if (wi < 0 || wi >= p.W) continue;
      const bf16 *x = p.in + ((size_t)(b * p.H + hi) * p.W + wi) * p.in_cs;
      const bf16 *wt = wbase + (r * p.kw + s) * p.cin;
      if (VEC) {
        for (int c = 0; c < p.cin; c += 8) {
          float xv[8];
          unpack8(*reinterpret_cast<const uint4 *>(x + c), xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float wv[8];
            unpack8(__ldg(reinterpret_cast<const uint4 *>(wt + (size_t)j * K + c)), wv);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j] = fmaf(xv[i], wv[i], acc[j]);
          }
        }
      } else {
        for (int c = 0; c < p.cin; ++c) {
          const float xv = __bfloat162float(x[c]);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, __bfloat162float(wt[(size_t)j * K + c]), acc[j]);
        }
      }
    }
  }
  const size_t opix = ((size_t)(b * p.Ho + ho) * p.Wo + wo);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = co0 + j;
    if (co >= p.cout) break;
    float v = acc[j] + p.bias[co];
    if (p.res_mode == B2Y_RES_PREADD_UP2)  // pre-activation add of the nearest-2x-upsampled half-resolution map
      v += __bfloat162float(p.res[((size_t)(b * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1)) * p.res_cs + co]);
    if (p.act) v = silu_f(v);
    if (p.res_mode == B2Y_RES_ADD || p.res_mode == B2Y_RES_MUL) {
      const float r = __bfloat162float(p.res[opix * p.res_cs + co]);
      v = (p.res_mode == B2Y_RES_ADD) ? v + r : v * r;
    }
    if (p.out_f32) reinterpret_cast<float *>(p.out)[opix * p.out_cs + co] = v;
    else reinterpret_cast<bf16 *>(p.out)[opix * p.out_cs + co] = __float2bfloat16_rn(v);
  }
}

int launch_conv_simt(const ConvParams &p, cudaStream_t st) {
  B2Y_REQUIRE(p.epi_mode == 0, B2Y_ERR_UNSUPPORTED,
              "fused Detect decode exists only in the tcgen05 conv epilogue (plan with fuse_decode=False for the SIMT path)");
  const long long npix = (long long)p.B * p.Ho * p.Wo;
  dim3 grid((unsigned)((npix + 127) / 128), (unsigned)ceil_div(p.cout, 8));
  const bool vec = (p.cin % 8 == 0) && (p.in_cs % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.w) & 15) == 0);
  if (vec) b2y_launch(conv_simt_kernel<true>, dim3(grid), dim3(128), 0, st, p);
  else b2y_launch(conv_simt_kernel<false>, dim3(grid), dim3(128), 0, st, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------ depthwise
// thread = 1 output pixel x V channels (V = 8: 16-byte, 4: 8-byte, 1: scalar accesses); weights fp32 [kh*kw][c].
template <int V>
__global__ void __launch_bounds__(256) dwconv_kernel(ConvParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int cg = p.cout / V;
  const long long total = (long long)p.B * p.Ho * p.Wo * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % cg) * V;
  const long long pix = idx / cg;
  const int wo = (int)(pix % p.Wo);
  const int ho = (int)((pix / p.Wo) % p.Ho);
  const int b = (int)(pix / ((long long)p.Wo * p.Ho));
  const float *w = reinterpret_cast<const float *>(p.w);
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = p.bias[c0 + j];
  for (int r = 0; r < p.kh; ++r) {
    const int hi = ho * p.sh - p.ph + r;
    if (hi < 0 || hi >= p.H) continue;
    for (int s = 0; s < p.kw; ++s) {
      const int wi = wo * p.sw - p.pw + s;
      if (wi < 0 || wi >= p.W) continue;
      const bf16 *x = p.in + ((size_t)(b * p.H + hi) * p.W + wi) * p.in_cs + c0;
      const float *wt = w + (size_t)(r * p.kw + s) * p.cout + c0;
      float xv[V];
      if (V == 8) unpack8(*reinterpret_cast<const uint4 *>(x), xv);
      else if (V == 4) unpack4(*reinterpret_cast<const uint2 *>(x), xv);
      else xv[0] = __bfloat162float(x[0]);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = fmaf(xv[j], __ldg(wt + j), acc[j]);
    }
  }
  const size_t opix = (size_t)pix;
  if (p.act) {
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = silu_f(acc[j]);
  }
  if (p.res_mode) {
    const bf16 *r = p.res + opix * p.res_cs + c0;
    float rv[V];
    if (V == 8) unpack8(*reinterpret_cast<const uint4 *>(r), rv);
    else if (V == 4) unpack4(*reinterpret_cast<const uint2 *>(r), rv);
    else rv[0] = __bfloat162float(r[0]);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = (p.res_mode == B2Y_RES_ADD) ? acc[j] + rv[j] : acc[j] * rv[j];
  }
  bf16 *o = reinterpret_cast<bf16 *>(p.out) + opix * p.out_cs + c0;
  if (V == 8) *reinterpret_cast<uint4 *>(o) = pack8(acc);
  else if (V == 4) *reinterpret_cast<uint2 *>(o) = pack4(acc);
  else o[0] = __float2bfloat16_rn(acc[0]);
}

// Register-blocked variant for 16-byte-aligned channel groups: thread = 4 consecutive output pixels of one row x 8
// channels.  Each input pixel of the (kw + 3*sw) wide window is loaded once and feeds up to 4 outputs; the weights of
// the whole layer sit in shared memory ([kh*kw][c] fp32 + bias) and are read as broadcast float4s.
template <int KW, int SW>  // KW == 0: runtime kernel width / stride
__global__ void __launch_bounds__(256) dwconv4_kernel(ConvParams p, int wq) {
  pdl_launch_dependents();
  extern __shared__ float wsm[];
  const int taps = p.kh * p.kw;
  for (int i = threadIdx.x; i < taps * p.cout; i += blockDim.x) wsm[i] = reinterpret_cast<const float *>(p.w)[i];
  for (int i = threadIdx.x; i < p.cout; i += blockDim.x) wsm[taps * p.cout + i] = p.bias[i];
  __syncthreads();
  pdl_wait();
  const int kw = KW ? KW : p.kw, sw = KW ? SW : p.sw;
  const int cg = p.cout / 8;
  const long long total = (long long)p.B * p.Ho * wq * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % cg) * 8;
  long long t = idx / cg;
  const int wo0 = (int)(t % wq) * 4;
  t /= wq;
  const int ho = (int)(t % p.Ho);
  const int b = (int)(t / p.Ho);
  float acc[4][8];
  {
    const float4 b0 = *reinterpret_cast<const float4 *>(wsm + taps * p.cout + c0);
    const float4 b1 = *reinterpret_cast<const float4 *>(wsm + taps * p.cout + c0 + 4);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc[q][0] = b0.x; acc[q][1] = b0.y; acc[q][2] = b0.z; acc[q][3] = b0.w;
      acc[q][4] = b1.x; acc[q][5] = b1.y; acc[q][6] = b1.z; acc[q][7] = b1.w;
    }
  }
  const int span = 3 * sw + kw;
  for (int r = 0; r < p.kh; ++r) {
    const int hi = ho * p.sh - p.ph + r;
    if (hi < 0 || hi >= p.H) continue;
    const bf16 *row = p.in + ((size_t)(b * p.H + hi) * p.W) * p.in_cs + c0;
    const float *wr = wsm + (size_t)r * kw * p.cout + c0;
#pragma unroll
    for (int xx = 0; xx < (KW ? 3 * SW + KW : 1); ++xx) {
      if (!KW) break;
      const int wi = wo0 * sw - p.pw + xx;
      if (wi < 0 || wi >= p.W) continue;
      float xv[8];
      unpack8(*reinterpret_cast<const uint4 *>(row + (size_t)wi * p.in_cs), xv);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int tp = xx - q * SW;
        if (tp >= 0 && tp < KW) {
          const float4 w0 = *reinterpret_cast<const float4 *>(wr + tp * p.cout);
          const float4 w1 = *reinterpret_cast<const float4 *>(wr + tp * p.cout + 4);
          acc[q][0] = fmaf(xv[0], w0.x, acc[q][0]); acc[q][1] = fmaf(xv[1], w0.y, acc[q][1]);
          acc[q][2] = fmaf(xv[2], w0.z, acc[q][2]); acc[q][3] = fmaf(xv[3], w0.w, acc[q][3]);
          acc[q][4] = fmaf(xv[4], w1.x, acc[q][4]); acc[q][5] = fmaf(xv[5], w1.y, acc[q][5]);
          acc[q][6] = fmaf(xv[6], w1.z, acc[q][6]); acc[q][7] = fmaf(xv[7], w1.w, acc[q][7]);
        }
      }
    }
    if (!KW) {
      for (int xx = 0; xx < span; ++xx) {
        const int wi = wo0 * sw - p.pw + xx;
        if (wi < 0 || wi >= p.W) continue;
        float xv[8];
        unpack8(*reinterpret_cast<const uint4 *>(row + (size_t)wi * p.in_cs), xv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int tp = xx - q * sw;
          if (tp >= 0 && tp < kw) {
            const float4 w0 = *reinterpret_cast<const float4 *>(wr + tp * p.cout);
            const float4 w1 = *reinterpret_cast<const float4 *>(wr + tp * p.cout + 4);
            acc[q][0] = fmaf(xv[0], w0.x, acc[q][0]); acc[q][1] = fmaf(xv[1], w0.y, acc[q][1]);
            acc[q][2] = fmaf(xv[2], w0.z, acc[q][2]); acc[q][3] = fmaf(xv[3], w0.w, acc[q][3]);
            acc[q][4] = fmaf(xv[4], w1.x, acc[q][4]); acc[q][5] = fmaf(xv[5], w1.y, acc[q][5]);
            acc[q][6] = fmaf(xv[6], w1.z, acc[q][6]); acc[q][7] = fmaf(xv[7], w1.w, acc[q][7]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int wo = wo0 + q;
    if (wo >= p.Wo) break;
    const size_t opix = ((size_t)(b * p.Ho + ho) * p.Wo + wo);
    if (p.act) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = silu_f(acc[q][j]);
    }
    if (p.res_mode) {
      float rv[8];
      unpack8(*reinterpret_cast<const uint4 *>(p.res + opix * p.res_cs + c0), rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = (p.res_mode == B2Y_RES_ADD) ? acc[q][j] + rv[j] : acc[q][j] * rv[j];
    }
    *reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(p.out) + opix * p.out_cs + c0) = pack8(acc[q]);
  }
}

// Row-strip variant for stride-1 kernels whose taps fit the register file (3x3, 1xk, kx1, 5x5):
// thread = 4 channels x one output column x a strip of R output rows.  The KH*KW*4 weights live in registers for the
// whole strip; every input row of the strip is loaded ONCE per thread (KW 8-byte pixels, column neighbours hit in L1) and
// scattered into KH rolling fp32 accumulators (slot r holds output row i + ph - r), so per 4-channel output there are
// KW loads, KH*KW*4 FMAs, one tanh.approx SiLU per channel and one 8-byte store.  Consecutive lanes cover consecutive
// channel groups, then consecutive columns: every warp access is one contiguous segment of the NHWC row.
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// WS: the KH*KW*4 weights of a thread do not fit the register file (7x7): they are staged in shared memory
// ([tap][C] fp32) and read back as 16-byte broadcasts.
template <int KH, int KW, bool WS>
__global__ void __launch_bounds__(256, (KH * KW <= 9) ? 3 : 1) dwconv_strip_kernel(ConvParams p, int R, int nstrips) {
  extern __shared__ float wsm_strip[];
  pdl_launch_dependents();
  if (WS) {
    for (int i = threadIdx.x; i < KH * KW * p.cout; i += blockDim.x) wsm_strip[i] = reinterpret_cast<const float *>(p.w)[i];
    __syncthreads();
  }
  const int cg = p.cout >> 2;
  const long long total = (long long)p.B * nstrips * p.Wo * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % cg) * 4;
  long long t = idx / cg;
  const int wo = (int)(t % p.Wo);
  t /= p.Wo;
  const int strip = (int)(t % nstrips);
  const int b = (int)(t / nstrips);
  float4 w[WS ? 1 : KH * KW];
  const float *wp = reinterpret_cast<const float *>(p.w) + c0;
  if (!WS) {
#pragma unroll
    for (int i = 0; i < KH * KW; ++i) w[i] = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)i * p.cout));
  }
  const float4 bias = __ldg(reinterpret_cast<const float4 *>(p.bias + c0));
  pdl_wait();  // weights are constants: only the activations depend on the preceding kernel
  const int ho0 = strip * R, ho1 = min(ho0 + R, p.Ho);
  float4 acc[KH];
#pragma unroll
  for (int r = 0; r < KH; ++r) acc[r] = bias;
  const bf16 *img = p.in + (size_t)b * p.H * p.W * p.in_cs + c0;
  // input rows are fetched one iteration ahead (raw 8-byte pixels, zeros outside the map) so that the loads of row i+1
  // are in flight while row i is being accumulated; the loop body itself is branch-free
  auto load_row = [&](int i, uint2(&buf)[KW]) {
    const bool row_in = (i >= 0 && i < p.H);
    const bf16 *row = img + (size_t)(row_in ? i : 0) * p.W * p.in_cs;
#pragma unroll
    for (int s_ = 0; s_ < KW; ++s_) {
      const int wi = wo - p.pw + s_;
      buf[s_] = (row_in && wi >= 0 && wi < p.W) ? __ldg(reinterpret_cast<const uint2 *>(row + (size_t)wi * p.in_cs))
                                                  : make_uint2(0u, 0u);
    }
  };
  const int i_end = ho1 - p.ph + KH - 1;
  uint2 nxt[KW];
  load_row(ho0 - p.ph, nxt);
  for (int i = ho0 - p.ph; i < i_end; ++i) {
    uint2 cur[KW];
#pragma unroll
    for (int s_ = 0; s_ < KW; ++s_) cur[s_] = nxt[s_];
    if (i + 1 < i_end) load_row(i + 1, nxt);
#pragma unroll
    for (int s_ = 0; s_ < KW; ++s_) {
      float x[4];
      unpack4(cur[s_], x);
#pragma unroll
      for (int r = 0; r < KH; ++r) {
        const float4 ww = WS ? *reinterpret_cast<const float4 *>(wsm_strip + (size_t)(r * KW + s_) * p.cout + c0) : w[r * KW + s_];
        acc[r].x = fmaf(x[0], ww.x, acc[r].x);
        acc[r].y = fmaf(x[1], ww.y, acc[r].y);
        acc[r].z = fmaf(x[2], ww.z, acc[r].z);
        acc[r].w = fmaf(x[3], ww.w, acc[r].w);
      }
    }
    const int ho = i + p.ph - (KH - 1);  // slot KH-1 is complete after input row i
    if (ho >= ho0) {
      float v[4] = {acc[KH - 1].x, acc[KH - 1].y, acc[KH - 1].z, acc[KH - 1].w};
      if (p.act) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = silu_tanh(v[j]);
      }
      const size_t opix = ((size_t)(b * p.Ho + ho) * p.Wo + wo);
      if (p.res_mode) {
        float rv[4];
        unpack4(__ldg(reinterpret_cast<const uint2 *>(p.res + opix * p.res_cs + c0)), rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (p.res_mode == B2Y_RES_ADD) ? v[j] + rv[j] : v[j] * rv[j];
      }
      *reinterpret_cast<uint2 *>(reinterpret_cast<bf16 *>(p.out) + opix * p.out_cs + c0) = pack4(v);
    }
#pragma unroll
    for (int r = KH - 1; r > 0; --r) acc[r] = acc[r - 1];
    acc[0] = bias;
  }
}

template <int KH, int KW, bool WS = false>
static int launch_dw_strip(const ConvParams &p, cudaStream_t st) {
  // strips of R output rows: long enough to amortise the weight loads and the KH-1 halo rows, short enough to fill the GPU
  const long long cols = (long long)p.B * p.Wo * (p.cout >> 2);
  int R = 16;
  while (R > 2 && cols * ceil_div(p.Ho, R) < 148LL * 1536) R >>= 1;
  const int nstrips = ceil_div(p.Ho, R);
  const long long total = cols * nstrips;
  const size_t smem = WS ? (size_t)KH * KW * p.cout * sizeof(float) : 0;
  if (smem > 48 * 1024)
    B2Y_CUDA_CHECK(cudaFuncSetAttribute(dwconv_strip_kernel<KH, KW, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  b2y_launch(dwconv_strip_kernel<KH, KW, WS>, dim3((unsigned)((total + 255) / 256)), dim3(256), smem, st, p, R, nstrips);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

int launch_dwconv(const ConvParams &p, cudaStream_t st) {
  B2Y_REQUIRE(p.cin == p.cout && !p.out_f32, B2Y_ERR_UNSUPPORTED, "depthwise conv needs cin == cout and bf16 output");
  B2Y_REQUIRE(p.res_mode != B2Y_RES_PREADD_UP2, B2Y_ERR_UNSUPPORTED, "depthwise conv: upsampled pre-activation residual");
  {
    static const int tile3 = [] { const char *e = getenv("B2Y_DW3_TILE"); return e ? atoi(e) : 1; }();  // A/B knob (default on)
    if (tile3 && dw3_tile_supported(p)) return launch_dw3_tile(p, st);
  }
  B2Y_REQUIRE(p.dw_chunk == 0, B2Y_ERR_UNSUPPORTED, "depthwise conv: a chunked input view needs the tiled 3x3 kernel (C %% 8, 3x3, pad 1)");
  auto aligned = [&](int v) {
    const uintptr_t m = (uintptr_t)v * 2 - 1;
    bool ok = (p.cout % v == 0) && (p.in_cs % v == 0) && (p.out_cs % v == 0) &&
              ((reinterpret_cast<uintptr_t>(p.in) & m) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & m) == 0);
    if (p.res_mode) ok = ok && (p.res_cs % v == 0) && ((reinterpret_cast<uintptr_t>(p.res) & m) == 0);
    return ok;
  };
  if (aligned(4) && p.sh == 1 && p.sw == 1 && p.Ho == p.H + 2 * p.ph - p.kh + 1) {
    if (p.kh == 3 && p.kw == 3) return launch_dw_strip<3, 3>(p, st);
    if (p.kh == 5 && p.kw == 5) return launch_dw_strip<5, 5>(p, st);
    if (p.kh == 1 && p.kw == 5) return launch_dw_strip<1, 5>(p, st);
    if (p.kh == 5 && p.kw == 1) return launch_dw_strip<5, 1>(p, st);
    if (p.kh == 1 && p.kw == 7) return launch_dw_strip<1, 7>(p, st);
    if (p.kh == 7 && p.kw == 1) return launch_dw_strip<7, 1>(p, st);
    if (p.kh == 1 && p.kw == 11) return launch_dw_strip<1, 11>(p, st);
    if (p.kh == 11 && p.kw == 1) return launch_dw_strip<11, 1>(p, st);
    const size_t wbytes = (size_t)p.kh * p.kw * p.cout * sizeof(float);
    if (p.kh == 7 && p.kw == 7 && wbytes <= 64 * 1024) return launch_dw_strip<7, 7, true>(p, st);
  }
  const size_t wsmem = (size_t)(p.kh * p.kw + 1) * p.cout * sizeof(float);
  // (the plain per-pixel kernel below was measured on these shapes too: 1.5x slower at 256 channels, stride 1 and 2)
  if (aligned(8) && wsmem <= 96 * 1024 && p.sh == p.sw) {
    const int wq = ceil_div(p.Wo, 4);
    const long long total = (long long)p.B * p.Ho * wq * (p.cout / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    auto launch = [&](auto kern) -> int {
      if (wsmem > 48 * 1024)
        B2Y_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
      b2y_launch(kern, dim3(grid), dim3(256), wsmem, st, p, wq);
      B2Y_LAUNCH_CHECK();
      return B2Y_OK;
    };
    if (p.kw == 3 && p.sw == 1) return launch(dwconv4_kernel<3, 1>);
    if (p.kw == 3 && p.sw == 2) return launch(dwconv4_kernel<3, 2>);
    if (p.kw == 5 && p.sw == 1) return launch(dwconv4_kernel<5, 1>);
    if (p.kw == 1 && p.sw == 1) return launch(dwconv4_kernel<1, 1>);
    return launch(dwconv4_kernel<0, 0>);
  }
  const int V = aligned(8) ? 8 : aligned(4) ? 4 : 1;
  const long long total = (long long)p.B * p.Ho * p.Wo * (p.cout / V);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (V == 8) b2y_launch(dwconv_kernel<8>, dim3(grid), dim3(256), 0, st, p);
  else if (V == 4) b2y_launch(dwconv_kernel<4>, dim3(grid), dim3(256), 0, st, p);
  else b2y_launch(dwconv_kernel<1>, dim3(grid), dim3(256), 0, st, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// CUDA-core convolution kernels: generic dense conv (validation twin of the tcgen05 kernel and the path for
// shapes that kernel does not take), the Cin=3 stem, and depthwise kh x kw.
// Reference semantics: ultralytics/nn/modules/conv.py:145-151 (act(conv(x)+b) with BN folded), DWConv :224-229.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ dense, generic
// thread = 1 output pixel x 8 output channels; weights [cout_pad][kh*kw*cin] bf16 are warp-uniform (broadcast loads).
template <bool VEC>
__global__ void __launch_bounds__(128) conv_simt_kernel(ConvParams p) {
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
    if (p.act) v = silu_f(v);
    if (p.res_mode) {
      const float r = __bfloat162float(p.res[opix * p.res_cs + co]);
      v = (p.res_mode == B2Y_RES_ADD) ? v + r : v * r;
    }
    if (p.out_f32) reinterpret_cast<float *>(p.out)[opix * p.out_cs + co] = v;
    else reinterpret_cast<bf16 *>(p.out)[opix * p.out_cs + co] = __float2bfloat16_rn(v);
  }
}

int launch_conv_simt(const ConvParams &p, cudaStream_t st) {
  const long long npix = (long long)p.B * p.Ho * p.Wo;
  dim3 grid((unsigned)((npix + 127) / 128), (unsigned)ceil_div(p.cout, 8));
  const bool vec = (p.cin % 8 == 0) && (p.in_cs % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.w) & 15) == 0);
  if (vec) conv_simt_kernel<true><<<grid, 128, 0, st>>>(p);
  else conv_simt_kernel<false><<<grid, 128, 0, st>>>(p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------ stem (Cin = 3)
// Input: NHWC bf16 with 4 channels per pixel (R,G,B,0) as written by the preprocess kernel (8-byte pixels).
// thread = 1 output pixel; the 27 inputs stay in registers while output channels are produced 16 at a time.
// weights fp32 [cout][27] (tap-major, channel fastest) staged in shared memory.
__global__ void __launch_bounds__(128) stem_kernel(ConvParams p) {
  extern __shared__ float sw[];  // cout*27 weights + cout bias
  const int nw = p.cout * 27;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = reinterpret_cast<const float *>(p.w)[i];
  for (int i = threadIdx.x; i < p.cout; i += blockDim.x) sw[nw + i] = p.bias[i];
  __syncthreads();
  const long long npix = (long long)p.B * p.Ho * p.Wo;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int wo = (int)(pix % p.Wo);
  const int ho = (int)((pix / p.Wo) % p.Ho);
  const int b = (int)(pix / ((long long)p.Wo * p.Ho));
  float x[27];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = ho * 2 - 1 + r;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int wi = wo * 2 - 1 + s;
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
        unpack4(*reinterpret_cast<const uint2 *>(p.in + ((size_t)(b * p.H + hi) * p.W + wi) * 4), f);
      x[(r * 3 + s) * 3 + 0] = f[0];
      x[(r * 3 + s) * 3 + 1] = f[1];
      x[(r * 3 + s) * 3 + 2] = f[2];
    }
  }
  bf16 *o = reinterpret_cast<bf16 *>(p.out) + (size_t)pix * p.out_cs;
  for (int c0 = 0; c0 < p.cout; c0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float *wj = sw + (c0 + j) * 27;
      float a = sw[nw + c0 + j];
#pragma unroll
      for (int i = 0; i < 27; ++i) a = fmaf(x[i], wj[i], a);
      v[j] = p.act ? silu_f(a) : a;
    }
    *reinterpret_cast<uint4 *>(o + c0) = pack8(v);
  }
}

int launch_stem(const ConvParams &p, cudaStream_t st) {
  B2Y_REQUIRE(p.cin == 3 && p.kh == 3 && p.kw == 3 && p.sh == 2 && p.sw == 2 && p.ph == 1 && p.pw == 1 &&
                  p.in_cs == 4 && p.cout % 8 == 0 && p.out_cs % 8 == 0 && !p.out_f32 && p.res_mode == 0,
              B2Y_ERR_UNSUPPORTED, "stem kernel: unsupported configuration");
  const long long npix = (long long)p.B * p.Ho * p.Wo;
  const size_t smem = (size_t)p.cout * 28 * sizeof(float);
  stem_kernel<<<(unsigned)((npix + 127) / 128), 128, smem, st>>>(p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------ depthwise
// thread = 1 output pixel x V channels (V = 8: 16-byte, 4: 8-byte, 1: scalar accesses); weights fp32 [kh*kw][c].
template <int V>
__global__ void __launch_bounds__(256) dwconv_kernel(ConvParams p) {
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

int launch_dwconv(const ConvParams &p, cudaStream_t st) {
  B2Y_REQUIRE(p.cin == p.cout && !p.out_f32, B2Y_ERR_UNSUPPORTED, "depthwise conv needs cin == cout and bf16 output");
  auto aligned = [&](int v) {
    const uintptr_t m = (uintptr_t)v * 2 - 1;
    bool ok = (p.cout % v == 0) && (p.in_cs % v == 0) && (p.out_cs % v == 0) &&
              ((reinterpret_cast<uintptr_t>(p.in) & m) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & m) == 0);
    if (p.res_mode) ok = ok && (p.res_cs % v == 0) && ((reinterpret_cast<uintptr_t>(p.res) & m) == 0);
    return ok;
  };
  const int V = aligned(8) ? 8 : aligned(4) ? 4 : 1;
  const long long total = (long long)p.B * p.Ho * p.Wo * (p.cout / V);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (V == 8) dwconv_kernel<8><<<grid, 256, 0, st>>>(p);
  else if (V == 4) dwconv_kernel<4><<<grid, 256, 0, st>>>(p);
  else dwconv_kernel<1><<<grid, 256, 0, st>>>(p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

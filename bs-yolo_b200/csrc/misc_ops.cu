// Memory-bound glue ops: SPPF chained max-pools, nearest-2x upsample / channel-slice copy, and the
// reductions / gates / elementwise pieces of the BS-YOLO add-on modules (ELA, MSCAAttention).
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ SPPF pools
// block = (image, 8-channel group); the H x W x 8 tile lives in shared memory and the three chained
// MaxPool2d(5,1,2) (ultralytics/nn/modules/block.py:3127,3147-3148; -inf padding) run as separable row/column
// passes, each stage written to its own channel slice of the concat buffer: out[:, j*c + ch], j = 0..2.
__global__ void __launch_bounds__(256) sppf_pool_kernel(const bf16 *__restrict__ in, int in_cs, bf16 *__restrict__ out,
                                                        int out_cs, int c, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint4 tile[];  // [2][H*W] of 8 x bf16
  uint4 *cur = tile, *tmp = tile + H * W;
  const int ngrp = c / 8;
  const int b = blockIdx.x / ngrp, g = blockIdx.x % ngrp;
  const size_t img = (size_t)b * H * W;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x)
    cur[i] = *reinterpret_cast<const uint4 *>(in + (img + i) * in_cs + g * 8);
  __syncthreads();
  for (int stage = 0; stage < 3; ++stage) {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {  // row pass: max over w-2..w+2
      const int h = i / W, w = i % W;
      __nv_bfloat162 m[4];
      const __nv_bfloat162 *c0 = reinterpret_cast<const __nv_bfloat162 *>(&cur[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = c0[k];
      for (int d = -2; d <= 2; ++d) {
        const int ww = w + d;
        if (d == 0 || ww < 0 || ww >= W) continue;
        const __nv_bfloat162 *v = reinterpret_cast<const __nv_bfloat162 *>(&cur[h * W + ww]);
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], v[k]);
      }
      tmp[i] = *reinterpret_cast<uint4 *>(m);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {  // column pass: max over h-2..h+2
      const int h = i / W, w = i % W;
      __nv_bfloat162 m[4];
      const __nv_bfloat162 *c0 = reinterpret_cast<const __nv_bfloat162 *>(&tmp[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = c0[k];
      for (int d = -2; d <= 2; ++d) {
        const int hh = h + d;
        if (d == 0 || hh < 0 || hh >= H) continue;
        const __nv_bfloat162 *v = reinterpret_cast<const __nv_bfloat162 *>(&tmp[hh * W + w]);
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], v[k]);
      }
      const uint4 r = *reinterpret_cast<uint4 *>(m);
      *reinterpret_cast<uint4 *>(out + (img + i) * out_cs + stage * c + g * 8) = r;
      cur[i] = r;  // safe: this pass only reads tmp
    }
    __syncthreads();
  }
}

int launch_sppf_pool(const bf16 *in, int in_cs, bf16 *out, int out_cs, int c, int B, int H, int W, cudaStream_t st) {
  B2Y_REQUIRE(c % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0, B2Y_ERR_UNSUPPORTED, "sppf pool: channels must be x8");
  const size_t smem = (size_t)2 * H * W * sizeof(uint4);
  B2Y_REQUIRE(smem <= 200 * 1024, B2Y_ERR_UNSUPPORTED, "sppf pool: %dx%d map does not fit shared memory", H, W);
  if (smem > 48 * 1024)
    B2Y_CUDA_CHECK(cudaFuncSetAttribute(sppf_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  b2y_launch(sppf_pool_kernel, dim3(B * (c / 8)), dim3(256), smem, st, in, in_cs, out, out_cs, c, H, W);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------ resample / copy
// out[b, h, w, :c] = in[b, h/scale, w/scale, :c]  (scale 1: concat copy; scale 2: nn.Upsample(2,"nearest")).
template <int V>
__global__ void __launch_bounds__(256) resample_kernel(const bf16 *__restrict__ in, int in_cs, bf16 *__restrict__ out,
                                                       int out_cs, int c, long long total, int Ho, int Wo, int scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = c / V;
  const int c0 = (int)(idx % cg) * V;
  const long long pix = idx / cg;
  const int wo = (int)(pix % Wo);
  const int ho = (int)((pix / Wo) % Ho);
  const long long b = pix / ((long long)Wo * Ho);
  const int Hi = Ho / scale, Wi = Wo / scale;
  const bf16 *s = in + ((size_t)(b * Hi + ho / scale) * Wi + wo / scale) * in_cs + c0;
  bf16 *d = out + (size_t)pix * out_cs + c0;
  if (V == 8) *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
  else if (V == 4) *reinterpret_cast<uint2 *>(d) = *reinterpret_cast<const uint2 *>(s);
  else d[0] = s[0];
}

int launch_resample(const bf16 *in, int in_cs, bf16 *out, int out_cs, int c, int B, int H, int W, int scale,
                    cudaStream_t st) {
  // H, W are OUTPUT dims
  auto ok = [&](int v) {
    const uintptr_t m = (uintptr_t)v * 2 - 1;
    return c % v == 0 && in_cs % v == 0 && out_cs % v == 0 && ((uintptr_t)in & m) == 0 && ((uintptr_t)out & m) == 0;
  };
  const int V = ok(8) ? 8 : ok(4) ? 4 : 1;
  const long long total = (long long)B * H * W * (c / V);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (V == 8) b2y_launch(resample_kernel<8>, dim3(grid), dim3(256), 0, st, in, in_cs, out, out_cs, c, total, H, W, scale);
  else if (V == 4) b2y_launch(resample_kernel<4>, dim3(grid), dim3(256), 0, st, in, in_cs, out, out_cs, c, total, H, W, scale);
  else b2y_launch(resample_kernel<1>, dim3(grid), dim3(256), 0, st, in, in_cs, out, out_cs, c, total, H, W, scale);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------ ELA
// rowm[b,h,c] = mean_w x[b,h,w,c]; colm[b,w,c] = mean_h x[b,h,w,c]   (ELA.py:74-84).  thread = channel.
__global__ void rowcol_mean_kernel(const bf16 *__restrict__ x, int cs, int c, int H, int W, float *__restrict__ rowm,
                                   float *__restrict__ colm) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, ch = blockIdx.y * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const int line = blockIdx.x;  // [0,H): row means; [H,H+W): column means
  const bf16 *base = x + (size_t)b * H * W * cs + ch;
  float s = 0.f;
  if (line < H) {
    for (int w = 0; w < W; ++w) s += __bfloat162float(base[(size_t)(line * W + w) * cs]);
    rowm[((size_t)b * H + line) * c + ch] = s / (float)W;
  } else {
    const int w = line - H;
    for (int h = 0; h < H; ++h) s += __bfloat162float(base[(size_t)(h * W + w) * cs]);
    colm[((size_t)b * W + w) * c + ch] = s / (float)H;
  }
}

int launch_rowcol_mean(const bf16 *x, int cs, int c, int B, int H, int W, float *rowm, float *colm, cudaStream_t st) {
  const int bt = c >= 128 ? 128 : 64;
  dim3 grid(H + W, ceil_div(c, bt), B);
  b2y_launch(rowcol_mean_kernel, dim3(grid), dim3(bt), 0, st, x, cs, c, H, W, rowm, colm);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// One block per (image, GroupNorm group); does both directions.  params: [sp_w (c*ks) | gn_w (c) | gn_b (c) | ch_w (c*ks)].
//   gate_dir[b,l,ch] = sigmoid(GN(conv1d_dil2(mean_dir)[ch]))  (ELA.py:77-86), GroupNorm over (c/groups channels x L)
//   chgate[b,ch]     = sigmoid(ch_w[ch][centre] * mean(x[b,:,:,ch]))  (ELA.py:49-58: Conv1d on a length-1 sequence)
__global__ void __launch_bounds__(256) ela_gate_kernel(const float *__restrict__ rowm, const float *__restrict__ colm,
                                                       const float *__restrict__ params, int c, int H, int W, int ks,
                                                       int groups, float eps, float *__restrict__ hgate,
                                                       float *__restrict__ wgate, float *__restrict__ chgate) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // conv outputs for this group: [cpg][L]
  __shared__ float red[64];
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = c / groups;
  const float *sp_w = params, *gn_w = params + (size_t)c * ks, *gn_b = gn_w + c, *ch_w = gn_b + c;
  for (int dir = 0; dir < 2; ++dir) {
    const int L = dir == 0 ? H : W;
    const float *src = (dir == 0 ? rowm : colm) + (size_t)b * L * c;
    float *dst = (dir == 0 ? hgate : wgate) + (size_t)b * L * c;
    const int n = cpg * L;
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int ci = i % cpg, l = i / cpg, ch = g * cpg + ci;
      float a = 0.f;
      for (int t = 0; t < ks; ++t) {
        const int ll = l + 2 * t - (ks - 1);  // dilation 2, padding ks-1
        if (ll >= 0 && ll < L) a = fmaf(sp_w[ch * ks + t], src[(size_t)ll * c + ch], a);
      }
      sm[i] = a;
      s1 += a;
      s2 += a * a;
    }
    // block reduce (sum, sum of squares)
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      red[threadIdx.x >> 5] = s1;
      red[32 + (threadIdx.x >> 5)] = s2;
    }
    __syncthreads();
    float t1 = 0.f, t2 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      t1 += red[w];
      t2 += red[32 + w];
    }
    const float mean = t1 / (float)n;
    const float var = fmaxf(t2 / (float)n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int ci = i % cpg, l = i / cpg, ch = g * cpg + ci;
      const float v = (sm[i] - mean) * rstd * gn_w[ch] + gn_b[ch];
      dst[(size_t)l * c + ch] = 1.f / (1.f + expf(-v));
    }
    __syncthreads();
  }
  for (int ci = threadIdx.x; ci < cpg; ci += blockDim.x) {
    const int ch = g * cpg + ci;
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += rowm[((size_t)b * H + h) * c + ch];
    const float gm = s / (float)H;
    chgate[(size_t)b * c + ch] = 1.f / (1.f + expf(-ch_w[ch * ks + (ks - 1) / 2] * gm));
  }
}

int launch_ela_gate(const float *rowm, const float *colm, const float *params, int c, int B, int H, int W, int ks,
                    int groups, float eps, float *hgate, float *wgate, float *chgate, cudaStream_t st) {
  B2Y_REQUIRE(c % groups == 0, B2Y_ERR_INVALID, "ELA: channels %d not divisible by groups %d", c, groups);
  const int L = H > W ? H : W;
  const size_t smem = (size_t)(c / groups) * L * sizeof(float);
  B2Y_REQUIRE(smem <= 48 * 1024, B2Y_ERR_UNSUPPORTED, "ELA gate: group tile too large");
  b2y_launch(ela_gate_kernel, dim3(B * groups), dim3(256), smem, st, rowm, colm, params, c, H, W, ks, groups, eps, hgate, wgate, chgate);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// out = x * (ch_c*chgate[b,c] + sp_c*hgate[b,h,c]*wgate[b,w,c]) + res_c*x   (ELA.py:94-100)
__global__ void __launch_bounds__(256) ela_apply_kernel(const bf16 *__restrict__ x, int cs, bf16 *__restrict__ out,
                                                        int out_cs, int c, long long total, int H, int W,
                                                        const float *__restrict__ hgate, const float *__restrict__ wgate,
                                                        const float *__restrict__ chgate, float ch_c, float sp_c,
                                                        float res_c) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = c / 8;
  const int c0 = (int)(idx % cg) * 8;
  const long long pix = idx / cg;
  const int w = (int)(pix % W);
  const int h = (int)((pix / W) % H);
  const long long b = pix / ((long long)W * H);
  float v[8];
  unpack8(*reinterpret_cast<const uint4 *>(x + (size_t)pix * cs + c0), v);
  const float *hg = hgate + ((size_t)b * H + h) * c + c0;
  const float *wg = wgate + ((size_t)b * W + w) * c + c0;
  const float *cgp = chgate + (size_t)b * c + c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = v[j] * (ch_c * cgp[j] + sp_c * (hg[j] * wg[j])) + res_c * v[j];
  *reinterpret_cast<uint4 *>(out + (size_t)pix * out_cs + c0) = pack8(v);
}

int launch_ela_apply(const bf16 *x, int cs, bf16 *out, int out_cs, int c, int B, int H, int W, const float *hgate,
                     const float *wgate, const float *chgate, float ch_c, float sp_c, float res_c, cudaStream_t st) {
  B2Y_REQUIRE(c % 8 == 0 && cs % 8 == 0 && out_cs % 8 == 0, B2Y_ERR_UNSUPPORTED, "ELA apply: channels must be x8");
  const long long total = (long long)B * H * W * (c / 8);
  b2y_launch(ela_apply_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, cs, out, out_cs, c, total, H, W, hgate, wgate,
                                                                    chgate, ch_c, sp_c, res_c);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------ MSCA
struct Ptr4 {
  const bf16 *p[4];
  int cs[4];
};

// gap[b, j, ch] = mean over pixels of branch j (MSCA.py:71-74).  block = (image, branch), thread = channel.
__global__ void gap4_kernel(Ptr4 br, int c, int HW, float *__restrict__ gap) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, j = blockIdx.y;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const bf16 *x = br.p[j] + (size_t)b * HW * br.cs[j] + ch;
    float s = 0.f;
    for (int i = 0; i < HW; ++i) s += __bfloat162float(x[(size_t)i * br.cs[j]]);
    gap[((size_t)b * 4 + j) * c + ch] = s / (float)HW;
  }
}

int launch_gap4(const bf16 *const *br, const int *br_cs, int c, int B, int HW, float *gap, cudaStream_t st) {
  Ptr4 q;
  for (int j = 0; j < 4; ++j) {
    q.p[j] = br[j];
    q.cs[j] = br_cs[j];
  }
  b2y_launch(gap4_kernel, dim3(dim3(B, 4)), dim3(256), 0, st, q, c, HW, gap);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// gate[b,j,ch] = softmax_j(sigmoid(SE_j(gap_j)[ch]))   (MSCA.py:71-77); se_w fp32 [4][c][c], se_b [4][c].
// block = (image, 32 output channels): one warp per output channel does the 4 dot products.
__global__ void __launch_bounds__(256) msca_gate_kernel(const float *__restrict__ gap, const float *__restrict__ se_w,
                                                        const float *__restrict__ se_b, int c, float *__restrict__ gate) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ch = blockIdx.y * 32 + warp; ch < min(c, (int)blockIdx.y * 32 + 32); ch += 8) {
    float z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float *w = se_w + ((size_t)j * c + ch) * c;
      const float *g = gap + ((size_t)b * 4 + j) * c;
      float s = 0.f;
      for (int k = lane; k < c; k += 32) s = fmaf(w[k], g[k], s);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      z[j] = 1.f / (1.f + expf(-(s + se_b[j * c + ch])));
    }
    if (lane == 0) {
      const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
      float e[4], t = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        e[j] = expf(z[j] - m);
        t += e[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) gate[((size_t)b * 4 + j) * c + ch] = e[j] / t;
    }
  }
}

int launch_msca_gate(const float *gap, const float *se_w, const float *se_b, int c, int B, float *gate,
                     cudaStream_t st) {
  b2y_launch(msca_gate_kernel, dim3(dim3(B, ceil_div(c, 32))), dim3(256), 0, st, gap, se_w, se_b, c, gate);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// out[b,p,ch] = sum_j gate[b,j,ch] * br_j[b,p,ch]   (MSCA.py:84)
__global__ void __launch_bounds__(256) wsum4_kernel(Ptr4 br, const float *__restrict__ gate, bf16 *__restrict__ out,
                                                    int out_cs, int c, long long total, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = c / 8;
  const int c0 = (int)(idx % cg) * 8;
  const long long pix = idx / cg;
  const long long b = pix / HW;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4 *>(br.p[j] + (size_t)pix * br.cs[j] + c0), v);
    const float *g = gate + ((size_t)b * 4 + j) * c + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(g[k], v[k], acc[k]);
  }
  *reinterpret_cast<uint4 *>(out + (size_t)pix * out_cs + c0) = pack8(acc);
}

int launch_wsum4(const bf16 *const *br, const int *br_cs, const float *gate, bf16 *out, int out_cs, int c, int B,
                 int HW, cudaStream_t st) {
  B2Y_REQUIRE(c % 8 == 0 && out_cs % 8 == 0, B2Y_ERR_UNSUPPORTED, "wsum4: channels must be x8");
  Ptr4 q;
  for (int j = 0; j < 4; ++j) {
    q.p[j] = br[j];
    q.cs[j] = br_cs[j];
    B2Y_REQUIRE(br_cs[j] % 8 == 0, B2Y_ERR_UNSUPPORTED, "wsum4: branch stride must be x8");
  }
  const long long total = (long long)B * HW * (c / 8);
  b2y_launch(wsum4_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, q, gate, out, out_cs, c, total, HW);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

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

// nearest 2x upsample, 16-byte chunks: thread = one INPUT chunk, read once and stored to its 2 x 2 output pixels
// (the per-output kernel above reads every input pixel four times).
__global__ void __launch_bounds__(256) upsample2_kernel(const bf16 *__restrict__ in, int in_cs, bf16 *__restrict__ out,
                                                        int out_cs, int c, long long total, int Hi, int Wi) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = c >> 3;
  const int c0 = (int)(idx % cg) * 8;
  const long long pix = idx / cg;
  const int wi = (int)(pix % Wi);
  const int hi = (int)((pix / Wi) % Hi);
  const long long b = pix / ((long long)Wi * Hi);
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + (size_t)pix * in_cs + c0));
  bf16 *d = out + ((size_t)(b * 2 * Hi + 2 * hi) * (2 * Wi) + 2 * wi) * out_cs + c0;
  *reinterpret_cast<uint4 *>(d) = v;
  *reinterpret_cast<uint4 *>(d + out_cs) = v;
  *reinterpret_cast<uint4 *>(d + (size_t)2 * Wi * out_cs) = v;
  *reinterpret_cast<uint4 *>(d + (size_t)2 * Wi * out_cs + out_cs) = v;
}

int launch_resample(const bf16 *in, int in_cs, bf16 *out, int out_cs, int c, int B, int H, int W, int scale,
                    cudaStream_t st) {
  // H, W are OUTPUT dims
  if (scale == 2 && c % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
      H % 2 == 0 && W % 2 == 0) {
    const long long total = (long long)B * (H / 2) * (W / 2) * (c / 8);
    b2y_launch(upsample2_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, in, in_cs, out, out_cs, c, total, H / 2, W / 2);
    B2Y_LAUNCH_CHECK();
    return B2Y_OK;
  }
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

// Banded single-pass variant (the fast path; C in {64,128,256}, W*C/8 <= 1024): one CTA = 8 rows of one image, read
// ONCE with fully coalesced 16-byte loads.  Thread t owns the 16-byte chunks t + 256 j of every row (chunk k = pixel
// k / (C/8), channel octet k % (C/8); 256 % (C/8) == 0, so a thread's octet and columns are the same in every row):
// column sums accumulate in registers, row sums are reduced by warp shuffles over the lanes that share an octet and then
// across the 8 warps through shared memory in a fixed order (deterministic).  Rows are fetched two at a time with no
// barrier in between (one barrier pair per `rs` rows), three CTAs per SM: ~100 KB of loads in flight per SM.  The column
// sums of the band are written as partial sums colm[b][w][band][c] (already divided by H); ela_gate adds the bands.
constexpr int ELA_BAND = 8, ELA_NT = 256;

template <int NJ>
__global__ void __launch_bounds__(ELA_NT, 3) ela_stats_kernel(const bf16 *__restrict__ x, int cs, int c, int H, int W, int nbands,
                                                              int rs_rows, float *__restrict__ rowm, float *__restrict__ colm) {
  pdl_launch_dependents();
  extern __shared__ float red[];  // [rs_rows][8 warps][c]
  const int b = blockIdx.x / nbands, band = blockIdx.x % nbands;
  const int n_oct = c >> 3, cpr = W * n_oct;  // 16-byte chunks per row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int o = tid % n_oct;
  const int h0 = band * ELA_BAND, h1 = min(h0 + ELA_BAND, H);
  bool has[NJ];
  int wq[NJ];
  float cs_acc[NJ][8];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    has[j] = tid + j * ELA_NT < cpr;
    wq[j] = (tid + j * ELA_NT) / n_oct;
#pragma unroll
    for (int i = 0; i < 8; ++i) cs_acc[j][i] = 0.f;
  }
  const bf16 *img = x + (size_t)b * H * W * cs + o * 8;
  const bf16 *pj[NJ];  // the thread's chunks in row 0 of the image; rows are `rowstride` elements apart (32-bit offsets)
#pragma unroll
  for (int j = 0; j < NJ; ++j) pj[j] = img + (has[j] ? wq[j] : 0) * cs;
  const int rowstride = W * cs;
  pdl_wait();
  for (int g0 = h0; g0 < h1; g0 += rs_rows) {
    for (int r0 = g0; r0 < min(g0 + rs_rows, h1); r0 += 2) {
      uint4 v[2][NJ];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int off = (r0 + r) * rowstride;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          v[r][j] = (r0 + r < h1 && has[j]) ? __ldg(reinterpret_cast<const uint4 *>(pj[j] + off)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float rs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float f[8];
          unpack8(v[r][j], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            rs[i] += f[i];
            cs_acc[j][i] += f[i];
          }
        }
        for (int st = n_oct; st < 32; st <<= 1)  // lanes l, l ^ st share the octet (n_oct is a power of two <= 32)
#pragma unroll
          for (int i = 0; i < 8; ++i) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], st);
        if (lane < n_oct) {
          float *d = red + ((size_t)(r0 + r - g0) * (ELA_NT / 32) + warp) * c + lane * 8;
          *reinterpret_cast<float4 *>(d) = make_float4(rs[0], rs[1], rs[2], rs[3]);
          *reinterpret_cast<float4 *>(d + 4) = make_float4(rs[4], rs[5], rs[6], rs[7]);
        }
      }
    }
    __syncthreads();
    const int nr = min(rs_rows, h1 - g0);
    for (int i = tid; i < nr * c; i += ELA_NT) {
      const int r = i / c, ch = i % c;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < ELA_NT / 32; ++w) s += red[((size_t)r * (ELA_NT / 32) + w) * c + ch];
      rowm[((size_t)b * H + g0 + r) * c + ch] = s / (float)W;
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if (has[j]) {
      float *d = colm + (((size_t)b * W + wq[j]) * nbands + band) * c + o * 8;
      const float inv = 1.f / (float)H;
      *reinterpret_cast<float4 *>(d) = make_float4(cs_acc[j][0] / (float)H, cs_acc[j][1] / (float)H, cs_acc[j][2] / (float)H, cs_acc[j][3] / (float)H);
      *reinterpret_cast<float4 *>(d + 4) = make_float4(cs_acc[j][4] / (float)H, cs_acc[j][5] / (float)H, cs_acc[j][6] / (float)H, cs_acc[j][7] / (float)H);
      (void)inv;
    }
}

bool ela_fast_path(int c, int W, int cs, const void *x) {
  return (c == 64 || c == 128 || c == 256) && W * (c >> 3) <= 4 * ELA_NT && cs % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

int launch_rowcol_mean(const bf16 *x, int cs, int c, int B, int H, int W, int nbands, float *rowm, float *colm, cudaStream_t st) {
  if (nbands >= 1) {  // the planner asked for the banded layout of colm (it applies the same shape test)
    B2Y_REQUIRE(ela_fast_path(c, W, cs, x) && nbands == ceil_div(H, ELA_BAND), B2Y_ERR_INVALID,
                "ELA stats: banded layout requested for an unsupported shape (c %d, W %d, bands %d)", c, W, nbands);
    const int rs_rows = c <= 128 ? ELA_BAND : 4;  // rows between barrier pairs: 16 / 32 / 32 KB of partial row sums
    const size_t smem = (size_t)rs_rows * (ELA_NT / 32) * c * sizeof(float);
    const int nj = ceil_div(W * (c >> 3), ELA_NT);
    auto go = [&](auto kern) -> int {
      b2y_launch(kern, dim3(B * nbands), dim3(ELA_NT), smem, st, x, cs, c, H, W, nbands, rs_rows, rowm, colm);
      B2Y_LAUNCH_CHECK();
      return B2Y_OK;
    };
    return nj <= 1 ? go(ela_stats_kernel<1>) : nj == 2 ? go(ela_stats_kernel<2>) : nj == 3 ? go(ela_stats_kernel<3>)
                                                                                            : go(ela_stats_kernel<4>);
  }
  const int bt = c >= 128 ? 128 : 64;
  dim3 grid(H + W, ceil_div(c, bt), B);
  b2y_launch(rowcol_mean_kernel, dim3(grid), dim3(bt), 0, st, x, cs, c, H, W, rowm, colm);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// One 128-thread block per (image, GroupNorm group); does both directions.
// params: [sp_w (c*ks) | gn_w (c) | gn_b (c) | ch_w (c*ks)].
//   gate_dir[b,l,ch] = sigmoid(GN(conv1d_dil2(mean_dir)[ch]))  (ELA.py:77-86), GroupNorm over (c/groups channels x L)
//   chgate[b,ch]     = sigmoid(ch_w[ch][centre] * mean(x[b,:,:,ch]))  (ELA.py:49-58: Conv1d on a length-1 sequence)
// Column means arrive as per-band partial sums (colm[b][w][band][c]): the band loads of an element are issued together
// (unrolled, predicated) and added in band order.  The kernel is pure latency: every phase is written so that a thread has
// all its global loads in flight at once (the first version's serial loops -- channel gate over H rows, bands one by one --
// cost 25-100 us per call for ~1 MFLOP).
constexpr int GATE_NT = 128, GATE_MAXNB = 16;
__global__ void __launch_bounds__(GATE_NT) ela_gate_kernel(const float *__restrict__ rowm, const float *__restrict__ colm,
                                                           const float *__restrict__ params, int c, int H, int W, int ks,
                                                           int groups, float eps, int nbands, float *__restrict__ hgate,
                                                           float *__restrict__ wgate, float *__restrict__ chgate) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // conv outputs for this group: [cpg][Lmax], then the group's means [Lmax][cpg]
  __shared__ float red[2 * (GATE_NT / 32)];
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = c / groups;
  const int cpg_sh = (cpg & (cpg - 1)) == 0 ? 31 - __clz(cpg) : -1;  // power of two: shifts instead of divisions
  const float *sp_w = params, *gn_w = params + (size_t)c * ks, *gn_b = gn_w + c, *ch_w = gn_b + c;
  float *mean_s = sm + cpg * (H > W ? H : W);
  auto split = [&](int i, int &ci, int &l) {
    if (cpg_sh >= 0) { ci = i & (cpg - 1); l = i >> cpg_sh; }
    else { ci = i % cpg; l = i / cpg; }
  };
  for (int dir = 0; dir < 2; ++dir) {
    const int L = dir == 0 ? H : W;
    const int nb = dir == 0 ? 1 : nbands;
    const float *src = (dir == 0 ? rowm : colm) + (size_t)b * L * nb * c;
    float *dst = (dir == 0 ? hgate : wgate) + (size_t)b * L * c;
    const int n = cpg * L;
    for (int i = threadIdx.x; i < n; i += GATE_NT) {
      int ci, l;
      split(i, ci, l);
      const int ch = g * cpg + ci;
      const float *sp = src + l * nb * c + ch;  // (32-bit offsets: an image's slice is < 2^31 elements)
      float m = 0.f;
      int k = 0;
      for (; k + 4 <= nb; k += 4) {  // four band loads in flight, added in band order
        const float p0 = __ldg(sp), p1 = __ldg(sp + c), p2 = __ldg(sp + 2 * c), p3 = __ldg(sp + 3 * c);
        sp += 4 * c;
        m += p0; m += p1; m += p2; m += p3;
      }
      for (; k < nb; ++k, sp += c) m += __ldg(sp);
      mean_s[i] = m;
    }
    __syncthreads();
    if (dir == 0) {  // channel gate from the row means just staged: one warp-wide strided sum per channel
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (int ci = warp; ci < cpg; ci += GATE_NT / 32) {
        const int ch = g * cpg + ci;
        float s = 0.f;
        for (int h = lane; h < H; h += 32) s += mean_s[h * cpg + ci];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) chgate[(size_t)b * c + ch] = __fdividef(1.f, 1.f + __expf(-ch_w[ch * ks + (ks - 1) / 2] * (s / (float)H)));
      }
    }
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < n; i += GATE_NT) {
      int ci, l;
      split(i, ci, l);
      const int ch = g * cpg + ci;
      float a = 0.f;
      const float *wt = sp_w + ch * ks;
      const float *ms = mean_s + ci + (l - (ks - 1)) * cpg;  // dilation 2, padding ks-1: tap t reads row l + 2t - (ks-1)
      for (int t = 0; t < ks; ++t) {
        const int ll = l + 2 * t - (ks - 1);
        if (ll >= 0 && ll < L) a = fmaf(__ldg(wt + t), ms[2 * t * cpg], a);
      }
      sm[i] = a;
      s1 += a;
      s2 += a * a;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
      red[threadIdx.x >> 5] = s1;
      red[GATE_NT / 32 + (threadIdx.x >> 5)] = s2;
    }
    __syncthreads();
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < GATE_NT / 32; ++w) {
      t1 += red[w];
      t2 += red[GATE_NT / 32 + w];
    }
    const float mean = t1 / (float)n;
    const float var = fmaxf(t2 / (float)n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    for (int i = threadIdx.x; i < n; i += GATE_NT) {
      int ci, l;
      split(i, ci, l);
      const int ch = g * cpg + ci;
      const float v = (sm[i] - mean) * rstd * __ldg(gn_w + ch) + __ldg(gn_b + ch);
      dst[l * c + ch] = __fdividef(1.f, 1.f + __expf(-v));
    }
    __syncthreads();
  }
}

int launch_ela_gate(const float *rowm, const float *colm, const float *params, int c, int B, int H, int W, int ks,
                    int groups, float eps, int nbands, float *hgate, float *wgate, float *chgate, cudaStream_t st) {
  B2Y_REQUIRE(c % groups == 0, B2Y_ERR_INVALID, "ELA: channels %d not divisible by groups %d", c, groups);
  B2Y_REQUIRE(nbands <= GATE_MAXNB, B2Y_ERR_UNSUPPORTED, "ELA gate: %d column bands (max %d)", nbands, GATE_MAXNB);
  const int L = H > W ? H : W;
  const size_t smem = (size_t)(c / groups) * L * sizeof(float) * 2;
  B2Y_REQUIRE(smem <= 48 * 1024, B2Y_ERR_UNSUPPORTED, "ELA gate: group tile too large");
  b2y_launch(ela_gate_kernel, dim3(B * groups), dim3(GATE_NT), smem, st, rowm, colm, params, c, H, W, ks, groups, eps,
             nbands < 1 ? 1 : nbands, hgate, wgate, chgate);
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

// Fast path (same shapes as ela_stats_kernel): CTA = 8 rows of one image, thread = the 16-byte chunks t + 256 j of every
// row.  A thread's column gate and channel gate are loop invariants held in registers; per row it reads one 32-byte row
// gate (shared by all threads of its octet: L1 hits), so the kernel moves 16 B in + 16 B out per chunk and little else --
// the per-pixel version above re-read 96 B of gates for every 16 B of x.  Four rows of loads are in flight per thread,
// two CTAs per SM.
template <int NJ>
__global__ void __launch_bounds__(ELA_NT, 2) ela_apply_band_kernel(const bf16 *__restrict__ x, int cs, bf16 *__restrict__ out,
                                                                   int out_cs, int c, int H, int W, int nbands,
                                                                   const float *__restrict__ hgate, const float *__restrict__ wgate,
                                                                   const float *__restrict__ chgate, float ch_c, float sp_c,
                                                                   float res_c) {
  pdl_launch_dependents();
  const int b = blockIdx.x / nbands, band = blockIdx.x % nbands;
  const int n_oct = c >> 3, cpr = W * n_oct;
  const int tid = threadIdx.x, o = tid % n_oct;
  const int h0 = band * ELA_BAND, h1 = min(h0 + ELA_BAND, H);
  bool has[NJ];
  int wq[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    has[j] = tid + j * ELA_NT < cpr;
    wq[j] = (tid + j * ELA_NT) / n_oct;
  }
  pdl_wait();
  float a[8], wg[NJ][8];
  {
    const float4 c0 = __ldg(reinterpret_cast<const float4 *>(chgate + (size_t)b * c + o * 8));
    const float4 c1 = __ldg(reinterpret_cast<const float4 *>(chgate + (size_t)b * c + o * 8) + 1);
    const float cg[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ch_c * cg[i];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float *g = wgate + ((size_t)b * W + (has[j] ? wq[j] : 0)) * c + o * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4 *>(g)), g1 = __ldg(reinterpret_cast<const float4 *>(g) + 1);
      wg[j][0] = g0.x; wg[j][1] = g0.y; wg[j][2] = g0.z; wg[j][3] = g0.w;
      wg[j][4] = g1.x; wg[j][5] = g1.y; wg[j][6] = g1.z; wg[j][7] = g1.w;
    }
  }
  const bf16 *img = x + (size_t)b * H * W * cs + o * 8;
  bf16 *oimg = out + (size_t)b * H * W * out_cs + o * 8;
  for (int r0 = h0; r0 < h1; r0 += 4) {
    uint4 v[4][NJ];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (r0 + r < h1 && has[j]) v[r][j] = __ldg(reinterpret_cast<const uint4 *>(img + ((size_t)(r0 + r) * W + wq[j]) * cs));
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r0 + r >= h1) break;
      const float *hgp = hgate + ((size_t)b * H + r0 + r) * c + o * 8;
      const float4 h0v = __ldg(reinterpret_cast<const float4 *>(hgp)), h1v = __ldg(reinterpret_cast<const float4 *>(hgp) + 1);
      const float hg[8] = {h0v.x, h0v.y, h0v.z, h0v.w, h1v.x, h1v.y, h1v.z, h1v.w};
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (has[j]) {
          float f[8];
          unpack8(v[r][j], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = f[i] * (a[i] + sp_c * (hg[i] * wg[j][i])) + res_c * f[i];
          *reinterpret_cast<uint4 *>(oimg + ((size_t)(r0 + r) * W + wq[j]) * out_cs) = pack8(f);
        }
    }
  }
}

int launch_ela_apply(const bf16 *x, int cs, bf16 *out, int out_cs, int c, int B, int H, int W, const float *hgate,
                     const float *wgate, const float *chgate, float ch_c, float sp_c, float res_c, cudaStream_t st) {
  B2Y_REQUIRE(c % 8 == 0 && cs % 8 == 0 && out_cs % 8 == 0, B2Y_ERR_UNSUPPORTED, "ELA apply: channels must be x8");
  if (ela_fast_path(c, W, cs, x) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int nbands = ceil_div(H, ELA_BAND);
    const int nj = ceil_div(W * (c >> 3), ELA_NT);
    auto go = [&](auto kern) -> int {
      b2y_launch(kern, dim3(B * nbands), dim3(ELA_NT), 0, st, x, cs, out, out_cs, c, H, W, nbands, hgate, wgate, chgate, ch_c, sp_c,
                 res_c);
      B2Y_LAUNCH_CHECK();
      return B2Y_OK;
    };
    return nj <= 1 ? go(ela_apply_band_kernel<1>) : nj == 2 ? go(ela_apply_band_kernel<2>) : nj == 3 ? go(ela_apply_band_kernel<3>)
                                                                                                      : go(ela_apply_band_kernel<4>);
  }
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
// block = (8 output channels, a chunk of images): the 4 x 8 weight rows (32 KB at c = 256) are staged in shared memory ONCE
// and reused for every image of the chunk; each WARP takes whole images: it loads the image's pooled vectors (4 x c floats,
// all loads in flight at once) into registers and runs the 32 dot products against shared memory, so eight images are in
// flight per block.  (v1 re-read the 1 MB weight set per image: 268 MB of L2 traffic at batch 256; v2 walked the images one
// by one with a memory latency each.)
constexpr int GATE_CH = 8;
template <int GATE_MAXM>  // ceil(c / 32): the pooled vectors and the 32 accumulators live in registers (ncu: with a fixed
                          // bound of 16 the kernel needed 255 registers and issued 6x the instructions at c = 256)
__global__ void __launch_bounds__(256) msca_gate_kernel(const float *__restrict__ gap, const float *__restrict__ se_w,
                                                        const float *__restrict__ se_b, int c, int B, int img_per_cta,
                                                        float *__restrict__ gate) {
  pdl_launch_dependents();
  extern __shared__ float wsm_gate[];  // [4][GATE_CH][c]
  const int ch0 = blockIdx.x * GATE_CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * GATE_CH * c; i += blockDim.x) {
    const int k = i % c, r = (i / c) % GATE_CH, j = i / (c * GATE_CH);
    wsm_gate[i] = (ch0 + r < c) ? __ldg(se_w + ((size_t)j * c + ch0 + r) * c + k) : 0.f;
  }
  __syncthreads();
  pdl_wait();
  const int b0 = blockIdx.y * img_per_cta, b1 = min(b0 + img_per_cta, B);
  const int nm = (c + 31) / 32;
  for (int b = b0 + warp; b < b1; b += 8) {
    float g[4][GATE_MAXM];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int m = 0; m < GATE_MAXM; ++m) {
        const int k = lane + 32 * m;
        g[j][m] = (m < nm && k < c) ? __ldg(gap + ((size_t)b * 4 + j) * c + k) : 0.f;
      }
    // 32 dot products (4 branches x 8 channels) per image: every lane accumulates its slice of all 32, then ONE butterfly
    // transposes and reduces them (31 shuffles instead of 32 x 5): lane l ends up with output l = branch (l >> 3), channel (l & 7)
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      const float *w = wsm_gate + (size_t)o * c;  // [j][r][c] with o = j * GATE_CH + r
      float s = 0.f;
#pragma unroll
      for (int m = 0; m < GATE_MAXM; ++m) {
        const int k = lane + 32 * m;
        if (m < nm && k < c) s = fmaf(w[k], g[o >> 3][m], s);
      }
      acc[o] = s;
    }
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const bool up = (lane & step) != 0;
#pragma unroll
      for (int i = 0; i < step; ++i) {
        const float send = up ? acc[i] : acc[i + step];
        const float keep = up ? acc[i + step] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
      }
    }
    const int j = lane >> 3, r = lane & 7;
    const bool live = ch0 + r < c;
    const float z = 1.f / (1.f + expf(-(acc[0] + (live ? __ldg(se_b + j * c + ch0 + r) : 0.f))));
    float mx = fmaxf(z, __shfl_xor_sync(0xffffffffu, z, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    const float e = expf(z - mx);
    float t = e + __shfl_xor_sync(0xffffffffu, e, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    if (live) gate[((size_t)b * 4 + j) * c + ch0 + r] = e / t;
  }
}

int launch_msca_gate(const float *gap, const float *se_w, const float *se_b, int c, int B, float *gate,
                     cudaStream_t st) {
  const size_t smem = (size_t)4 * GATE_CH * c * sizeof(float);
  B2Y_REQUIRE(smem <= 96 * 1024 && c <= 512, B2Y_ERR_UNSUPPORTED, "msca_gate: c = %d too large", c);
  const int tiles = ceil_div(c, GATE_CH);
  int chunks = ceil_div(148 * 2, tiles);  // ~2 CTAs per SM in total
  if (chunks > ceil_div(B, 8)) chunks = ceil_div(B, 8);
  const int per = ceil_div(B, chunks);
  auto go = [&](auto kern) -> int {
    if (smem > 48 * 1024) B2Y_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b2y_launch(kern, dim3(dim3(tiles, ceil_div(B, per))), dim3(256), smem, st, gap, se_w, se_b, c, B, per, gate);
    B2Y_LAUNCH_CHECK();
    return B2Y_OK;
  };
  const int nm = ceil_div(c, 32);
  return nm <= 1 ? go(msca_gate_kernel<1>) : nm <= 2 ? go(msca_gate_kernel<2>) : nm <= 4 ? go(msca_gate_kernel<4>)
       : nm <= 8 ? go(msca_gate_kernel<8>) : go(msca_gate_kernel<16>);
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

// Fast path (256 % (c/8) == 0): CTA = (image, a run of pixels), thread = a fixed channel octet whose 4 x 8 gates stay in
// registers while it walks over pixels (the per-pixel kernel above re-reads 128 B of gates for every 64 B of branch data).
__global__ void __launch_bounds__(256) wsum4_band_kernel(Ptr4 br, const float *__restrict__ gate, bf16 *__restrict__ out,
                                                         int out_cs, int c, int HW, int px_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  const int n_oct = c >> 3;
  const int b = blockIdx.y, o = threadIdx.x % n_oct, pstep = blockDim.x / n_oct;
  float g[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 *gp = reinterpret_cast<const float4 *>(gate + ((size_t)b * 4 + j) * c + o * 8);
    const float4 a = __ldg(gp), d = __ldg(gp + 1);
    g[j][0] = a.x; g[j][1] = a.y; g[j][2] = a.z; g[j][3] = a.w; g[j][4] = d.x; g[j][5] = d.y; g[j][6] = d.z; g[j][7] = d.w;
  }
  const int p0 = blockIdx.x * px_per_cta, p1 = min(p0 + px_per_cta, HW);
  for (int px = p0 + threadIdx.x / n_oct; px < p1; px += 2 * pstep) {
    uint4 v[2][4];
    const bool two = px + pstep < p1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[0][j] = __ldg(reinterpret_cast<const uint4 *>(br.p[j] + ((size_t)b * HW + px) * br.cs[j] + o * 8));
      if (two) v[1][j] = __ldg(reinterpret_cast<const uint4 *>(br.p[j] + ((size_t)b * HW + px + pstep) * br.cs[j] + o * 8));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8(v[u][j], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(g[j][k], f[k], acc[k]);
      }
      *reinterpret_cast<uint4 *>(out + ((size_t)b * HW + px + u * pstep) * out_cs + o * 8) = pack8(acc);
    }
  }
}

int launch_wsum4(const bf16 *const *br, const int *br_cs, const float *gate, bf16 *out, int out_cs, int c, int B,
                 int HW, cudaStream_t st) {
  B2Y_REQUIRE(c % 8 == 0 && out_cs % 8 == 0, B2Y_ERR_UNSUPPORTED, "wsum4: channels must be x8");
  if (256 % (c >> 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(gate) & 15) == 0) {
    Ptr4 q;
    bool ok = true;
    for (int j = 0; j < 4; ++j) {
      q.p[j] = br[j];
      q.cs[j] = br_cs[j];
      ok = ok && br_cs[j] % 8 == 0 && (reinterpret_cast<uintptr_t>(br[j]) & 15) == 0;
    }
    if (ok) {
      const int per_iter = 2 * (256 / (c >> 3));
      int px_per_cta = per_iter * 4;  // four iterations per thread
      if (px_per_cta > HW) px_per_cta = HW;
      b2y_launch(wsum4_band_kernel, dim3(dim3(ceil_div(HW, px_per_cta), B)), dim3(256), 0, st, q, gate, out, out_cs, c, HW, px_per_cta);
      B2Y_LAUNCH_CHECK();
      return B2Y_OK;
    }
  }
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

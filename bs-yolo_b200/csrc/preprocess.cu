// Fused letterbox + BGR->RGB + HWC->(NHWC|NCHW) + /255 preprocess.
// Reference: LetterBox.__call__ (ultralytics/data/augment.py:1535-1601: resize INTER_LINEAR :1586, border 114 :1589-1591)
// and BasePredictor.preprocess (ultralytics/engine/predictor.py:116-134: [..., ::-1], transpose, /255).
// The resize reproduces OpenCV's uint8 INTER_LINEAR bit-for-bit for down-scaling (11-bit fixed-point coefficients,
// integer two-pass blend; SURVEY §8 "a1 fine print"); identity when the source already has the unpadded size.
#include "common.cuh"

namespace {
struct PreParams {
  const uint8_t *src;
  void *dst;
  int B, h, w, H, W, top, left, uh, uw, fmt, resize;
  double scale_x, scale_y;
};

__device__ __forceinline__ void axis_coeff(int d, double scale, int src, int &s, int &a0, int &a1) {
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= src - 1) { s = src - 1; f = 0.f; }
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  a0 = 2048 - a1;
}

__global__ void __launch_bounds__(256) preprocess_kernel(PreParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)p.B * p.H * p.W;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % p.W);
  const int y = (int)((idx / p.W) % p.H);
  const int b = (int)(idx / ((long long)p.W * p.H));
  int bgr[3] = {114, 114, 114};
  const int ux = x - p.left, uy = y - p.top;
  if (ux >= 0 && ux < p.uw && uy >= 0 && uy < p.uh) {
    const uint8_t *img = p.src + (size_t)b * p.h * p.w * 3;
    if (!p.resize) {
      const uint8_t *s = img + ((size_t)uy * p.w + ux) * 3;
      bgr[0] = s[0]; bgr[1] = s[1]; bgr[2] = s[2];
    } else {
      int sx, xa0, xa1, sy, ya0, ya1;
      axis_coeff(ux, p.scale_x, p.w, sx, xa0, xa1);
      axis_coeff(uy, p.scale_y, p.h, sy, ya0, ya1);
      const int sx1 = min(sx + 1, p.w - 1), sy1 = min(sy + 1, p.h - 1);
      const uint8_t *r0 = img + (size_t)sy * p.w * 3, *r1 = img + (size_t)sy1 * p.w * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = r0[sx * 3 + c] * xa0 + r0[sx1 * 3 + c] * xa1;
        const int h1 = r1[sx * 3 + c] * xa0 + r1[sx1 * 3 + c] * xa1;
        bgr[c] = (((ya0 * (h0 >> 4)) >> 16) + ((ya1 * (h1 >> 4)) >> 16) + 2) >> 2;
      }
    }
  }
  if (p.fmt == 0) {  // space-to-depth: pixel (y,x) -> super-pixel (y/2, x/2), channel group (y&1)*2 + (x&1)
    float f[4] = {__fdiv_rn((float)bgr[2], 255.f), __fdiv_rn((float)bgr[1], 255.f), __fdiv_rn((float)bgr[0], 255.f), 0.f};
    const size_t sp = ((size_t)b * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
    reinterpret_cast<uint2 *>(p.dst)[sp * 4 + ((y & 1) * 2 + (x & 1))] = pack4(f);
  } else if (p.fmt == 1) {
    float *o = reinterpret_cast<float *>(p.dst) + (size_t)b * 3 * p.H * p.W + (size_t)y * p.W + x;
    const size_t plane = (size_t)p.H * p.W;
    o[0] = __fdiv_rn((float)bgr[2], 255.f);
    o[plane] = __fdiv_rn((float)bgr[1], 255.f);
    o[2 * plane] = __fdiv_rn((float)bgr[0], 255.f);
  } else {
    uint8_t *o = reinterpret_cast<uint8_t *>(p.dst) + (size_t)idx * 3;
    o[0] = (uint8_t)bgr[0]; o[1] = (uint8_t)bgr[1]; o[2] = (uint8_t)bgr[2];
  }
}
// Network-input fast path (no resize): one space-to-depth super-pixel = 2 x 2 source pixels per thread and round, kUnroll
// rounds in flight (persistent grid, grid-stride over blocks of 256 * kUnroll consecutive super-pixels).
// Reads 2 x 6 source bytes (three 16-bit loads per row: x is even, so the BGR pair is 2-byte aligned whenever the row
// pitch and the placement are even) and stores the 16 channels ((y&1)*2 + (x&1))*4 + {R,G,B,0} as two 16-byte vectors:
// consecutive lanes write consecutive 32-byte super-pixels.  v/255 in bf16: float(v) * (1/255.f) rounded to bf16 gives the
// same bits as the generic path's correctly rounded divide for all 256 byte values (checked exhaustively in the tests).
constexpr int kPreUnroll = 4;

__device__ __forceinline__ uint32_t norm2_bf16(uint32_t lo, uint32_t hi) {  // {bf16(lo/255), bf16(hi/255)} packed
  const float k = 1.f / 255.f;
  const float a = __fmul_rn(__fsub_rn(__uint_as_float(0x4B000000u | lo), 8388608.f), k);
  const float b = __fmul_rn(__fsub_rn(__uint_as_float(0x4B000000u | hi), 8388608.f), k);
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

__global__ void __launch_bounds__(256) preprocess_s2d_kernel(PreParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int W2 = p.W >> 1, H2 = p.H >> 1;
  const long long total = (long long)p.B * H2 * W2;
  const bool fast_ok = ((p.w & 1) == 0) && ((p.left & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.src) & 1) == 0);
  for (long long base = (long long)blockIdx.x * (256 * kPreUnroll); base < total; base += (long long)gridDim.x * (256 * kPreUnroll)) {
    uint32_t raw[kPreUnroll][2][3];  // per round and row: B0G0 | R0B1 | G1R1 (16-bit words)
#pragma unroll
    for (int u = 0; u < kPreUnroll; ++u) {
      const long long idx = base + u * 256 + threadIdx.x;
      const int xs = (int)(idx % W2);
      const int ys = (int)((idx / W2) % H2);
      const int b = (int)(idx / ((long long)W2 * H2));
      const uint8_t *img = p.src + (size_t)b * p.h * p.w * 3;
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const int uy = 2 * ys + row - p.top, ux = 2 * xs - p.left;
        uint32_t px[2][3] = {{114, 114, 114}, {114, 114, 114}};  // B, G, R of the two pixels
        bool packed = false;
        if (idx < total && uy >= 0 && uy < p.uh) {
          const uint8_t *s = img + ((size_t)uy * p.w + ux) * 3;
          if (fast_ok && ux >= 0 && ux + 1 < p.uw) {
            const uint16_t *s16 = reinterpret_cast<const uint16_t *>(s);
            raw[u][row][0] = __ldg(s16);
            raw[u][row][1] = __ldg(s16 + 1);
            raw[u][row][2] = __ldg(s16 + 2);
            packed = true;
          } else {
#pragma unroll
            for (int q = 0; q < 2; ++q)
              if (ux + q >= 0 && ux + q < p.uw) {
                px[q][0] = s[q * 3]; px[q][1] = s[q * 3 + 1]; px[q][2] = s[q * 3 + 2];
              }
          }
        }
        if (!packed) {
          raw[u][row][0] = px[0][0] | (px[0][1] << 8);
          raw[u][row][1] = px[0][2] | (px[1][0] << 8);
          raw[u][row][2] = px[1][1] | (px[1][2] << 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kPreUnroll; ++u) {
      const long long idx = base + u * 256 + threadIdx.x;
      if (idx >= total) continue;
      uint32_t o[8];
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const uint32_t a = raw[u][row][0], c = raw[u][row][1], d = raw[u][row][2];
        const uint32_t b0 = a & 255u, g0 = a >> 8, r0 = c & 255u, b1 = c >> 8, g1 = d & 255u, r1 = d >> 8;
        o[row * 4 + 0] = norm2_bf16(r0, g0);          // R, G
        o[row * 4 + 1] = norm2_bf16(b0, 0u) & 0xFFFFu;  // B, 0
        o[row * 4 + 2] = norm2_bf16(r1, g1);
        o[row * 4 + 3] = norm2_bf16(b1, 0u) & 0xFFFFu;
      }
      // one 256-bit store per super-pixel: every lane writes a whole 32-byte sector
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(reinterpret_cast<uint4 *>(p.dst) + (size_t)idx * 2),
                   "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                   : "memory");
    }
  }
}
}  // namespace

extern "C" int b2y_preprocess(const void *src, int32_t B, int32_t h, int32_t w, void *dst, int32_t H, int32_t W,
                              int32_t top, int32_t left, int32_t uh, int32_t uw, int32_t dst_format, void *stream) {
  B2Y_REQUIRE(src && dst && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, B2Y_ERR_INVALID, "preprocess: bad arguments");
  B2Y_REQUIRE(top >= 0 && left >= 0 && top + uh <= H && left + uw <= W && uh > 0 && uw > 0, B2Y_ERR_INVALID,
              "preprocess: letterbox geometry (%d,%d,%dx%d) does not fit %dx%d", top, left, uh, uw, H, W);
  B2Y_REQUIRE(dst_format >= 0 && dst_format <= 2, B2Y_ERR_INVALID, "preprocess: unknown dst_format %d", dst_format);
  B2Y_REQUIRE(dst_format != 0 || (H % 2 == 0 && W % 2 == 0), B2Y_ERR_INVALID, "preprocess: network input needs even H, W");
  B2YDeviceGuard guard(b2y_device_of(dst));
  PreParams p;
  p.src = reinterpret_cast<const uint8_t *>(src);
  p.dst = dst;
  p.B = B; p.h = h; p.w = w; p.H = H; p.W = W; p.top = top; p.left = left; p.uh = uh; p.uw = uw; p.fmt = dst_format;
  p.resize = !(uh == h && uw == w);
  p.scale_x = 1.0 / ((double)uw / (double)w);
  p.scale_y = 1.0 / ((double)uh / (double)h);
  if (dst_format == 0 && !p.resize && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    const long long nsp = (long long)B * (H >> 1) * (W >> 1);
    const long long blocks = (nsp + 256 * kPreUnroll - 1) / (256 * kPreUnroll);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long cap = (long long)sms * 8;  // persistent: 8 CTAs of 256 threads per SM
    b2y_launch(preprocess_s2d_kernel, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(256), 0, (cudaStream_t)stream, p);
    B2Y_LAUNCH_CHECK();
    return B2Y_OK;
  }
  const long long total = (long long)B * H * W;
  b2y_launch(preprocess_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// Segment head: the two device pieces the Detect kernels do not already cover.
//
//  d2s_kernel ............ depth-to-space shuffle behind Proto.upsample (nn.ConvTranspose2d(c, c, 2, 2, 0),
//                          ultralytics/nn/modules/block.py:91).  Kernel = stride = 2, so each output pixel is one tap of one
//                          input pixel: the planner runs the four taps as ONE tensor-core 1x1 conv with 4c output channels
//                          ordered (dy, dx, co) and this kernel scatters them to (2y+dy, 2x+dx).
//  process_mask_kernel ... ops.process_mask(protos, masks_in, bboxes, shape, upsample=True)
//                          (ultralytics/utils/ops.py:663-693) for every kept row of a batch in one launch:
//                            masks = masks_in @ protos.view(c, -1)                      (:680)
//                            crop_mask against the box scaled to the prototype grid     (:681-690, :644-660)
//                            F.interpolate(bilinear, align_corners=False) to (H, W)      (:692)
//                            masks.gt_(0.0)                                              (:693)
//                          A CTA owns one 64 x 64 output tile of one kept row: the (<= 18 x 18) low-resolution mask values
//                          it needs are computed once into shared memory (nm-long dot products against bf16 prototypes,
//                          fp32 accumulate) and each thread then blends / thresholds 16 output pixels.  The low-resolution
//                          mask, its cropped copy and the upsampled float mask of the reference never touch HBM.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) d2s_kernel(const bf16 *__restrict__ in, int in_cs, bf16 *__restrict__ out, int out_cs,
                                                  int c, long long total, int Hi, int Wi) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk of the INPUT view
  if (idx >= total) return;
  const int cg = (4 * c) >> 3;
  const int c0 = (int)(idx % cg) * 8;  // channel within the 4c-channel input pixel
  const long long pix = idx / cg;
  const int wi = (int)(pix % Wi);
  const int hi = (int)((pix / Wi) % Hi);
  const long long b = pix / ((long long)Wi * Hi);
  const int tap = c0 / c, co = c0 - tap * c;  // c % 8 == 0: a chunk never straddles two taps
  const int dy = tap >> 1, dx = tap & 1;
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + (size_t)pix * in_cs + c0));
  bf16 *d = out + ((size_t)(b * 2 * Hi + 2 * hi + dy) * (2 * Wi) + 2 * wi + dx) * out_cs + co;
  *reinterpret_cast<uint4 *>(d) = v;
}

constexpr int TILE = 64;       // output tile edge
constexpr int LOW = TILE / 4;  // prototype-grid pixels under one tile edge
constexpr int LR = LOW + 2;    // + one halo row / column on each side for the bilinear taps
constexpr int MAX_NM = 64;

struct MaskParams {
  const bf16 *protos;  // (B, mh, mw, cs) NHWC, nm channels used
  int cs, nm, mh, mw, H, W;
  const float *coeffs;  // (B, max_det, nm)
  const float *dets;    // (B, max_det, 6) boxes in network-input pixels (before scale_boxes)
  const int *row_img, *row_det;  // per output row: image index, detection index within the image
  int rows, max_det;
  float *out;  // (rows, H, W) fp32 0 / 1
};

__global__ void __launch_bounds__(256) process_mask_kernel(MaskParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float coef[MAX_NM];
  __shared__ float low[LR][LR + 1];
  const int r = blockIdx.z;
  const int b = p.row_img[r], k = p.row_det[r];
  const int oy0 = blockIdx.y * TILE, ox0 = blockIdx.x * TILE;
  const int ly0 = oy0 / 4 - 1, lx0 = ox0 / 4 - 1;  // prototype-grid origin of the shared tile (may be -1)
  if (threadIdx.x < p.nm) coef[threadIdx.x] = p.coeffs[((size_t)b * p.max_det + k) * p.nm + threadIdx.x];
  // box on the prototype grid (ops.py:681-688: x *= mw / iw, y *= mh / ih; fp32 products like the reference's in-place muls)
  const float *box = p.dets + ((size_t)b * p.max_det + k) * 6;
  const float wr = (float)p.mw / (float)p.W, hr = (float)p.mh / (float)p.H;
  const float x1 = __fmul_rn(box[0], wr), y1 = __fmul_rn(box[1], hr), x2 = __fmul_rn(box[2], wr), y2 = __fmul_rn(box[3], hr);
  __syncthreads();
  for (int i = threadIdx.x; i < LR * LR; i += blockDim.x) {
    const int ly = ly0 + i / LR, lx = lx0 + i % LR;
    float v = 0.f;
    if (ly >= 0 && ly < p.mh && lx >= 0 && lx < p.mw) {
      const float fx = (float)lx, fy = (float)ly;
      if (fx >= x1 && fx < x2 && fy >= y1 && fy < y2) {  // crop_mask (ops.py:657-660): zero outside the box
        const bf16 *q = p.protos + ((size_t)(b * p.mh + ly) * p.mw + lx) * p.cs;
        for (int c = 0; c < p.nm; c += 8) {
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4 *>(q + c)), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) v = fmaf(coef[c + j], f[j], v);
        }
      }
    }
    low[i / LR][i % LR] = v;
  }
  __syncthreads();
  // bilinear, align_corners=False, scale 1/4 (ATen upsample_bilinear2d: src = 0.25 * (dst + 0.5) - 0.5 clamped at 0,
  // i1 = i0 + (i0 < in - 1), lambda1 = src - i0)
  const int tx = threadIdx.x & 63, ty4 = threadIdx.x >> 6;  // 64 columns x 4 row groups of 16 rows
  const int ox = ox0 + tx;
  if (ox >= p.W) return;
  float sx = 0.25f * ((float)ox + 0.5f) - 0.5f;
  if (sx < 0.f) sx = 0.f;
  const int ix0 = (int)sx, ix1 = ix0 + (ix0 < p.mw - 1 ? 1 : 0);
  const float lx1 = sx - (float)ix0, lx0w = 1.f - lx1;
  const int cx0 = ix0 - lx0, cx1 = ix1 - lx0;
  float *o = p.out + (size_t)r * p.H * p.W + ox;
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int oy = oy0 + ty4 * 16 + j;
    if (oy >= p.H) break;
    float sy = 0.25f * ((float)oy + 0.5f) - 0.5f;
    if (sy < 0.f) sy = 0.f;
    const int iy0 = (int)sy, iy1 = iy0 + (iy0 < p.mh - 1 ? 1 : 0);
    const float ly1 = sy - (float)iy0, ly0w = 1.f - ly1;
    const int cy0 = iy0 - ly0, cy1 = iy1 - ly0;
    const float v = ly0w * (lx0w * low[cy0][cx0] + lx1 * low[cy0][cx1]) + ly1 * (lx0w * low[cy1][cx0] + lx1 * low[cy1][cx1]);
    o[(size_t)oy * p.W] = v > 0.f ? 1.f : 0.f;
  }
}
}  // namespace

int launch_d2s(const bf16 *in, int in_cs, bf16 *out, int out_cs, int c, int B, int Hi, int Wi, cudaStream_t st) {
  B2Y_REQUIRE(c > 0 && c % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0,
              B2Y_ERR_UNSUPPORTED, "depth-to-space needs 16-byte aligned views and c %% 8 == 0 (c = %d)", c);
  const long long total = (long long)B * Hi * Wi * (4 * c / 8);
  b2y_launch(d2s_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, in, in_cs, out, out_cs, c, total, Hi, Wi);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

extern "C" int b2y_process_mask(const void *protos, int32_t proto_cs, int32_t nm, int32_t mh, int32_t mw, const void *coeffs,
                                const void *dets, int32_t max_det, const void *row_img, const void *row_det, int32_t rows,
                                int32_t H, int32_t W, void *masks, void *stream) {
  B2Y_REQUIRE(protos && coeffs && dets && masks && row_img && row_det, B2Y_ERR_INVALID, "process_mask: null argument");
  B2Y_REQUIRE(rows > 0 && max_det > 0 && H > 0 && W > 0, B2Y_ERR_INVALID, "process_mask: bad sizes");
  B2Y_REQUIRE(nm > 0 && nm <= MAX_NM && nm % 8 == 0 && proto_cs >= nm && proto_cs % 8 == 0 && ((uintptr_t)protos & 15) == 0,
              B2Y_ERR_UNSUPPORTED, "process_mask: nm = %d (pitch %d) must be a multiple of 8, at most %d, 16-byte aligned", nm,
              proto_cs, MAX_NM);
  B2Y_REQUIRE(mh * 4 == H && mw * 4 == W, B2Y_ERR_UNSUPPORTED,
              "process_mask: prototypes %dx%d must be the network input %dx%d divided by 4 (Proto on the stride-8 map)", mh, mw,
              H, W);
  B2Y_REQUIRE(rows <= 65535, B2Y_ERR_UNSUPPORTED, "process_mask: more than 65535 rows in one call");
  B2YDeviceGuard guard(b2y_device_of(masks));
  MaskParams p;
  p.protos = reinterpret_cast<const bf16 *>(protos);
  p.cs = proto_cs; p.nm = nm; p.mh = mh; p.mw = mw; p.H = H; p.W = W;
  p.coeffs = reinterpret_cast<const float *>(coeffs);
  p.dets = reinterpret_cast<const float *>(dets);
  p.row_img = reinterpret_cast<const int *>(row_img);
  p.row_det = reinterpret_cast<const int *>(row_det);
  p.rows = rows; p.max_det = max_det;
  p.out = reinterpret_cast<float *>(masks);
  b2y_launch(process_mask_kernel, dim3(ceil_div(W, TILE), ceil_div(H, TILE), rows), dim3(256), 0, (cudaStream_t)stream, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

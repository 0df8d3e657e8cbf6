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
//                          A CTA owns one 64 x 64 output tile of one kept row: the 18 x 18 low-resolution mask values it
//                          needs are computed once into shared memory (nm-long dot products against bf16 prototypes,
//                          fp32 accumulate), each thread then blends / thresholds 4 x 4 output pixels and writes them
//                          with 16-byte stores; tiles that miss the box are zero-filled without touching the
//                          prototypes.  The low-resolution mask, its cropped copy and the upsampled float mask of the
//                          reference never touch HBM: the kernel is a pure HBM writer (rows x H x W x 4 bytes).
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
  // box on the prototype grid (ops.py:681-688: x *= mw / iw, y *= mh / ih; fp32 products like the reference's in-place muls)
  const float *box = p.dets + ((size_t)b * p.max_det + k) * 6;
  const float wr = (float)p.mw / (float)p.W, hr = (float)p.mh / (float)p.H;
  const float x1 = __fmul_rn(box[0], wr), y1 = __fmul_rn(box[1], hr), x2 = __fmul_rn(box[2], wr), y2 = __fmul_rn(box[3], hr);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 quads of 4 columns x 16 rows per pass, 4 passes
  const int ox = ox0 + 4 * tx;
  float *o = p.out + (size_t)r * p.H * p.W + ox;
  // the tile's prototype cells (clamped to the grid) against the box: no overlap -> the tile is all zeros (crop_mask)
  const float cx_lo = (float)max(lx0, 0), cx_hi = (float)min(lx0 + LR - 1, p.mw - 1);
  const float cy_lo = (float)max(ly0, 0), cy_hi = (float)min(ly0 + LR - 1, p.mh - 1);
  if (!(cx_hi >= x1 && cx_lo < x2 && cy_hi >= y1 && cy_lo < y2)) {
    if (ox < p.W) {
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int oy = oy0 + pass * 16 + ty;
        if (oy < p.H) *reinterpret_cast<float4 *>(o + (size_t)oy * p.W) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }
  if (threadIdx.x < p.nm) coef[threadIdx.x] = p.coeffs[((size_t)b * p.max_det + k) * p.nm + threadIdx.x];
  __syncthreads();
  // Low-resolution cropped mask values under the tile, one halo cell on each side.  Coordinates are CLAMPED to the grid
  // (edge replication): ATen's bilinear rule clamps the source index at 0 and takes i1 = i0 on the last cell, which is the
  // same as blending a cell with its own replica.
  for (int i = threadIdx.x; i < LR * LR; i += blockDim.x) {
    const int ly = min(max(ly0 + i / LR, 0), p.mh - 1), lx = min(max(lx0 + i % LR, 0), p.mw - 1);
    const float fx = (float)lx, fy = (float)ly;
    float v = 0.f;
    if (fx >= x1 && fx < x2 && fy >= y1 && fy < y2) {  // crop_mask (ops.py:657-660): zero outside the box
      const bf16 *q = p.protos + ((size_t)(b * p.mh + ly) * p.mw + lx) * p.cs;
      for (int c = 0; c < p.nm; c += 8) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4 *>(q + c)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v = fmaf(coef[c + j], f[j], v);
      }
    }
    low[i / LR][i % LR] = v;
  }
  __syncthreads();
  if (ox >= p.W) return;
  // Bilinear x4, align_corners=False (ATen upsample_bilinear2d: src = 0.25 * (dst + 0.5) - 0.5): output 4m + i lies between
  // cells m - 1 and m (i = 0, 1; weight of the upper cell 0.625, 0.875) or m and m + 1 (i = 2, 3; 0.125, 0.375).  A thread
  // owns four consecutive columns (one cell triple) and writes them with one 16-byte store per row.
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int row = pass * 16 + ty, oy = oy0 + row;
    if (oy >= p.H) break;
    const int m = row >> 2, i = row & 3;
    const int cy0 = m + (i >> 1);
    const float wy1 = (i == 0) ? 0.625f : (i == 1) ? 0.875f : (i == 2) ? 0.125f : 0.375f, wy0 = 1.f - wy1;
    const float c0 = wy0 * low[cy0][tx] + wy1 * low[cy0 + 1][tx];
    const float c1 = wy0 * low[cy0][tx + 1] + wy1 * low[cy0 + 1][tx + 1];
    const float c2 = wy0 * low[cy0][tx + 2] + wy1 * low[cy0 + 1][tx + 2];
    float4 v;
    v.x = (0.375f * c0 + 0.625f * c1) > 0.f ? 1.f : 0.f;
    v.y = (0.125f * c0 + 0.875f * c1) > 0.f ? 1.f : 0.f;
    v.z = (0.875f * c1 + 0.125f * c2) > 0.f ? 1.f : 0.f;
    v.w = (0.625f * c1 + 0.375f * c2) > 0.f ? 1.f : 0.f;
    *reinterpret_cast<float4 *>(o + (size_t)oy * p.W) = v;
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
  B2Y_REQUIRE(W % 4 == 0 && ((uintptr_t)masks & 15) == 0, B2Y_ERR_UNSUPPORTED, "process_mask: W %% 4 and a 16-byte aligned output");
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

// C3k2 (c3k = False, one Bottleneck) as ONE kernel:
//   [y0 | y1] = SiLU(cv1 x)                      1x1, cin -> 2c            (block.py:4148-4156 -> C2f.forward :3306-3310)
//   h         = SiLU(conv3x3(y1; W1) + b1)       c -> chid                 (Bottleneck, block.py:3405-3419)
//   y2        = y1 + SiLU(conv3x3(h; W2) + b2)   chid -> c
//   out       = SiLU(cv2 [y0 | y1 | y2])         1x1, 3c -> cout
// EXPERIMENTAL (opt-in through the planner: B2Y_FUSE_C3K2=1; the default plan never emits B2Y_OP_C3K2).  First version, written
// at the end of round 1: parity-green on the GPU (tests/test_gpu_parity.py::test_fused_c3k2_vs_emulator with the flag set) but
// not yet faster than the three-launch lowering -- layer 2: 196 us vs 58 + 75 + 87, layer 4 (one CTA per SM: 208 KB of shared
// memory, 512 TMEM columns): 145 us vs 124; the CUDA-graphed network 1.91 ms vs 1.86 ms.  The four phases of a tile are
// serial and drain 21 16-column chunks through x16 loads; see profiles/r01_notes.md for what to change first.
//
// The three launches of such a block (cv1, fused Bottleneck, cv2) move 682 MB for layer 2 of yolo11n at batch 64; this kernel
// reads x once (with a 2-pixel halo) and writes out once (315 MB).  Per tile of H2 x TW output pixels (P = TW + 4):
//   * TMA loads the input halo tile X0: (H2+4) x P pixels, cin channels (OOB zero fill; single buffered: only GEMM 0 reads it).
//   * GEMM 0 (1x1): D0[q] = X0[q] . W0 for every halo pixel q -> epilogue 0: +b0, SiLU, ZERO outside the image (the
//     Bottleneck's conv sees zero padding of y1, not cv1 of zero-padded x) -> bf16 rows of the shared-memory tile Y
//     (2c channels per pixel, K-major swizzled: the A operand of GEMM 1 and of GEMM 3).
//   * GEMM 1 (3x3 over the y1 half of Y, row-shifted descriptors) -> epilogue 1: +b1, SiLU, zero outside -> tile A2.
//   * GEMM 2 (3x3 over A2) -> epilogue 2: +b2, SiLU, + y1 (read back from Y) -> bf16 tile Y2 (row m = output pixel m).
//   * GEMM 3 (1x1): D3[m] = Y[m + 2P + 2] . W3[:, 0:2c] + Y2[m] . W3[:, 2c:3c] -> epilogue 3: +b3, SiLU -> staging ->
//     TMA store.  D3 reuses the TMEM columns of D0; the staging tile reuses the shared memory of Y.
// The phases of a tile are strictly serial inside a CTA (one mbarrier per hand-over, one use per tile); two co-resident
// CTAs per SM overlap each other.  Warp roles as in conv_pair.cu: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator,
// warps 2..9 epilogue.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int kEpiWarps = 8, kEpiThreads = 32 * kEpiWarps, kNumThreads = 64 + kEpiThreads;

struct C3k2Geom {
  int B, H, W, cin, c, chid, cout;
  int TH, TW, P, MB, H2, NB0, NB1, tiles_h, tiles_w;
  long long ntiles;
  int row0, rowY, row2, rowY2;          // bytes per pixel row of X0 / Y / A2 / Y2
  int sw0_mask, swY_mask, sw2_mask, swY2_mask, o_sw_mask;
  int x_tx_bytes, w_tx_bytes;
  int w1_tile, w2_tile;                 // bytes per tap tile of W1 / W2 (1024-byte aligned)
  int x_off, y_off, a2_off, y2_off, w0_off, w1_off, w2_off, w3a_off, w3b_off, bias_off;  // dynamic smem byte offsets
  int o_box_bytes, n_oboxes;            // staging (aliases Y): boxes of 64 channels, MB*128 rows each
  int tmem_cols, d1_col, d2_col;        // D0 and D3 at column 0, D1 at d1_col, D2 at d2_col
  unsigned dhi0, dhiY, dhiW1, dhi2, dhiY2;  // descriptor high words by row width: X0/W0, Y/W3a, W1, A2/W2, Y2/W3b
  unsigned idesc0, idesc1, idesc2, idesc3;
  const float *bias;                    // fp32 [2c | chid | c | cout]
};

__device__ __forceinline__ void store_chunk(uint32_t base, uint32_t off, uint32_t sw_mask, const float *v) {
  const uint32_t xr = ((off >> 7) & sw_mask) << 4;
  const uint4 p0 = pack8(v), p1 = pack8(v + 8);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (off ^ xr)), "r"(p0.x), "r"(p0.y), "r"(p0.z), "r"(p0.w) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((off + 16u) ^ xr)), "r"(p1.x), "r"(p1.y), "r"(p1.z), "r"(p1.w) : "memory");
}

__device__ __forceinline__ void silu16(float (&v)[16], const float *bias_half, bool keep) {
  const float4 *bp = reinterpret_cast<const float4 *>(bias_half);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 hb = bp[k];
    v[4 * k] = keep ? silu_half(fmaf(v[4 * k], 0.5f, hb.x)) : 0.f;
    v[4 * k + 1] = keep ? silu_half(fmaf(v[4 * k + 1], 0.5f, hb.y)) : 0.f;
    v[4 * k + 2] = keep ? silu_half(fmaf(v[4 * k + 2], 0.5f, hb.z)) : 0.f;
    v[4 * k + 3] = keep ? silu_half(fmaf(v[4 * k + 3], 0.5f, hb.w)) : 0.f;
  }
}

template <int KS0, int KC>  // k-steps (16 channels each) of cv1's input / of one half c of cv1's output
__global__ void __launch_bounds__(kNumThreads, 2)
c3k2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW0,
            const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
            const __grid_constant__ CUtensorMap tmW3a, const __grid_constant__ CUtensorMap tmW3b,
            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ C3k2Geom g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t fullX = smem_u32(&bars[0]), emptyX = smem_u32(&bars[1]), wfull = smem_u32(&bars[2]);
  const uint32_t d0_full = smem_u32(&bars[3]), y_full = smem_u32(&bars[4]);
  const uint32_t d1_full = smem_u32(&bars[5]), a2_full = smem_u32(&bars[6]);
  const uint32_t d2_full = smem_u32(&bars[7]), y2_full = smem_u32(&bars[8]);
  const uint32_t d3_full = smem_u32(&bars[9]), d0_empty = smem_u32(&bars[10]);
  float *bias_s = reinterpret_cast<float *>(smem_gen + g.bias_off);  // [2c | chid | c | cout], pre-halved (SiLU form)
  const int nbias = 2 * g.c + g.chid + g.c + g.cout;

  if (threadIdx.x == 0) {
    mbar_init(fullX, 1);
    mbar_init(emptyX, 1);
    mbar_init(wfull, 1);
    mbar_init(d0_full, 1);
    mbar_init(y_full, kEpiThreads);
    mbar_init(d1_full, 1);
    mbar_init(a2_full, kEpiThreads);
    mbar_init(d2_full, 1);
    mbar_init(y2_full, kEpiThreads);
    mbar_init(d3_full, 1);
    mbar_init(d0_empty, kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < nbias; i += kNumThreads) bias_s[i] = 0.5f * g.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_launch_dependents();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t ntiles = (uint32_t)g.ntiles;
  const uint32_t x_base = smem_base + (uint32_t)g.x_off, y_base = smem_base + (uint32_t)g.y_off;
  const uint32_t a2_base = smem_base + (uint32_t)g.a2_off, y2_base = smem_base + (uint32_t)g.y2_off;
  const uint32_t w0_base = smem_base + (uint32_t)g.w0_off, w1_base = smem_base + (uint32_t)g.w1_off;
  const uint32_t w2_base = smem_base + (uint32_t)g.w2_off;
  const uint32_t w3a_base = smem_base + (uint32_t)g.w3a_off, w3b_base = smem_base + (uint32_t)g.w3b_off;
  const int tiles_per_img = g.tiles_h * g.tiles_w;
  const int c = g.c;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
      mbar_expect_tx(wfull, (uint32_t)g.w_tx_bytes);
      tma_load_2d(w0_base, &tmW0, wfull, 0, 0);
      for (int t = 0; t < 9; ++t) {
        tma_load_2d(w1_base + t * g.w1_tile, &tmW1, wfull, t * c, 0);
        tma_load_2d(w2_base + t * g.w2_tile, &tmW2, wfull, t * g.chid, 0);
      }
      tma_load_2d(w3a_base, &tmW3a, wfull, 0, 0);
      tma_load_2d(w3b_base, &tmW3b, wfull, 2 * c, 0);
    }
    __syncwarp();
    pdl_wait();
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = (int)(tile / tiles_per_img), r = (int)(tile % tiles_per_img);
      const int h0 = (r / g.tiles_w) * g.H2, w0 = (r % g.tiles_w) * g.TW;
      mbar_wait(emptyX, (it & 1) ^ 1);  // GEMM 0 of the previous tile has consumed X0
      if (elect_one()) {
        mbar_expect_tx(fullX, (uint32_t)g.x_tx_bytes);
        tma_load_4d(x_base, &tmX, fullX, 0, w0 - 2, h0 - 2, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    const uint64_t one16 = (uint64_t)(1u << 16);
    const uint64_t desc0 = ((uint64_t)g.dhi0 << 32) | one16, descY = ((uint64_t)g.dhiY << 32) | one16;
    const uint64_t descW1 = ((uint64_t)g.dhiW1 << 32) | one16, desc2 = ((uint64_t)g.dhi2 << 32) | one16;
    const uint64_t descY2 = ((uint64_t)g.dhiY2 << 32) | one16;
    const uint32_t P = (uint32_t)g.P;
    const uint32_t x16 = (x_base & 0x3FFFFu) >> 4, y16 = (y_base & 0x3FFFFu) >> 4;
    const uint32_t a16 = (a2_base & 0x3FFFFu) >> 4, y216 = (y2_base & 0x3FFFFu) >> 4;
    mbar_wait(wfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      // ---- GEMM 0: cv1 over every pixel of the halo tile
      mbar_wait(fullX, ph);
      mbar_wait(d0_empty, ph ^ 1);  // epilogue 3 of the previous tile has drained D3 (same TMEM columns)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint64_t db = desc0 | (uint64_t)((w0_base & 0x3FFFFu) >> 4);
        for (int j = 0; j < g.NB0; ++j) {
          const uint32_t d = tmem_base + (uint32_t)(j * 2 * c);
          const uint64_t da = desc0 | (uint64_t)(x16 + (((uint32_t)j * 128u * (uint32_t)g.row0) >> 4));
#pragma unroll
          for (int k = 0; k < KS0; ++k) umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), g.idesc0, k ? 1u : 0u);
        }
        umma_commit(d0_full);
        umma_commit(emptyX);  // X0 may be reloaded once these MMAs have retired
      }
      __syncwarp();
      // ---- GEMM 1: Bottleneck conv1 over the y1 half of Y (channels [c, 2c): byte offset 2c inside the Y row)
      mbar_wait(y_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t koff = (uint32_t)(2 * c) >> 4;
        for (int j = 0; j < g.NB1; ++j) {
          const uint32_t d = tmem_base + (uint32_t)(g.d1_col + j * g.chid);
          const uint32_t aj = y16 + (((uint32_t)j * 128u * (uint32_t)g.rowY) >> 4) + koff;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t shift = (uint32_t)((t / 3) * P + (t % 3)) * (uint32_t)g.rowY;
            const uint64_t da = descY | (uint64_t)(aj + (shift >> 4));
            const uint64_t db = descW1 | (uint64_t)(((w1_base + t * g.w1_tile) & 0x3FFFFu) >> 4);
#pragma unroll
            for (int k = 0; k < KC; ++k) umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), g.idesc1, (t | k) ? 1u : 0u);
          }
        }
        umma_commit(d1_full);
      }
      __syncwarp();
      // ---- GEMM 2: Bottleneck conv2 over the hidden tile (chid == 16: one k-step per tap)
      mbar_wait(a2_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        for (int bq = 0; bq < g.MB; ++bq) {
          const uint32_t d = tmem_base + (uint32_t)(g.d2_col + bq * c);
          const uint32_t aj = a16 + (((uint32_t)(bq * g.TH) * P * (uint32_t)g.row2) >> 4);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t shift = (uint32_t)((t / 3) * P + (t % 3)) * (uint32_t)g.row2;
            const uint64_t da = desc2 | (uint64_t)(aj + (shift >> 4));
            const uint64_t db = desc2 | (uint64_t)(((w2_base + t * g.w2_tile) & 0x3FFFFu) >> 4);
            umma_bf16(d, da, db, g.idesc2, t ? 1u : 0u);
          }
        }
        umma_commit(d2_full);
      }
      __syncwarp();
      // ---- GEMM 3: cv2 over [y0 | y1] (Y, centre pixels: row shift 2P + 2) and y2 (Y2)
      mbar_wait(y2_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint64_t dba = descY | (uint64_t)((w3a_base & 0x3FFFFu) >> 4);
        const uint64_t dbb = descY2 | (uint64_t)((w3b_base & 0x3FFFFu) >> 4);
        for (int bq = 0; bq < g.MB; ++bq) {
          const uint32_t d = tmem_base + (uint32_t)(bq * g.cout);
          const uint32_t rowa = (uint32_t)(bq * g.TH) * P + 2u * P + 2u;
          const uint64_t daa = descY | (uint64_t)(y16 + ((rowa * (uint32_t)g.rowY) >> 4));
          const uint64_t dab = descY2 | (uint64_t)(y216 + (((uint32_t)bq * 128u * (uint32_t)g.rowY2) >> 4));
#pragma unroll
          for (int k = 0; k < 2 * KC; ++k) umma_bf16(d, daa + (uint64_t)(2 * k), dba + (uint64_t)(2 * k), g.idesc3, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < KC; ++k) umma_bf16(d, dab + (uint64_t)(2 * k), dbb + (uint64_t)(2 * k), g.idesc3, 1u);
        }
        umma_commit(d3_full);
      }
      __syncwarp();
    }
  } else {
    // ===================================================== epilogue (four phases): two warps per TMEM lane quadrant
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int P = g.P;
    const int th = m / P, tw = m - th * P;
    const bool row_ok = (th < g.TH) && (tw < g.TW);
    const uint32_t rr0 = (uint32_t)(th * g.TW + tw), blk_out_rows = (uint32_t)(g.TH * g.TW);
    const int nchY = (2 * c) >> 4, nch1 = g.chid >> 4, nch2 = c >> 4, nch3 = g.cout >> 4;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float *bias0 = bias_s, *bias1 = bias_s + 2 * c, *bias2 = bias1 + g.chid, *bias3 = bias2 + c;
    pdl_wait();
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int b = (int)(tile / tiles_per_img), r = (int)(tile % tiles_per_img);
      const int h0 = (r / g.tiles_w) * g.H2, w0 = (r % g.tiles_w) * g.TW;
      // ---------------- phase 0: [y0 | y1] -> Y (shared memory), zero outside the image
      mbar_wait(d0_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging tile (same memory) is free again
      named_bar_sync(1, kEpiThreads);
      const int n0 = g.NB0 * nchY;
      for (int f = half; f < n0; f += 2) {
        const int j = f / nchY, cc = f - j * nchY;
        float v[16];
        tmem_ld16_issue(taddr + (uint32_t)(j * 2 * c + cc * 16), v);
        tmem_ld_wait(v);
        const int q = j * 128 + m;  // halo pixel (ii - 2, jj - 2) relative to the tile origin
        const int ii = q / P, jj = q - ii * P;
        const int hi = h0 - 2 + ii, wi = w0 - 2 + jj;
        const bool inside = (hi >= 0) && (hi < g.H) && (wi >= 0) && (wi < g.W);
        silu16(v, bias0 + cc * 16, inside);
        store_chunk(y_base, (uint32_t)q * (uint32_t)g.rowY + (uint32_t)(cc * 32), (uint32_t)g.swY_mask, v);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // Y writes -> visible to the tensor core
      mbar_arrive(y_full);
      // ---------------- phase 1: hidden map -> A2
      mbar_wait(d1_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int n1 = g.NB1 * nch1;
      for (int f = half; f < n1; f += 2) {
        const int j = f / nch1, cc = f - j * nch1;
        float v[16];
        tmem_ld16_issue(taddr + (uint32_t)(g.d1_col + j * g.chid + cc * 16), v);
        tmem_ld_wait(v);
        const int q = j * 128 + m;  // hidden pixel (ii - 1, jj - 1) relative to the tile origin
        const int ii = q / P, jj = q - ii * P;
        const int hi = h0 - 1 + ii, wi = w0 - 1 + jj;
        const bool inside = (hi >= 0) && (hi < g.H) && (wi >= 0) && (wi < g.W);
        silu16(v, bias1 + cc * 16, inside);
        store_chunk(a2_base, (uint32_t)q * (uint32_t)g.row2 + (uint32_t)(cc * 32), (uint32_t)g.sw2_mask, v);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a2_full);
      // ---------------- phase 2: y2 = y1 + SiLU(conv2) -> Y2 (row bq*128 + m: no compaction)
      mbar_wait(d2_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int n2 = g.MB * nch2;
      for (int f = half; f < n2; f += 2) {
        const int bq = f / nch2, cc = f - bq * nch2;
        float v[16];
        tmem_ld16_issue(taddr + (uint32_t)(g.d2_col + bq * c + cc * 16), v);
        tmem_ld_wait(v);
        silu16(v, bias2 + cc * 16, true);
        if (row_ok) {  // residual y1(th, tw): Y row (bq*TH + th + 2)*P + tw + 2, channels [c + 16cc, c + 16cc + 16)
          uint32_t yo = (uint32_t)((bq * g.TH + th + 2) * P + tw + 2) * (uint32_t)g.rowY + (uint32_t)(2 * c + cc * 32);
          const uint32_t yx = ((yo >> 7) & (uint32_t)g.swY_mask) << 4;
          uint4 r0, r1;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(y_base + (yo ^ yx)));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(y_base + ((yo + 16u) ^ yx)));
          float rf[16];
          unpack8(r0, rf);
          unpack8(r1, rf + 8);
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] += rf[k];
        }
        store_chunk(y2_base, (uint32_t)(bq * 128 + m) * (uint32_t)g.rowY2 + (uint32_t)(cc * 32), (uint32_t)g.swY2_mask, v);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(y2_full);
      // ---------------- phase 3: output blocks -> staging (the memory of Y: GEMM 3 has retired) -> TMA store
      mbar_wait(d3_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int n3 = g.MB * nch3;
      for (int f = half; f < n3; f += 2) {
        const int bq = f / nch3, cc = f - bq * nch3;
        float v[16];
        tmem_ld16_issue(taddr + (uint32_t)(bq * g.cout + cc * 16), v);
        tmem_ld_wait(v);
        silu16(v, bias3 + cc * 16, true);
        if (row_ok) {
          const int col = cc * 16;
          const uint32_t rowoff = (rr0 + (uint32_t)bq * blk_out_rows) * 128u;  // staging rows: 64 channels = 128 bytes
          store_chunk(y_base + (uint32_t)(col >> 6) * (uint32_t)g.o_box_bytes, rowoff + (uint32_t)((col & 63) * 2), (uint32_t)g.o_sw_mask, v);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(d0_empty);  // D3 drained: GEMM 0 of the next tile may overwrite the columns
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      named_bar_sync(2, kEpiThreads);
      if (warp == 2) {
        if (elect_one()) {
          for (int bx = 0; bx < g.n_oboxes; ++bx) tma_store_4d(&tmO, y_base + (uint32_t)(bx * g.o_box_bytes), bx * 64, w0, h0, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
    }
    if (warp == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
  }
}

typedef void (*C3k2Fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                       const CUtensorMap, const CUtensorMap, const C3k2Geom);

C3k2Fn pick_c3k2(int ks0, int kc) {
  switch (ks0 * 8 + kc) {
    case 2 * 8 + 1: return c3k2_kernel<2, 1>;  // yolo11n layer 2: 32 -> 2 x 16 -> 64
    case 4 * 8 + 2: return c3k2_kernel<4, 2>;  // yolo11n layer 4: 64 -> 2 x 32 -> 128
    case 4 * 8 + 1: return c3k2_kernel<4, 1>;
    case 2 * 8 + 2: return c3k2_kernel<2, 2>;
    default: return nullptr;
  }
}

int sw_mask_for_row(int row_bytes) { return row_bytes == 128 ? 7 : (row_bytes == 64 ? 3 : 1); }
}  // namespace

// Tile geometry and shared-memory / TMEM carve-up (pure host arithmetic: also answers b2y_c3k2_fits for the planner).
static bool c3k2_geometry(const C3k2Params &p, int num_sms, C3k2Geom &g, size_t &smem_out) {
  memset(&g, 0, sizeof(g));
  g.B = p.B; g.H = p.H; g.W = p.W; g.cin = p.cin; g.c = p.c; g.chid = p.chid; g.cout = p.cout;
  g.row0 = p.cin * 2; g.rowY = 4 * p.c; g.row2 = p.chid * 2; g.rowY2 = 2 * p.c;
  g.sw0_mask = sw_mask_for_row(g.row0); g.swY_mask = sw_mask_for_row(g.rowY);
  g.sw2_mask = sw_mask_for_row(g.row2); g.swY2_mask = sw_mask_for_row(g.rowY2);
  g.o_sw_mask = 7;  // staging boxes are 64 channels = 128-byte rows
  g.bias = p.bias;
  {  // spatial block of GEMM 2 / GEMM 3: rows m = th*P + tw, P = TW + 4 (same search as the fused Bottleneck)
    double best = -1.0;
    int bth = 1, btw = 1;
    for (int tw = 1; tw <= (p.W < 124 ? p.W : 124); ++tw) {
      const int P = tw + 4;
      int th = (128 - tw) / P + 1;
      if (th > p.H) th = p.H;
      if (th < 1) continue;
      const double eff = (double)p.H * p.W / ((double)ceil_div(p.H, th) * ceil_div(p.W, tw) * 128.0);
      const double score = eff - 0.03 * ((double)(th + 4) * P / ((double)th * tw)) + 1e-4 * tw;
      if (score > best) { best = score; bth = th; btw = tw; }
    }
    g.TH = bth; g.TW = btw; g.P = btw + 4;
  }
  g.tiles_w = ceil_div(p.W, g.TW);
  g.n_oboxes = p.cout / 64;
  g.w1_tile = round_up(p.chid * g.rowY2, 1024);   // [chid][c] per tap
  g.w2_tile = round_up(p.c * g.row2, 1024);       // [c][chid] per tap
  const int w0_bytes = round_up(2 * p.c * g.row0, 1024);
  const int w3a_bytes = round_up(p.cout * g.rowY, 1024), w3b_bytes = round_up(p.cout * g.rowY2, 1024);
  g.w_tx_bytes = 2 * p.c * g.row0 + 9 * (p.chid * g.rowY2 + p.c * g.row2) + p.cout * g.rowY + p.cout * g.rowY2;
  const int bias_bytes = round_up((2 * p.c + p.chid + p.c + p.cout) * 4, 1024);
  int best_mb = 0;
  for (int pass = 0; pass < 2 && !best_mb; ++pass) {  // pass 0: two CTAs per SM; pass 1: one
    const int smem_budget = pass == 0 ? 113 * 1024 : 220 * 1024, tmem_budget = pass == 0 ? 256 : 512;
    for (int mb = 4; mb >= 1; mb >>= 1) {
      if (mb > 1 && (mb - 1) * g.TH >= p.H) continue;
      const int H2 = mb * g.TH, P = g.P;
      const int NB0 = ceil_div((H2 + 4) * P, 128), NB1 = ceil_div((H2 + 2) * P, 128);
      const int d03 = round_up(NB0 * 2 * p.c > mb * p.cout ? NB0 * 2 * p.c : mb * p.cout, 32);
      const int d1 = round_up(NB1 * p.chid, 32), d2 = round_up(mb * p.c, 32);
      if (d03 + d1 + d2 > tmem_budget) continue;
      const long long tiles = (long long)p.B * ceil_div(p.H, H2) * g.tiles_w;
      if (mb > 1 && tiles < 3LL * num_sms) continue;
      const int x_rows = NB0 * 128;  // GEMM 0 reads whole 128-row blocks (the TMA box fills (H2+4)*P of them)
      int y_rows = NB0 * 128;
      const int y_need1 = (NB1 - 1) * 128 + 2 * P + 2 + 128, y_need3 = (mb - 1) * g.TH * P + 2 * P + 2 + 128;
      if (y_need1 > y_rows) y_rows = y_need1;
      if (y_need3 > y_rows) y_rows = y_need3;
      const int a2_need = (mb - 1) * g.TH * P + 2 * P + 2 + 128;
      const int a2_rows = NB1 * 128 > a2_need ? NB1 * 128 : a2_need;
      const int x_bytes = round_up(x_rows * g.row0, 1024);
      int y_bytes = round_up(y_rows * g.rowY, 1024);
      const int o_bytes = g.n_oboxes * mb * 128 * 128;
      if (o_bytes > y_bytes) y_bytes = o_bytes;
      const int a2_bytes = round_up(a2_rows * g.row2, 1024), y2_bytes = round_up(mb * 128 * g.rowY2, 1024);
      const long long total = (long long)x_bytes + y_bytes + a2_bytes + y2_bytes + w0_bytes + 9LL * (g.w1_tile + g.w2_tile) +
                              w3a_bytes + w3b_bytes + bias_bytes + 1024;
      if (total > smem_budget) continue;
      best_mb = mb;
      g.MB = mb; g.H2 = H2; g.NB0 = NB0; g.NB1 = NB1;
      g.d1_col = d03; g.d2_col = d03 + d1;
      g.x_tx_bytes = (H2 + 4) * P * g.row0;
      g.o_box_bytes = mb * 128 * 128;
      g.x_off = 0;
      g.y_off = x_bytes;
      g.a2_off = g.y_off + y_bytes;
      g.y2_off = g.a2_off + a2_bytes;
      g.w0_off = g.y2_off + y2_bytes;
      g.w1_off = g.w0_off + w0_bytes;
      g.w2_off = g.w1_off + 9 * g.w1_tile;
      g.w3a_off = g.w2_off + 9 * g.w2_tile;
      g.w3b_off = g.w3a_off + w3a_bytes;
      g.bias_off = g.w3b_off + w3b_bytes;
      smem_out = (size_t)g.bias_off + bias_bytes + 1024;
      int cols = 32;
      while (cols < d03 + d1 + d2) cols <<= 1;
      g.tmem_cols = cols;
      break;
    }
  }
  if (!best_mb) return false;
  g.tiles_h = ceil_div(p.H, g.H2);
  g.ntiles = (long long)p.B * g.tiles_h * g.tiles_w;
  return g.ntiles < (1ll << 31);
}

struct C3k2Plan {
  CUtensorMap tmX, tmW0, tmW1, tmW2, tmW3a, tmW3b, tmO;
  C3k2Geom g;
  int grid;
  size_t smem;
  C3k2Fn fn;
};

bool c3k2_supported(const C3k2Params &p, char *why, int why_len) {
  auto no = [&](const char *msg) {
    if (why) snprintf(why, why_len, "%s", msg);
    return false;
  };
  if (p.cin != 32 && p.cin != 64) return no("cin must be 32 or 64 (one swizzled row per pixel)");
  if (p.c != 16 && p.c != 32) return no("c (half of cv1's output) must be 16 or 32");
  if (p.chid != 16) return no("Bottleneck hidden channels must pad to 16");
  if (p.cout != 64 && p.cout != 128) return no("cout must be 64 or 128");
  if (!pick_c3k2(p.cin / 16, p.c / 16)) return no("no kernel variant for this channel combination");
  if (p.in_cs % 8 || p.out_cs % 8 || (reinterpret_cast<uintptr_t>(p.in) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 15))
    return no("views not 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(p.w) & 15) return no("weights not 16-byte aligned");
  return true;
}

int c3k2_plan_create(const C3k2Params &p, int device, C3k2Plan **out) {
  char why[128];
  B2Y_REQUIRE(c3k2_supported(p, why, sizeof(why)), B2Y_ERR_UNSUPPORTED, "fused C3k2: %s", why);
  EncodeTiledFn encode = get_encode_fn();
  B2Y_REQUIRE(encode != nullptr, B2Y_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cudaDeviceProp prop;
  B2Y_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  const int num_sms = prop.multiProcessorCount;
  C3k2Plan *pl = new C3k2Plan();
  C3k2Geom &g = pl->g;
  if (!c3k2_geometry(p, num_sms, g, pl->smem)) {
    delete pl;
    b2y_set_error("fused C3k2: no tile configuration fits");
    return B2Y_ERR_UNSUPPORTED;
  }
  g.dhi0 = desc_hi_for_row(g.row0);
  g.dhiY = desc_hi_for_row(g.rowY);
  g.dhiW1 = desc_hi_for_row(g.rowY2);  // W1 tiles are [chid][c]: c*2 bytes per row
  g.dhi2 = desc_hi_for_row(g.row2);
  g.dhiY2 = desc_hi_for_row(g.rowY2);
  g.idesc0 = idesc_bf16(2 * p.c);
  g.idesc1 = idesc_bf16(p.chid);
  g.idesc2 = idesc_bf16(p.c);
  g.idesc3 = idesc_bf16(p.cout);
  // packed weights: W0 [2c][cin] | W1 [chid][9c] | W2 [c][9 chid] | W3 [cout][3c]   (bf16, K-major)
  const char *w = reinterpret_cast<const char *>(p.w);
  const char *w0 = w, *w1 = w0 + (size_t)2 * p.c * p.cin * 2, *w2 = w1 + (size_t)p.chid * 9 * p.c * 2;
  const char *w3 = w2 + (size_t)p.c * 9 * p.chid * 2;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  auto map2d = [&](CUtensorMap *tm, const char *base, int K, int rows, int boxK) {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)boxK, (cuuint32_t)rows};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<char *>(base), dims, strides, box, ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row(boxK * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.in_cs * 2, (cuuint64_t)p.W * p.in_cs * 2, (cuuint64_t)p.H * p.W * p.in_cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)p.cin, (cuuint32_t)g.P, (cuuint32_t)(g.H2 + 4), 1};
    r = encode(&pl->tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16 *>(p.in), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row(g.row0), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) r = map2d(&pl->tmW0, w0, p.cin, 2 * p.c, p.cin);
  if (r == CUDA_SUCCESS) r = map2d(&pl->tmW1, w1, 9 * p.c, p.chid, p.c);
  if (r == CUDA_SUCCESS) r = map2d(&pl->tmW2, w2, 9 * p.chid, p.c, p.chid);
  if (r == CUDA_SUCCESS) r = map2d(&pl->tmW3a, w3, 3 * p.c, p.cout, 2 * p.c);
  if (r == CUDA_SUCCESS) r = map2d(&pl->tmW3b, w3, 3 * p.c, p.cout, p.c);
  if (r == CUDA_SUCCESS) {
    const cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.out_cs * 2, (cuuint64_t)p.W * p.out_cs * 2, (cuuint64_t)p.H * p.W * p.out_cs * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)g.TW, (cuuint32_t)g.H2, 1};
    r = encode(&pl->tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p.out, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    delete pl;
    b2y_set_error("cuTensorMapEncodeTiled failed with CUresult %d (fused C3k2)", (int)r);
    return B2Y_ERR_CUDA;
  }
  const int per_sm = (pl->smem <= (size_t)114 * 1024 && g.tmem_cols <= 256) ? 2 : 1;
  const long long max_ctas = (long long)num_sms * per_sm;
  pl->grid = (int)(g.ntiles < max_ctas ? g.ntiles : max_ctas);
  pl->fn = pick_c3k2(p.cin / 16, p.c / 16);
  {
    const char *vb = getenv("B2Y_VERBOSE");
    if (vb && vb[0] == '1')
      fprintf(stderr, "[b2y c3k2] %dx%d cin %d c %d cout %d: TH %d TW %d MB %d NB0 %d NB1 %d tiles %lld grid %d smem %zu tmem %d\n",
              p.H, p.W, p.cin, p.c, p.cout, g.TH, g.TW, g.MB, g.NB0, g.NB1, g.ntiles, pl->grid, pl->smem, g.tmem_cols);
  }
  B2Y_CUDA_CHECK(cudaFuncSetAttribute(pl->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  *out = pl;
  return B2Y_OK;
}

void c3k2_plan_destroy(C3k2Plan *p) { delete p; }

int launch_c3k2(const C3k2Plan *pl, cudaStream_t st) {
  b2y_launch(pl->fn, dim3(pl->grid), dim3(kNumThreads), pl->smem, st, pl->tmX, pl->tmW0, pl->tmW1, pl->tmW2, pl->tmW3a,
             pl->tmW3b, pl->tmO, pl->g);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// Planner query (no CUDA calls): would a C3k2 block with these channels / map size get a fused-kernel configuration?
extern "C" int b2y_c3k2_fits(int32_t cin, int32_t c, int32_t chid, int32_t cout, int32_t B, int32_t H, int32_t W) {
  C3k2Params p;
  memset(&p, 0, sizeof(p));
  p.cin = cin; p.c = c; p.chid = chid; p.cout = cout; p.B = B; p.H = H; p.W = W; p.in_cs = cin; p.out_cs = cout;
  if (!c3k2_supported(p, nullptr, 0)) return 0;
  C3k2Geom g;
  size_t smem = 0;
  return c3k2_geometry(p, 148, g, smem) ? 1 : 0;
}

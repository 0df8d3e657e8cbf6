// Bottleneck as ONE kernel:  out = x + SiLU(conv3x3(SiLU(conv3x3(x; W1) + b1); W2) + b2)
// (ultralytics/nn/modules/block.py:3405-3419, BN folded per utils/torch_utils.py:242-269; c1 == c2 <= 64, hidden <= 64).
//
// The thin Bottlenecks of yolo11n (16->8->16 on 160x160 ... 64->32->64 on 40x40) are bound by per-launch latency and by the
// round trip of the hidden map through HBM.  Here a tile keeps everything on chip:
//   * TMA loads the input halo tile X: (H2+4) x P pixels (P = TW+4; OOB zero-fill = conv1's "same" padding).
//   * GEMM 1 (tcgen05, accumulators D1 in TMEM) evaluates conv1 on the (H2+2) x P halo region of the hidden map: hidden
//     pixel q = (i+1)*P + (j+1) is accumulator row q, its taps are X rows q + r*P + s (row-shifted descriptors).
//   * epilogue 1: D1 -> +b1 -> SiLU -> ZERO outside the image (conv2's padding) -> bf16 rows of the shared-memory tile A2
//     in the K-major swizzled layout GEMM 2 consumes (row q of A2 = hidden pixel q: no compaction, no HBM).
//   * GEMM 2 evaluates conv2 over A2 with the same shifted-descriptor scheme (output pixel m = th*P + tw, taps at m + r*P + s).
//   * epilogue 2: D2 -> +b2 -> SiLU -> + x (read back from the X tile in shared memory) -> staging -> TMA store.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2..9 epilogue (both phases).  Both weight
// matrices stay resident in shared memory; X is double buffered so the next tile's load overlaps this tile's math.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int kEpiWarps = 8, kEpiThreads = 32 * kEpiWarps, kNumThreads = 64 + kEpiThreads;

struct PairGeom {
  int B, H, W, cin, chid, cout;
  int TH, TW, P, MB, H2, NB1, tiles_h, tiles_w;
  long long ntiles;
  int row1, row2;                       // bytes per X row (cin*2) / per A2 row (chid*2)
  int x_stage_bytes, x_tx_bytes, a2_bytes, w1_tile, w2_tile, w_tx_bytes;
  int x_off, a2_off, w1_off, w2_off, o_off, bias_off;  // byte offsets in dynamic smem (x_off: 2 stages)
  int o_stage_bytes, oc, o_row_bytes, o_sw_mask, sw2_mask, sw1_mask;
  int tmem_cols, d2_col;                // D1 at column 0 (NB1 x chid), D2 at d2_col (MB x cout)
  unsigned desc_hi1, desc_hi2, idesc1, idesc2;
  const float *bias1, *bias2;
  int add;                              // residual x + ...
};

template <int KS1, int KS2>  // k-steps (16 channels each) of conv1 / conv2
__global__ void __launch_bounds__(kNumThreads, 2)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ PairGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t fullX = smem_u32(&bars[0]), emptyX = smem_u32(&bars[2]);  // 2 stages each
  const uint32_t wfull = smem_u32(&bars[4]);
  const uint32_t d1_full = smem_u32(&bars[5]), d1_empty = smem_u32(&bars[6]);
  const uint32_t a2_full = smem_u32(&bars[7]), a2_empty = smem_u32(&bars[8]);
  const uint32_t d2_full = smem_u32(&bars[9]), d2_empty = smem_u32(&bars[10]);
  float *bias_s = reinterpret_cast<float *>(smem_gen + g.bias_off);  // [chid | cout], pre-halved (SiLU form)

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(fullX + 8 * s, 1);
      mbar_init(emptyX + 8 * s, kEpiThreads);
    }
    mbar_init(wfull, 1);
    mbar_init(d1_full, 1);
    mbar_init(d1_empty, kEpiThreads);
    mbar_init(a2_full, kEpiThreads);
    mbar_init(a2_empty, 1);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < g.chid; i += kNumThreads) bias_s[i] = 0.5f * g.bias1[i];
  for (int i = threadIdx.x; i < g.cout; i += kNumThreads) bias_s[g.chid + i] = 0.5f * g.bias2[i];
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
  const uint32_t x_base = smem_base + (uint32_t)g.x_off, a2_base = smem_base + (uint32_t)g.a2_off;
  const uint32_t w1_base = smem_base + (uint32_t)g.w1_off, w2_base = smem_base + (uint32_t)g.w2_off;
  const int tiles_per_img = g.tiles_h * g.tiles_w;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
      mbar_expect_tx(wfull, (uint32_t)g.w_tx_bytes);
      for (int t = 0; t < 9; ++t) {
        tma_load_2d(w1_base + t * g.w1_tile, &tmW1, wfull, t * g.cin, 0);
        tma_load_2d(w2_base + t * g.w2_tile, &tmW2, wfull, t * g.chid, 0);
      }
    }
    __syncwarp();
    pdl_wait();
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const int b = (int)(tile / tiles_per_img), r = (int)(tile % tiles_per_img);
      const int h0 = (r / g.tiles_w) * g.H2, w0 = (r % g.tiles_w) * g.TW;
      mbar_wait(emptyX + 8 * s, ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(fullX + 8 * s, (uint32_t)g.x_tx_bytes);
        tma_load_4d(x_base + s * g.x_stage_bytes, &tmX, fullX + 8 * s, 0, w0 - 2, h0 - 2, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    const uint64_t desc1 = ((uint64_t)g.desc_hi1 << 32) | (uint64_t)(1u << 16);
    const uint64_t desc2 = ((uint64_t)g.desc_hi2 << 32) | (uint64_t)(1u << 16);
    const uint32_t P = (uint32_t)g.P;
    mbar_wait(wfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = it & 1;
      // ---- GEMM 1: hidden halo region
      mbar_wait(fullX + 8 * s, (it >> 1) & 1);
      mbar_wait(d1_empty, ph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t x16 = ((x_base + s * g.x_stage_bytes) & 0x3FFFFu) >> 4;
        for (int j = 0; j < g.NB1; ++j) {
          const uint32_t d = tmem_base + (uint32_t)(j * g.chid);
          const uint32_t aj = x16 + (((uint32_t)j * 128u * (uint32_t)g.row1) >> 4);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t shift = (uint32_t)((t / 3) * P + (t % 3)) * (uint32_t)g.row1;
            const uint64_t da = desc1 | (uint64_t)(aj + (shift >> 4));
            const uint64_t db = desc1 | (uint64_t)(((w1_base + t * g.w1_tile) & 0x3FFFFu) >> 4);
#pragma unroll
            for (int k = 0; k < KS1; ++k) umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), g.idesc1, (t | k) ? 1u : 0u);
          }
        }
        umma_commit(d1_full);
      }
      __syncwarp();
      // ---- GEMM 2: output blocks over the hidden tile written by epilogue 1
      mbar_wait(a2_full, ph);
      mbar_wait(d2_empty, ph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a16 = (a2_base & 0x3FFFFu) >> 4;
        for (int bq = 0; bq < g.MB; ++bq) {
          const uint32_t d = tmem_base + (uint32_t)(g.d2_col + bq * g.cout);
          const uint32_t aj = a16 + (((uint32_t)(bq * g.TH) * P * (uint32_t)g.row2) >> 4);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t shift = (uint32_t)((t / 3) * P + (t % 3)) * (uint32_t)g.row2;
            const uint64_t da = desc2 | (uint64_t)(aj + (shift >> 4));
            const uint64_t db = desc2 | (uint64_t)(((w2_base + t * g.w2_tile) & 0x3FFFFu) >> 4);
#pragma unroll
            for (int k = 0; k < KS2; ++k) umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), g.idesc2, (t | k) ? 1u : 0u);
          }
        }
        umma_commit(d2_full);
        umma_commit(a2_empty);  // A2 may be rewritten once these MMAs have retired
      }
      __syncwarp();
    }
  } else {
    // ===================================================== epilogue (both phases): two warps per TMEM lane quadrant
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int P = g.P;
    const int th = m / P, tw = m - th * P;
    const bool row_ok = (th < g.TH) && (tw < g.TW);
    const uint32_t rr0 = (uint32_t)(th * g.TW + tw), blk_out_rows = (uint32_t)(g.TH * g.TW);
    const uint32_t o_base = smem_base + (uint32_t)g.o_off;
    const int nch1 = g.chid >> 4, nch2 = g.cout >> 4;
    const int oc_shift = g.oc == 64 ? 6 : (g.oc == 32 ? 5 : 4);
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    pdl_wait();
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = it & 1;
      const int b = (int)(tile / tiles_per_img), r = (int)(tile % tiles_per_img);
      const int h0 = (r / g.tiles_w) * g.H2, w0 = (r % g.tiles_w) * g.TW;
      // ---------------- phase 1: hidden map -> A2 (shared memory)
      mbar_wait(d1_full, ph);
      mbar_wait(a2_empty, ph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int n1 = g.NB1 * nch1;
      for (int f = half; f < n1; f += 2) {
        const int j = f / nch1, c = f - j * nch1;
        float v[16];
        tmem_ld16_issue(taddr + (uint32_t)(j * g.chid + c * 16), v);
        tmem_ld_wait(v);
        const int q = j * 128 + m;            // hidden pixel (ii-1, jj-1) relative to the tile origin
        const int ii = q / P, jj = q - ii * P;
        const int hi = h0 - 1 + ii, wi = w0 - 1 + jj;
        const bool inside = (hi >= 0) && (hi < g.H) && (wi >= 0) && (wi < g.W);
        const float4 *bp = reinterpret_cast<const float4 *>(bias_s + c * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 hb = bp[k];
          v[4 * k] = inside ? silu_half(fmaf(v[4 * k], 0.5f, hb.x)) : 0.f;
          v[4 * k + 1] = inside ? silu_half(fmaf(v[4 * k + 1], 0.5f, hb.y)) : 0.f;
          v[4 * k + 2] = inside ? silu_half(fmaf(v[4 * k + 2], 0.5f, hb.z)) : 0.f;
          v[4 * k + 3] = inside ? silu_half(fmaf(v[4 * k + 3], 0.5f, hb.w)) : 0.f;
        }
        uint32_t off = (uint32_t)q * (uint32_t)g.row2 + (uint32_t)(c * 32);
        const uint32_t xr = ((off >> 7) & (uint32_t)g.sw2_mask) << 4;
        const uint4 p0 = pack8(v), p1 = pack8(v + 8);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a2_base + (off ^ xr)), "r"(p0.x), "r"(p0.y), "r"(p0.z), "r"(p0.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a2_base + ((off + 16u) ^ xr)), "r"(p1.x), "r"(p1.y), "r"(p1.z), "r"(p1.w) : "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(d1_empty);                                          // D1 drained: GEMM 1 of the next tile may start
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A2 writes -> visible to the tensor core
      mbar_arrive(a2_full);
      // ---------------- phase 2: output blocks
      mbar_wait(d2_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging buffer free again
      named_bar_sync(1, kEpiThreads);
      const uint32_t x_stage = x_base + (uint32_t)(s * g.x_stage_bytes);
      const int n2 = g.MB * nch2;
      for (int f = half; f < n2; f += 2) {
        const int bq = f / nch2, c = f - bq * nch2;
        float v[16];
        tmem_ld16_issue(taddr + (uint32_t)(g.d2_col + bq * g.cout + c * 16), v);
        tmem_ld_wait(v);
        const float4 *bp = reinterpret_cast<const float4 *>(bias_s + g.chid + c * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 hb = bp[k];
          v[4 * k] = silu_half(fmaf(v[4 * k], 0.5f, hb.x));
          v[4 * k + 1] = silu_half(fmaf(v[4 * k + 1], 0.5f, hb.y));
          v[4 * k + 2] = silu_half(fmaf(v[4 * k + 2], 0.5f, hb.z));
          v[4 * k + 3] = silu_half(fmaf(v[4 * k + 3], 0.5f, hb.w));
        }
        if (row_ok) {
          if (g.add) {  // residual: x(th, tw) sits at X row (bq*TH + th + 2)*P + tw + 2 of this tile's stage
            uint32_t xo = (uint32_t)((bq * g.TH + th + 2) * P + tw + 2) * (uint32_t)g.row1 + (uint32_t)(c * 32);
            const uint32_t xx = ((xo >> 7) & (uint32_t)g.sw1_mask) << 4;
            uint4 r0, r1;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(x_stage + (xo ^ xx)));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(x_stage + ((xo + 16u) ^ xx)));
            float rf[16];
            unpack8(r0, rf);
            unpack8(r1, rf + 8);
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] += rf[k];
          }
          const int col = c * 16;
          const uint32_t rowoff = (rr0 + (uint32_t)bq * blk_out_rows) * (uint32_t)g.o_row_bytes;
          const uint32_t xr = ((rowoff >> 7) & (uint32_t)g.o_sw_mask) << 4;
          const uint32_t dst = o_base + rowoff + (uint32_t)(col >> oc_shift) * (uint32_t)g.o_stage_bytes;
          const uint32_t u = ((uint32_t)col & ((uint32_t)g.oc - 1u)) * 2u;
          const uint4 p0 = pack8(v), p1 = pack8(v + 8);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (u ^ xr)), "r"(p0.x), "r"(p0.y), "r"(p0.z), "r"(p0.w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((u + 16u) ^ xr)), "r"(p1.x), "r"(p1.y), "r"(p1.z), "r"(p1.w) : "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(d2_empty);        // D2 drained
      mbar_arrive(emptyX + 8 * s);  // residual rows read: the X stage may be reloaded
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      named_bar_sync(2, kEpiThreads);
      if (warp == 2) {
        if (elect_one()) {
          tma_store_4d(&tmO, o_base, 0, w0, h0, b);
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

typedef void (*PairFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const PairGeom);

PairFn pick_pair(int ks1, int ks2) {
  switch (ks1 * 8 + ks2) {
    case 1 * 8 + 1: return conv_pair_kernel<1, 1>;
    case 2 * 8 + 1: return conv_pair_kernel<2, 1>;
    case 2 * 8 + 2: return conv_pair_kernel<2, 2>;
    case 4 * 8 + 2: return conv_pair_kernel<4, 2>;
    case 4 * 8 + 4: return conv_pair_kernel<4, 4>;
    default: return nullptr;
  }
}
}  // namespace

// Tile geometry and shared-memory / TMEM carve-up (pure host arithmetic: also answers b2y_conv_pair_fits for the planner).
static bool pair_geometry(const PairParams &p, int num_sms, PairGeom &g, size_t &smem_out) {
  memset(&g, 0, sizeof(g));
  g.B = p.B; g.H = p.H; g.W = p.W; g.cin = p.cin; g.chid = p.chid; g.cout = p.cout;
  g.row1 = p.cin * 2; g.row2 = p.chid * 2;
  g.bias1 = p.b1; g.bias2 = p.b2; g.add = p.add;
  // spatial block of GEMM 2: rows m = th*P + tw, P = TW + 4
  {
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
  g.oc = p.cout <= 16 ? 16 : (p.cout <= 32 ? 32 : 64);
  g.o_row_bytes = g.oc * 2;
  g.o_sw_mask = g.oc == 64 ? 7 : (g.oc == 32 ? 3 : 1);
  g.sw1_mask = g.row1 == 128 ? 7 : (g.row1 == 64 ? 3 : 1);
  g.sw2_mask = g.row2 == 128 ? 7 : (g.row2 == 64 ? 3 : 1);
  g.w1_tile = round_up(p.chid * g.row1, 1024);
  g.w2_tile = round_up(p.cout * g.row2, 1024);
  g.w_tx_bytes = 9 * (p.chid * g.row1 + p.cout * g.row2);
  const int bias_bytes = round_up((p.chid + p.cout) * 4, 1024);
  const int smem_budget = 110 * 1024;  // two CTAs per SM
  int best_mb = 0;
  for (int mb = 4; mb >= 1; mb >>= 1) {
    if (mb > 1 && (mb - 1) * g.TH >= p.H) continue;
    const int H2 = mb * g.TH;
    const int R1 = (H2 + 2) * g.P, NB1 = ceil_div(R1, 128);
    const int d2_col = round_up(NB1 * p.chid, 32);
    if (d2_col + mb * p.cout > 256) continue;  // TMEM: two co-resident CTAs
    const long long tiles = (long long)p.B * ceil_div(p.H, H2) * g.tiles_w;
    if (mb > 1 && tiles < 3LL * num_sms) continue;
    const int x_rows = ((H2 + 4) * g.P > NB1 * 128 + 2 * g.P + 2) ? (H2 + 4) * g.P : NB1 * 128 + 2 * g.P + 2;
    const int x_stage = round_up(x_rows * g.row1, 1024);
    const int a2_need = (mb - 1) * g.TH * g.P + 2 * g.P + 2 + 128;
    const int a2_rows = NB1 * 128 > a2_need ? NB1 * 128 : a2_need;
    const int a2_bytes = round_up(a2_rows * g.row2, 1024);
    const int o_stage = mb * 128 * g.o_row_bytes;  // cout <= 64: one channel box
    const long long total = 2LL * x_stage + a2_bytes + 9LL * (g.w1_tile + g.w2_tile) + o_stage + bias_bytes + 1024;
    if (total > smem_budget && mb > 1) continue;
    if (total > 220 * 1024) continue;
    best_mb = mb;
    g.MB = mb; g.H2 = H2; g.NB1 = NB1; g.d2_col = d2_col;
    g.x_stage_bytes = x_stage;
    g.x_tx_bytes = (H2 + 4) * g.P * g.row1;
    g.a2_bytes = a2_bytes;
    g.o_stage_bytes = o_stage;
    g.x_off = 0;
    g.a2_off = 2 * x_stage;
    g.w1_off = g.a2_off + a2_bytes;
    g.w2_off = g.w1_off + 9 * g.w1_tile;
    g.o_off = g.w2_off + 9 * g.w2_tile;
    g.bias_off = g.o_off + o_stage;
    smem_out = (size_t)g.bias_off + bias_bytes + 1024;
    break;
  }
  if (!best_mb) return false;
  g.tiles_h = ceil_div(p.H, g.H2);
  g.ntiles = (long long)p.B * g.tiles_h * g.tiles_w;
  return g.ntiles < (1ll << 31);
}

struct ConvPairPlan {
  CUtensorMap tmX, tmW1, tmW2, tmO;
  PairGeom g;
  int grid;
  size_t smem;
  PairFn fn;
};

bool conv_pair_supported(const PairParams &p, char *why, int why_len) {
  auto no = [&](const char *msg) {
    if (why) snprintf(why, why_len, "%s", msg);
    return false;
  };
  if (p.cin != p.cout) return no("fused Bottleneck needs c1 == c2");
  if (p.cin % 16 || p.chid % 16 || p.cin > 64 || p.chid > 64) return no("channels must be multiples of 16, at most 64");
  if (p.cin != 16 && p.cin != 32 && p.cin != 64) return no("cin must be 16, 32 or 64 (one swizzled row per pixel)");
  if (p.chid != 16 && p.chid != 32 && p.chid != 64) return no("hidden channels must be 16, 32 or 64");
  if (!pick_pair(p.cin / 16, p.chid / 16)) return no("no kernel variant for this channel pair");
  if (p.in_cs % 8 || p.out_cs % 8 || (reinterpret_cast<uintptr_t>(p.in) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 15))
    return no("views not 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(p.w1) & 15) || (reinterpret_cast<uintptr_t>(p.w2) & 15)) return no("weights not 16-byte aligned");
  return true;
}

int conv_pair_plan_create(const PairParams &p, int device, ConvPairPlan **out) {
  char why[128];
  B2Y_REQUIRE(conv_pair_supported(p, why, sizeof(why)), B2Y_ERR_UNSUPPORTED, "fused Bottleneck: %s", why);
  EncodeTiledFn encode = get_encode_fn();
  B2Y_REQUIRE(encode != nullptr, B2Y_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cudaDeviceProp prop;
  B2Y_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  const int num_sms = prop.multiProcessorCount;
  ConvPairPlan *pl = new ConvPairPlan();
  PairGeom &g = pl->g;
  if (!pair_geometry(p, num_sms, g, pl->smem)) {
    delete pl;
    b2y_set_error("fused Bottleneck: no tile configuration fits");
    return B2Y_ERR_UNSUPPORTED;
  }
  int cols = 32;
  while (cols < g.d2_col + g.MB * p.cout) cols <<= 1;
  g.tmem_cols = cols;
  g.desc_hi1 = desc_hi_for_row(g.row1);
  g.desc_hi2 = desc_hi_for_row(g.row2);
  g.idesc1 = idesc_bf16(p.chid);
  g.idesc2 = idesc_bf16(p.cout);
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  CUresult r;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.in_cs * 2, (cuuint64_t)p.W * p.in_cs * 2, (cuuint64_t)p.H * p.W * p.in_cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)p.cin, (cuuint32_t)g.P, (cuuint32_t)(g.H2 + 4), 1};
    r = encode(&pl->tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16 *>(p.in), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row(g.row1), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) {
    const cuuint64_t dims[2] = {(cuuint64_t)9 * p.cin, (cuuint64_t)p.chid};
    const cuuint64_t strides[1] = {(cuuint64_t)9 * p.cin * 2};
    const cuuint32_t box[2] = {(cuuint32_t)p.cin, (cuuint32_t)p.chid};
    r = encode(&pl->tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(p.w1), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row(g.row1), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) {
    const cuuint64_t dims[2] = {(cuuint64_t)9 * p.chid, (cuuint64_t)p.cout};
    const cuuint64_t strides[1] = {(cuuint64_t)9 * p.chid * 2};
    const cuuint32_t box[2] = {(cuuint32_t)p.chid, (cuuint32_t)p.cout};
    r = encode(&pl->tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(p.w2), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row(g.row2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) {
    const cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.out_cs * 2, (cuuint64_t)p.W * p.out_cs * 2, (cuuint64_t)p.H * p.W * p.out_cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)g.oc, (cuuint32_t)g.TW, (cuuint32_t)g.H2, 1};
    const CUtensorMapSwizzle osw = g.oc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : g.oc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    r = encode(&pl->tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p.out, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               osw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    delete pl;
    b2y_set_error("cuTensorMapEncodeTiled failed with CUresult %d (fused Bottleneck)", (int)r);
    return B2Y_ERR_CUDA;
  }
  const int per_sm = pl->smem <= (size_t)110 * 1024 ? 2 : 1;
  const long long max_ctas = (long long)num_sms * per_sm;
  pl->grid = (int)(g.ntiles < max_ctas ? g.ntiles : max_ctas);
  pl->fn = pick_pair(p.cin / 16, p.chid / 16);
  {
    const char *vb = getenv("B2Y_VERBOSE");
    if (vb && vb[0] == '1')
      fprintf(stderr, "[b2y pair] %dx%d c %d hid %d: TH %d TW %d MB %d NB1 %d tiles %lld grid %d smem %zu tmem %d\n", p.H, p.W,
              p.cin, p.chid, g.TH, g.TW, g.MB, g.NB1, g.ntiles, pl->grid, pl->smem, g.tmem_cols);
  }
  B2Y_CUDA_CHECK(cudaFuncSetAttribute(pl->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  *out = pl;
  return B2Y_OK;
}

void conv_pair_plan_destroy(ConvPairPlan *p) { delete p; }

int launch_conv_pair(const ConvPairPlan *pl, cudaStream_t st) {
  b2y_launch(pl->fn, dim3(pl->grid), dim3(kNumThreads), pl->smem, st, pl->tmX, pl->tmW1, pl->tmW2, pl->tmO, pl->g);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// Planner query (no CUDA calls): would a Bottleneck with these channels / map size get a fused-kernel configuration?
extern "C" int b2y_conv_pair_fits(int32_t cin, int32_t chid, int32_t B, int32_t H, int32_t W) {
  PairParams p;
  memset(&p, 0, sizeof(p));
  p.cin = cin; p.chid = chid; p.cout = cin; p.B = B; p.H = H; p.W = W; p.in_cs = cin; p.out_cs = cin; p.add = 1;
  if (!conv_pair_supported(p, nullptr, 0)) return 0;
  PairGeom g;
  size_t smem = 0;
  return pair_geometry(p, 148, g, smem) ? 1 : 0;
}

// Dense convolution as an implicit GEMM on the 5th-generation tensor cores (sm_100a):
//   D[pixels, cout] = sum over taps (r,s) and channel chunks of  A_tap[pixels, BK] * W[cout, tap*cin + chunk]^T
//
//   * A operand: NHWC bf16 activations fetched by TMA straight out of the (possibly channel-sliced) activation
//     buffer, ONE halo box per (tile, channel chunk): {BK channels, TW+2, rows+2} for 3x3 s1, {BK, TW, rows} for 1x1.
//     Out-of-bounds coordinates are zero-filled by the TMA unit, which *is* the "same" padding (conv.py:29-35).
//     The nine taps are not nine loads: they are nine *row-shifted shared-memory descriptors* over the same halo tile.
//     Accumulator row m corresponds to halo pixel (m / P, m % P) with pitch P = TW+2, so tap (r,s) is the operand
//     starting r*P+s rows further down; the P-TW garbage columns per row are simply not stored.  Stride-2 3x3 convs
//     view the input as {2*Cs, W/2, 2, H/2, B}: four parity planes, each a dense halo box, each tap = (plane, shift).
//     No im2col buffer, no gather, L2->SM traffic ~1.3x the input instead of 9x.
//   * A tile is MB stacked 128-row accumulator blocks (MB*TH output rows x TW columns) fed from ONE halo load and
//     drained by ONE TMA store: the thin, memory-bound layers of yolo11n (16..64 channels on 160x160 / 80x80 maps)
//     amortise every barrier round trip over MB blocks.
//   * B operand: BN-folded weights packed [cout_pad][kh*kw*cin] bf16 (K-major).  When the whole matrix fits it is
//     loaded ONCE per persistent CTA and stays resident in shared memory; otherwise it is streamed per (chunk, tap).
//   * Both operands sit in shared memory in the canonical K-major swizzled layout (128B/64B/32B swizzle by BK) that
//     tcgen05.mma consumes through shared-memory descriptors (row-shifted views are plain start-address offsets: the
//     swizzle XOR acts on absolute smem address bits, measured); accumulators (128 lanes x MB*BN fp32 columns, double
//     buffered) live in TMEM.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one lane, fully unrolled descriptor arithmetic) + TMEM
//     allocator, warps 2..9 = epilogue (tcgen05.ld -> +bias -> SiLU -> residual add|mul -> bf16 into a swizzled smem
//     staging tile -> TMA store into the output channel slice; fp32 head outputs are stored from registers).
//     mbarrier pipelines: A full/empty, B full/empty (or one "weights resident" barrier), TMEM full/empty.
//     Persistent CTAs walk the tile list.
// Reference semantics: ultralytics/nn/modules/conv.py:145-151 (act(bn(conv(x))), BN folded per
// utils/torch_utils.py:242-269), residuals of Bottleneck (block.py:3416-3419), PSABlock (:4379-4383), MSCA (:88).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kNumThreads = 64 + kEpiThreads;
constexpr int kMaxStages = 8;
constexpr int kMaxMB = 8;

struct TcGeom {
  int B, Ho, Wo, cin, cout, cout_pad;
  int taps, nplanes, stride2, in_cs;
  int BK, nkc;
  int TH, TW, P, MB, tiles_h, tiles_w;  // TH: output rows per 128-row block; a tile has MB blocks (MB*TH rows)
  int BN, n_ntiles;
  long long ntiles;
  int sa, sb, resident;                 // A stages, B stages (streamed), weights resident flag
  int plane_bytes, a_stage_bytes, a_tx_bytes, b_tile_bytes, b_tx_bytes, w_bytes_tx;
  int row_bytes, blk_rows_bytes;        // bytes per halo row; byte offset between consecutive blocks' A operands
  int tap_off16[9];                     // tap sequence position -> operand start offset inside its A stage, 16-byte units
  int tap_seq[9];                       // tap sequence position -> tap index r*kw+s (stride 2: grouped by parity plane)
  int tmem_cols;
  unsigned desc_hi;                     // descriptor bits [32,64): SBO, version, layout type
  unsigned idesc;
  int ox, oy;                           // halo origin offsets: TMA coordinate = tile origin - (ox, oy)
  // epilogue
  void *out; int out_cs, out_f32;
  const bf16 *res; int res_cs, res_mode;
  const float *bias;
  int act, vec_ok;
  int grouped;                          // groups of 16 channels: chunk kc only feeds accumulator columns [16kc, 16kc+16)
  int flat, real_hw, real_w;            // 1x1 convs run over the flattened pixel axis; real map size for the fused decode
  unsigned long long *trace;            // debug (B2Y_TRACE_OP): CTA 0 logs %globaltimer at pipeline events, 3 x 64 slots
  int fast_epi;                         // bf16 TMA-store output, whole 16-column chunks, vectorisable residual: lean epilogue
  unsigned dv_nt_mul, dv_nt_shr, dv_tw_mul, dv_tw_shr, dv_th_mul, dv_th_shr;  // tile decode by multiply-shift
  // bf16 outputs leave through shared memory + TMA store
  int tma_out, oc, n_oboxes, obox_bytes, o_row_bytes, o_sw_mask, o_stages, o_stage_bytes;
  int o_off, bias_off;                  // byte offsets of the output staging area / bias table in dynamic smem
  // fused Detect decode (head.py:100-131): epi_mode 1 = DFL + dist2bbox of the box branch, 2 = class sigmoid; both
  // write straight into y (B, 4+nc, A) fp32 -- the raw head maps are never materialised
  int epi_mode, a_off, A, nc;
  float det_stride;
};

__device__ __forceinline__ void trace_evt(const TcGeom &g, int role, int &slot) {
  if (g.trace != nullptr && blockIdx.x == 0 && slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g.trace[role * 64 + slot++] = t;
  }
}

struct TileCoord {
  int b, th_i, tw_i, nt;
};
__device__ __forceinline__ TileCoord decode_tile(const TcGeom &g, uint32_t tile) {
  TileCoord t;
  const uint32_t sp = fast_div(tile, g.dv_nt_mul, g.dv_nt_shr);
  t.nt = (int)(tile - sp * g.n_ntiles);
  const uint32_t q = fast_div(sp, g.dv_tw_mul, g.dv_tw_shr);
  t.tw_i = (int)(sp - q * g.tiles_w);
  const uint32_t q2 = fast_div(q, g.dv_th_mul, g.dv_th_shr);
  t.th_i = (int)(q - q2 * g.tiles_h);
  t.b = (int)q2;
  return t;
}

// Shared-memory operand descriptors (K-major, swizzled): bits [0,14) start address >> 4, [16,30) LBO (unused for
// swizzled K-major, 1), [32,46) SBO >> 4 = 8 rows, [46,48) version 1, [61,64) swizzle mode.  Measured on B200: the
// swizzle XOR is applied to the absolute shared-memory address bits (exactly like the TMA write side), so a view that
// starts at any 16-byte-aligned row of a swizzle-aligned tile needs NO base_offset (bits [49,52) stay 0); setting
// base_offset = (addr >> 7) & 7 for shifted starts produces garbage (profiles/r01_notes.md).
// Stride-2 3x3 convs (S2): the four parity planes of a tile are four separate pipeline stages (taps grouped by plane:
// {centre}, {left,right}, {up,down}, {corners}), so the TMA producer runs several planes ahead of the MMAs instead of
// the whole 4-plane tile being one (single-buffered) stage.
// EPI selects the epilogue compiled into the instantiation (keeps each kernel's code small: the thin layers run a code
// path once or twice per launch, so instruction-cache misses are a first-order cost): 0 generic (fp32 / ragged / unaligned
// outputs), 1 lean bf16 (staging + TMA store), 2 lean fused Detect decode.
template <int TAPS, int KSTEPS, bool S2, int EPI>
__global__ void __launch_bounds__(kNumThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const TcGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 5];
  __shared__ uint32_t tmem_base_smem;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t fullA = smem_u32(&bars[0]), emptyA = smem_u32(&bars[kMaxStages]);
  const uint32_t fullB = smem_u32(&bars[2 * kMaxStages]), emptyB = smem_u32(&bars[3 * kMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[4 * kMaxStages]), tempty0 = smem_u32(&bars[4 * kMaxStages + 2]);
  const uint32_t wfull = smem_u32(&bars[4 * kMaxStages + 4]);
  // smem carve-up: [A ring: sa stages][B: resident weights (taps*nkc tiles) | ring of sb tiles][out staging][bias]
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)(g.sa * g.a_stage_bytes);
  float *bias_s = reinterpret_cast<float *>(smem_gen + g.bias_off);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(fullA + 8 * s, 1);
      mbar_init(emptyA + 8 * s, 1);
      mbar_init(fullB + 8 * s, 1);
      mbar_init(emptyB + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, EPI == 1 ? kEpiWarps : kEpiThreads);  // lean epilogue: one arrival per warp
    }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // bias table (pre-halved when the activation is SiLU: silu(x+b) = h + h*tanh(h), h = 0.5x + 0.5b)
  for (int i = threadIdx.x; i < g.cout_pad; i += kNumThreads) bias_s[i] = g.act ? 0.5f * g.bias[i] : g.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // PDL: this CTA owns its shared memory and TMEM columns now, so the next kernel of the stream may become resident and
  // run its own prologue under our main loop (triggering before the TMEM allocation could let a later kernel starve it)
  pdl_launch_dependents();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t ntiles = (uint32_t)g.ntiles;
  const int rows_tile = g.MB * g.TH;

  if (warp == 0) {
    // ===================================================== TMA producer
    // the whole warp walks the tile list (converged waits); one elected lane issues the TMA instructions
    {
      if (elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        if (g.resident) {
          mbar_expect_tx(wfull, (uint32_t)g.w_bytes_tx);
          for (int t = 0; t < TAPS * g.nkc; ++t)  // tile (tap, chunk); grouped: only the chunk's own 16 output rows
            tma_load_2d(b_base + t * g.b_tile_bytes, &tmB, wfull, t * g.BK, g.grouped ? (t % g.nkc) * 16 : 0);
        }
      }
      __syncwarp();
      int tslot = 0;
      if (lane == 0) trace_evt(g, 0, tslot);  // [0] producer start (prologue done)
      pdl_wait();  // weights are constants; the activations below were written by the preceding kernel(s)
      if (lane == 0) trace_evt(g, 0, tslot);  // [1] dependency satisfied
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(g, tile);
        const int h0 = tc.th_i * rows_tile - g.oy, w0 = tc.tw_i * g.TW - g.ox, n0 = tc.nt * g.BN;
        for (int kc = 0; kc < g.nkc; ++kc) {
          const int c0 = kc * g.BK;
#pragma unroll
          for (int sub = 0; sub < (S2 ? 4 : 1); ++sub) {
            mbar_wait(emptyA + 8 * sa, pha ^ 1);
            if (kc == 0 && sub == 0 && lane == 0) trace_evt(g, 0, tslot);  // [2+i] first load of tile i issued
            const uint32_t dst = a_base + sa * g.a_stage_bytes;
            if (elect_one()) {
              mbar_expect_tx(fullA + 8 * sa, (uint32_t)g.a_tx_bytes);
              if (!S2) tma_load_4d(dst, &tmA, fullA + 8 * sa, c0, w0, h0, tc.b);
              else tma_load_5d(dst, &tmA, fullA + 8 * sa, (sub & 1) * g.in_cs + c0, w0, sub >> 1, h0, tc.b);
            }
            __syncwarp();
            if (++sa == g.sa) { sa = 0; pha ^= 1; }
            if (!g.resident) {  // this stage's weight tiles, in the tap order the MMA warp consumes
              constexpr int kT0[5] = {0, 1, 3, 5, 9};
              const int t0 = S2 ? kT0[sub] : 0, t1 = S2 ? kT0[sub + 1] : TAPS;
              for (int t = t0; t < t1; ++t) {
                mbar_wait(emptyB + 8 * sb, phb ^ 1);
                if (elect_one()) {
                  mbar_expect_tx(fullB + 8 * sb, (uint32_t)g.b_tx_bytes);
                  tma_load_2d(b_base + sb * g.b_tile_bytes, &tmB, fullB + 8 * sb, g.tap_seq[t] * g.cin + c0, n0);
                }
                __syncwarp();
                if (++sb == g.sb) { sb = 0; phb ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    int sa = 0, sb = 0, as = 0;
    uint32_t pha = 0, phb = 0, aphase = 0;
    uint32_t tap_off[TAPS], tap_b[TAPS];  // registers: the tap loops below are fully unrolled
    const uint32_t btile16 = (uint32_t)g.b_tile_bytes >> 4;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      tap_off[t] = (uint32_t)g.tap_off16[t];
      tap_b[t] = (uint32_t)(g.tap_seq[t] * g.nkc) * btile16;  // resident weights: tile (tap, chunk) sits at tap*nkc + chunk
    }
    const uint64_t desc0 = ((uint64_t)g.desc_hi << 32) | (uint64_t)(1u << 16);
    const uint32_t blk16 = (uint32_t)g.blk_rows_bytes >> 4;
    const uint32_t idesc = g.idesc;
    const int MB = g.MB, nkc = g.nkc;
    if (g.resident) {
      mbar_wait(wfull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    int tslot = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(tempty0 + 8 * as, aphase ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (g.trace != nullptr && lane == 0) trace_evt(g, 1, tslot);  // [2i] accumulator stage free
      const uint32_t tmem_d0 = tmem_base + (uint32_t)(as * MB * g.BN);
      for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
        for (int sub = 0; sub < (S2 ? 4 : 1); ++sub) {
          // taps served by this A stage: all of them, or one parity plane's {1, 2, 2, 4}
          constexpr int kT0[4] = {0, 1, 3, 5};
          const int t0 = S2 ? kT0[sub] : 0;
          const int nt = S2 ? (sub == 0 ? 1 : (sub == 3 ? 4 : 2)) : TAPS;
          mbar_wait(fullA + 8 * sa, pha);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a16 = ((a_base + sa * g.a_stage_bytes) & 0x3FFFFu) >> 4;
          if (g.resident) {
            if (elect_one()) {
              const uint32_t b16 = ((b_base & 0x3FFFFu) >> 4) + (uint32_t)kc * btile16;
              const uint32_t gcol = g.grouped ? (uint32_t)kc * 16u : 0u;      // grouped: this chunk's own 16 columns,
              const uint32_t acc0 = g.grouped ? 0u : (uint32_t)(kc != 0);     // each started afresh by its first tap
              for (int j = 0; j < MB; ++j) {
                const uint32_t tmem_d = tmem_d0 + (uint32_t)(j * g.BN) + gcol;
                const uint32_t aj = a16 + (uint32_t)j * blk16;
#pragma unroll
                for (int q = 0; q < TAPS; ++q) {
                  if (q >= nt) break;
                  const int t = t0 + q;
                  const uint64_t da = desc0 | (uint64_t)(aj + tap_off[t]);
                  const uint64_t db = desc0 | (uint64_t)(b16 + tap_b[t]);
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (t | k) ? 1u : acc0);
                }
              }
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int q = 0; q < nt; ++q) {
              const int t = t0 + q;
              mbar_wait(fullB + 8 * sb, phb);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
                const uint64_t db = desc0 | (uint64_t)(((b_base + sb * g.b_tile_bytes) & 0x3FFFFu) >> 4);
                for (int j = 0; j < MB; ++j) {
                  const uint64_t da = desc0 | (uint64_t)(a16 + (uint32_t)j * blk16 + (uint32_t)g.tap_off16[t]);
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    umma_bf16(tmem_d0 + (uint32_t)(j * g.BN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                              (uint32_t)((kc | t | k) != 0));
                }
                umma_commit(emptyB + 8 * sb);
              }
              __syncwarp();
              if (++sb == g.sb) { sb = 0; phb ^= 1; }
            }
          }
          if (elect_one()) {
            umma_commit(emptyA + 8 * sa);  // frees the halo tile / plane when the MMAs retire
            if (kc == nkc - 1 && sub == (S2 ? 3 : 0)) {
              umma_commit(tfull0 + 8 * as);  // accumulators complete
              trace_evt(g, 1, tslot);        // [2i+1] all MMAs of tile i issued
            }
          }
          __syncwarp();
          if (++sa == g.sa) { sa = 0; pha ^= 1; }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================================================== epilogue: 8 warps, two per TMEM lane quadrant (warp % 4)
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;  // accumulator row == halo pixel index within the block
    const int th = m / g.P, tw = m - th * g.P;
    const bool row_ok = (th < g.TH) && (tw < g.TW);
    const uint32_t rr0 = (uint32_t)(th * g.TW + tw);  // row in the compacted output block (halo columns dropped)
    const uint32_t blk_out_rows = (uint32_t)(g.TH * g.TW);
    const bool elected = (threadIdx.x == 64);
    const uint32_t o_base = smem_base + (uint32_t)g.o_off;
    const int nch = g.BN >> 4, nflat = g.MB * nch;
    int as = 0, os = 0;
    uint32_t aphase = 0;
    float dfl_first = 0.f;
    int eslot = 0;
    pdl_wait();  // residual reads / output writes touch buffers of preceding kernels
    const int oc_shift = g.oc == 64 ? 6 : (g.oc == 32 ? 5 : 4);
    if constexpr (EPI == 1) {
      // ---- lean epilogue (the common case: bf16 output through the staging buffer + TMA store, no ragged chunks).
      // The tile's accumulators are MB*BN/16 chunks of 16 TMEM columns (chunk f = block f / nch, columns 16*(f % nch)).
      // A work item is a pair of adjacent chunks (one 32-column tcgen05.ld); the two warps of a lane quadrant take the
      // even / odd items (single chunks when the tile has only one pair).  The epilogue warps are latency-bound (two per
      // scheduler): everything per thread that does not depend on the chunk is computed once per tile, the residual rows
      // of the next item are prefetched into L1 while the current one is processed, and the accumulator stage is released
      // with one mbarrier arrival per warp.
      const int nchq = g.BN >> 4, nflat_q = g.MB * nchq;
      const bool pair_mode = nflat_q > 2;
      const int n_items = pair_mode ? (nflat_q + 1) >> 1 : nflat_q;
      const uint32_t nch_mul = (65536u + (uint32_t)nchq - 1u) / (uint32_t)nchq;  // f / nch == (f * mul) >> 16 for f < 256
      const uint32_t oc_mask = (uint32_t)g.oc - 1u, sw_mask = (uint32_t)g.o_sw_mask, orb = (uint32_t)g.o_row_bytes;
      const uint32_t obox = (uint32_t)g.obox_bytes;
      const uint32_t row_off0 = rr0 * orb, blk_off = blk_out_rows * orb;
      const int act = g.act, res_mode = g.res_mode, THq = g.TH, BNq = g.BN;
      const size_t res_blk = (size_t)THq * g.Wo * g.res_cs;             // residual elements between consecutive blocks
      const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
      for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(g, tile);
        const int wo = tc.tw_i * g.TW + tw, n0 = tc.nt * BNq;
        const int ho0 = tc.th_i * rows_tile + th;
        const uint32_t taddr = tlane + (uint32_t)(as * nflat_q * 16);
        const uint32_t o_stage = o_base + (uint32_t)(os * g.o_stage_bytes);
        const bool col_ok = row_ok && wo < g.Wo;
        const float *bias_t = bias_s + n0;
        const size_t pix0 = ((size_t)tc.b * g.Ho + ho0) * g.Wo + wo;
        const bf16 *res0 = g.res + pix0 * g.res_cs + n0;

        auto chunk_jc = [&](int f, int &j, int &c) {
          j = (int)(((uint32_t)f * nch_mul) >> 16);
          c = f - j * nchq;
        };
        // residual rows (2 x 16 B) of chunk f; nullptr outside the image
        auto res_ptr = [&](int f) -> const uint4 * {
          int j, c;
          chunk_jc(f, j, c);
          if (!(col_ok && ho0 + j * THq < g.Ho)) return nullptr;
          if (res_mode == B2Y_RES_PREADD_UP2) {  // half-resolution map: pixel (b, y/2, x/2) of the real (unflattened) image
            const size_t pix = pix0 + (size_t)j * THq * g.Wo;
            const int rb = (int)(pix / (size_t)g.real_hw);
            const int a = (int)(pix - (size_t)rb * g.real_hw);
            const int y = a / g.real_w, x = a - y * g.real_w;
            const size_t hp = ((size_t)rb * (g.real_hw >> 2)) + (size_t)(y >> 1) * (g.real_w >> 1) + (x >> 1);
            return reinterpret_cast<const uint4 *>(g.res + hp * g.res_cs + n0 + c * 16);
          }
          return reinterpret_cast<const uint4 *>(res0 + (size_t)j * res_blk + c * 16);
        };
        auto res_load = [&](int f, uint4 *r) -> bool {
          const uint4 *rp = res_ptr(f);
          if (rp) {
            r[0] = __ldg(rp);
            r[1] = __ldg(rp + 1);
          }
          return rp != nullptr;
        };
        auto res_prefetch = [&](int f) {  // next item: pull its rows into L1 without holding registers
          const uint4 *rp = res_ptr(f);
          if (rp) asm volatile("prefetch.global.L1 [%0];" ::"l"(rp));
        };
        // item p -> first chunk and whether it has a second one
        auto item_f = [&](int p, bool &two) -> int {
          const int f0 = pair_mode ? 2 * p : p;
          two = pair_mode && (f0 + 1 < nflat_q);
          return f0;
        };

        // bias (+SiLU) (+residual) of one chunk, packed to bf16 into the swizzled staging tile
        auto finish = [&](int f, float(&v)[16], const uint4 *r, const bool r_ok) {
          int j, c;
          chunk_jc(f, j, c);
          const int col = c * 16;
          const float4 *bp = reinterpret_cast<const float4 *>(bias_t + col);
          if (res_mode == B2Y_RES_PREADD_UP2 && r_ok) {  // pre-activation: conv + upsampled half-resolution partial sum
            float rf[16];
            unpack8(r[0], rf);
            unpack8(r[1], rf + 8);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] += rf[q];
          }
          if (act) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 hb = bp[q];
              v[4 * q] = silu_half(fmaf(v[4 * q], 0.5f, hb.x));
              v[4 * q + 1] = silu_half(fmaf(v[4 * q + 1], 0.5f, hb.y));
              v[4 * q + 2] = silu_half(fmaf(v[4 * q + 2], 0.5f, hb.z));
              v[4 * q + 3] = silu_half(fmaf(v[4 * q + 3], 0.5f, hb.w));
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bb = bp[q];
              v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
            }
          }
          if (r_ok && res_mode != B2Y_RES_PREADD_UP2) {
            float rf[16];
            unpack8(r[0], rf);
            unpack8(r[1], rf + 8);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (res_mode == B2Y_RES_ADD) ? v[q] + rf[q] : v[q] * rf[q];
          }
          if (row_ok) {
            const uint32_t rowoff = row_off0 + (uint32_t)j * blk_off;
            const uint32_t xr = ((rowoff >> 7) & sw_mask) << 4;
            const uint32_t dst = o_stage + rowoff + (uint32_t)(col >> oc_shift) * obox;
            const uint32_t u = (((uint32_t)col & oc_mask) * 2u);  // byte offset of the chunk inside the staging row
            const uint4 p0 = pack8(v), p1 = pack8(v + 8);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (u ^ xr)), "r"(p0.x), "r"(p0.y), "r"(p0.z), "r"(p0.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((u + 16u) ^ xr)), "r"(p1.x), "r"(p1.y), "r"(p1.z), "r"(p1.w) : "memory");
          }
        };

        uint4 rc[4];  // residual rows of the current item: rc[0..1] first chunk, rc[2..3] second chunk
        bool rcA_ok = false, rcB_ok = false;
        const bool has_items = half < n_items;
        if (res_mode && has_items) {  // independent of the MMAs: in flight across the accumulator wait below
          bool two;
          const int f0 = item_f(half, two);
          rcA_ok = res_load(f0, rc);
          rcB_ok = two && res_load(f0 + 1, rc + 2);
        }

        mbar_wait(tfull0 + 8 * as, aphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elected) trace_evt(g, 2, eslot);  // [5i] accumulators ready

        bool reclaimed = false;
        for (int p = half; p < n_items; p += 2) {
          bool two;
          const int f0 = item_f(p, two);  // warp-uniform
          float va[16], vb[16];
          if (two) tmem_ld32_issue(taddr + (uint32_t)(f0 * 16), va, vb);
          else tmem_ld16_issue(taddr + (uint32_t)(f0 * 16), va);
          const bool last = p + 2 >= n_items;
          if (res_mode) {
            if (p != half) {  // (the first item's rows were requested before the accumulator wait)
              rcA_ok = res_load(f0, rc);
              rcB_ok = two && res_load(f0 + 1, rc + 2);
            }
            if (!last) {
              bool two_n;
              const int fn = item_f(p + 2, two_n);
              res_prefetch(fn);
              if (two_n) res_prefetch(fn + 1);
            }
          }
          tmem_ld_wait(va);
          if (two) tmem_ld_wait(vb);
          if (last) {  // this warp's last tcgen05.ld of the tile has landed: the MMA warp may overwrite the stage
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * as);
          }
          if (!reclaimed) {
            // the staging buffer about to be written was last read by the TMA store issued o_stages tiles ago
            if (warp == 2) {
              if (g.o_stages == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            named_bar_sync(1, kEpiThreads);
            reclaimed = true;
            if (elected) trace_evt(g, 2, eslot);  // [5i+1] staging buffer reclaimed (all warps arrived)
          }
          finish(f0, va, rc, res_mode && rcA_ok);
          if (two) finish(f0 + 1, vb, rc + 2, res_mode && rcB_ok);
        }
        if (!has_items) {  // a warp without items still takes part in the barrier protocol
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * as);
          if (warp == 2) {
            if (g.o_stages == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          named_bar_sync(1, kEpiThreads);
          if (elected) trace_evt(g, 2, eslot);
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
        if (elected) trace_evt(g, 2, eslot);  // [5i+2] this thread's math + staging writes done
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to TMA
        named_bar_sync(2, kEpiThreads);
        if (elected) trace_evt(g, 2, eslot);  // [5i+3] every warp's staging writes done
        if (warp == 2) {
          if (elect_one()) {
            for (int bx = 0; bx < g.n_oboxes; ++bx)
              if (n0 + bx * g.oc < g.cout)
                tma_store_4d(&tmO, o_stage + (uint32_t)bx * obox, n0 + bx * g.oc, tc.tw_i * g.TW, tc.th_i * rows_tile, tc.b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
          if (elected) trace_evt(g, 2, eslot);  // [5i+4] TMA store issued
        }
        if (++os == g.o_stages) os = 0;
      }
    } else if constexpr (EPI == 2) {
      // ---- lean fused Detect._inference (head.py:100-131) for the 1x1 output convs of a level: thread = anchor.
      // epi_mode 1 (box conv, 64 = 4 sides x 16 DFL bins): the warp pair of a lane quadrant splits the sides, half 0 takes
      // left/right (-> cx, w), half 1 top/bottom (-> cy, h): softmax-expectation (block.py:73-76), make_anchors +0.5
      // (tal.py:371-383), dist2bbox xywh (tal.py:386-395), * stride.  epi_mode 2 (class conv): sigmoid of nc columns,
      // contiguous column ranges per warp.  Results go straight to y (B, 4+nc, A) fp32: 32 lanes = 32 consecutive anchors.
      const int MBq = g.MB, BNq = g.BN, THq = g.TH, mode = g.epi_mode;
      const int c_lo = (mode == 1) ? half : (half ? (nch + 1) >> 1 : 0);
      const int c_hi = (mode == 1) ? 4 : (half ? nch : (nch + 1) >> 1);
      const int c_step = (mode == 1) ? 4 : 2;  // mode 1: one item = chunks (half, half + 2)
      const float kLog2e = 1.4426950408889634f;
      const size_t An = (size_t)g.A;
      for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(g, tile);
        const int wo = tc.tw_i * g.TW + tw, n0 = tc.nt * BNq;
        const int ho0 = tc.th_i * rows_tile + th;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * MBq * BNq);
        mbar_wait(tfull0 + 8 * as, aphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bool released = false;
        for (int j = 0; j < MBq; ++j) {
          const int ho = ho0 + j * THq;
          const bool valid = row_ok && ho < g.Ho && wo < g.Wo;
          // image / anchor coordinates (1x1 convs run over the flattened pixel axis)
          int rb = tc.b, rho = ho, rwo = wo;
          size_t apix = (size_t)ho * g.Wo + wo;
          if (g.flat) {
            rb = (int)(apix / (size_t)g.real_hw);
            apix -= (size_t)rb * g.real_hw;
            rho = (int)apix / g.real_w;
            rwo = (int)apix - rho * g.real_w;
          }
          float *yrow = reinterpret_cast<float *>(g.out) + (size_t)rb * (4 + g.nc) * An + g.a_off + apix;
          for (int c = c_lo; c < c_hi; c += c_step) {
            const bool two = (mode == 1) || (c + 1 < c_hi);
            const bool last = (j == MBq - 1) && (c + c_step >= c_hi);
            float va[16], vb[16];
            const uint32_t ta = taddr + (uint32_t)(j * BNq + c * 16);
            if (mode == 1) {
              tmem_ld16_issue(ta, va);
              tmem_ld16_issue(ta + 32u, vb);
            } else if (two) {
              tmem_ld32_issue(ta, va, vb);
            } else {
              tmem_ld16_issue(ta, va);
            }
            tmem_ld_wait(va);
            if (two) tmem_ld_wait(vb);
            if (last) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              mbar_arrive(tempty0 + 8 * as);
              released = true;
            }
            if (mode == 1) {
              auto dfl = [&](float(&v)[16], const int cc) -> float {
                const float4 *bp = reinterpret_cast<const float4 *>(bias_s + n0 + cc * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 bb = bp[q];
                  v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
                }
                float mx = v[0];
#pragma unroll
                for (int q = 1; q < 16; ++q) mx = fmaxf(mx, v[q]);
                float se = 0.f, sk = 0.f;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                  const float e = exp2f((v[q] - mx) * kLog2e);
                  se += e;
                  sk = fmaf(e, (float)q, sk);
                }
                return __fdividef(sk, se);
              };
              const float d1 = dfl(va, c), d2 = dfl(vb, c + 2);  // (left, right) or (top, bottom)
              if (valid) {
                const float anc = (half == 0 ? (float)rwo : (float)rho) + 0.5f;
                const float p1 = anc - d1, p2 = anc + d2;
                yrow[(size_t)half * An] = ((p1 + p2) * 0.5f) * g.det_stride;        // cx | cy
                yrow[(size_t)(2 + half) * An] = (p2 - p1) * g.det_stride;          // w  | h
              }
            } else {
              auto sig16 = [&](float(&v)[16], const int cc) {
                const int n = n0 + cc * 16;
                if (n >= g.cout) return;
                const float4 *bp = reinterpret_cast<const float4 *>(bias_s + n);
                float *yp = yrow + (size_t)(4 + n) * An;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 bb = bp[q];
                  v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
                }
                if (valid) {
#pragma unroll
                  for (int q = 0; q < 16; ++q)
                    if (n + q < g.cout) yp[(size_t)q * An] = __fdividef(1.f, 1.f + exp2f(-v[q] * kLog2e));
                }
              };
              sig16(va, c);
              if (two) sig16(vb, c + 1);
            }
          }
        }
        if (!released) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(tempty0 + 8 * as);
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    } else {
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(g, tile);
      const int wo = tc.tw_i * g.TW + tw, n0 = tc.nt * g.BN;
      const int ho0 = tc.th_i * rows_tile + th;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * g.MB * g.BN);
      const uint32_t o_stage = o_base + (uint32_t)(os * g.o_stage_bytes);

      // one 16-column chunk of accumulator block j: +bias, SiLU, then fused decode | residual + store
      auto chunk = [&](float(&v)[16], const int j, const int c) {
        const int col = c * 16, n = n0 + col;
        if (n >= g.cout) return;  // warp-uniform
        const int ho = ho0 + j * g.TH;
        const bool valid = row_ok && ho < g.Ho && wo < g.Wo;
        const size_t pix = ((size_t)tc.b * g.Ho + ho) * g.Wo + wo;
        const float4 *bp = reinterpret_cast<const float4 *>(bias_s + n);
        if (g.act) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 hb = bp[q];
            v[4 * q] = silu_half(fmaf(v[4 * q], 0.5f, hb.x));
            v[4 * q + 1] = silu_half(fmaf(v[4 * q + 1], 0.5f, hb.y));
            v[4 * q + 2] = silu_half(fmaf(v[4 * q + 2], 0.5f, hb.z));
            v[4 * q + 3] = silu_half(fmaf(v[4 * q + 3], 0.5f, hb.w));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bb = bp[q];
            v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
          }
        }
        if (g.epi_mode) {
          // ---- fused Detect._inference: thread = one anchor (pixel); chunk c of the box conv = one box side (16 DFL bins)
          // image / anchor coordinates (the conv itself may run over the flattened pixel axis)
          int rb = tc.b, rho = ho, rwo = wo;
          size_t apix = (size_t)ho * g.Wo + wo;
          if (g.flat) {
            rb = (int)(apix / (size_t)g.real_hw);
            apix -= (size_t)rb * g.real_hw;
            rho = (int)apix / g.real_w;
            rwo = (int)apix - rho * g.real_w;
          }
          if (g.epi_mode == 1) {
            float mx = v[0];
#pragma unroll
            for (int q = 1; q < 16; ++q) mx = fmaxf(mx, v[q]);
            float se = 0.f, sk = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float e = expf(v[q] - mx);
              se += e;
              sk = fmaf(e, (float)q, sk);
            }
            const float dist = sk / se;  // DFL expectation (block.py:73-76)
            if (c < 2) {
              dfl_first = dist;          // left (c = 0) or top (c = 1); the partner side is this warp's next chunk
            } else if (valid) {
              const float anc = (c == 2 ? (float)rwo : (float)rho) + 0.5f;  // make_anchors (tal.py:371-383)
              const float p1 = anc - dfl_first, p2 = anc + dist;            // dist2bbox (tal.py:386-395)
              float *yb = reinterpret_cast<float *>(g.out) + ((size_t)rb * (4 + g.nc) + (c - 2)) * g.A + g.a_off + apix;
              yb[0] = ((p1 + p2) * 0.5f) * g.det_stride;         // cx | cy
              yb[(size_t)2 * g.A] = (p2 - p1) * g.det_stride;   // w  | h
            }
          } else if (valid) {
            float *yb = reinterpret_cast<float *>(g.out) + ((size_t)rb * (4 + g.nc) + 4 + n) * g.A + g.a_off + apix;
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (n + q < g.cout) yb[(size_t)q * g.A] = 1.f / (1.f + expf(-v[q]));
          }
          return;
        }
        const bool full16 = (n + 16 <= g.cout);
        if (g.res_mode && valid) {
          if (g.vec_ok && full16) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(g.res + pix * g.res_cs + n);
            float r[16];
            unpack8(__ldg(rp), r);
            unpack8(__ldg(rp + 1), r + 8);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (g.res_mode == B2Y_RES_ADD) ? v[q] + r[q] : v[q] * r[q];
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q)  // static indices only: v[] must stay in registers
              if (n + q < g.cout) {
                const float r = __bfloat162float(g.res[pix * g.res_cs + n + q]);
                v[q] = (g.res_mode == B2Y_RES_ADD) ? v[q] + r : v[q] * r;
              }
          }
        }
        if (g.tma_out) {
          if (row_ok) {
            const int bx = col >> oc_shift, cb = (col - (bx << oc_shift)) * 2;  // channel box, byte offset of the chunk in the row
            uint32_t off = (rr0 + (uint32_t)j * blk_out_rows) * (uint32_t)g.o_row_bytes + (uint32_t)cb;
            const uint32_t off2 = off + 16;
            off ^= ((off >> 7) & (uint32_t)g.o_sw_mask) << 4;
            const uint32_t o2 = off2 ^ (((off2 >> 7) & (uint32_t)g.o_sw_mask) << 4);
            const uint32_t dst = o_stage + (uint32_t)(bx * g.obox_bytes);
            const uint4 p0 = pack8(v), p1 = pack8(v + 8);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + off), "r"(p0.x), "r"(p0.y), "r"(p0.z), "r"(p0.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + o2), "r"(p1.x), "r"(p1.y), "r"(p1.z), "r"(p1.w) : "memory");
          }
        } else if (valid) {
          if (g.vec_ok && full16) {
            if (g.out_f32) {
              float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(g.out) + pix * g.out_cs + n);
#pragma unroll
              for (int q = 0; q < 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
              uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(g.out) + pix * g.out_cs + n);
              op[0] = pack8(v);
              op[1] = pack8(v + 8);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (n + q < g.cout) {
                if (g.out_f32) reinterpret_cast<float *>(g.out)[pix * g.out_cs + n + q] = v[q];
                else reinterpret_cast<bf16 *>(g.out)[pix * g.out_cs + n + q] = __float2bfloat16_rn(v[q]);
              }
          }
        }
      };

      mbar_wait(tfull0 + 8 * as, aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // this warp's chunks f = half, half+2, ... (f = block j * nch + 16-column chunk c), two per round: both TMEM loads
      // are in flight while the staging buffer is being reclaimed; the accumulator stage is released to the MMA warp as
      // soon as the last load has landed in registers, before the math and the stores.
      int j = 0, c = half;
      while (c >= nch) { c -= nch; ++j; }
      bool reclaimed = false, released = false;
      for (int f = half; f < nflat; f += 4) {
        const int j0 = j, c0 = c;
        c += 2;
        while (c >= nch) { c -= nch; ++j; }
        const int j1 = j, c1 = c;
        c += 2;
        while (c >= nch) { c -= nch; ++j; }
        const bool two = f + 2 < nflat;  // warp-uniform
        float v0[16], v1[16];
        tmem_ld16_issue(taddr + (uint32_t)(j0 * g.BN + c0 * 16), v0);  // warp-collective: every lane, valid or not
        if (two) tmem_ld16_issue(taddr + (uint32_t)(j1 * g.BN + c1 * 16), v1);
        if (!reclaimed && g.tma_out) {
          // the staging buffer about to be written was last read by the TMA store issued o_stages tiles ago
          if (warp == 2) {
            if (g.o_stages == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          named_bar_sync(1, kEpiThreads);
        }
        reclaimed = true;
        tmem_ld_wait(v0);
        if (two) tmem_ld_wait(v1);
        if (f + 4 >= nflat) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(tempty0 + 8 * as);  // accumulator stage drained: the MMA warp may overwrite it
          released = true;
        }
        chunk(v0, j0, c0);
        if (two) chunk(v1, j1, c1);
      }
      if (!reclaimed && g.tma_out) {  // a warp without chunks still takes part in the barrier protocol
        if (warp == 2) {
          if (g.o_stages == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        named_bar_sync(1, kEpiThreads);
      }
      if (!released) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(tempty0 + 8 * as);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
      if (g.tma_out) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to TMA
        named_bar_sync(2, kEpiThreads);
        if (warp == 2) {
          if (elect_one()) {
            for (int bx = 0; bx < g.n_oboxes; ++bx)
              if (n0 + bx * g.oc < g.cout)
                tma_store_4d(&tmO, o_stage + (uint32_t)(bx * g.obox_bytes), n0 + bx * g.oc, tc.tw_i * g.TW,
                             tc.th_i * rows_tile, tc.b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
        }
        if (++os == g.o_stages) os = 0;
      }
    }
    }
    if (g.tma_out && warp == 2) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (elected) trace_evt(g, 2, eslot);  // last: all stores complete
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
void find_div(unsigned d, unsigned &mul, unsigned &shr) {  // n / d == umulhi(n, mul) >> shr for n < 2^31 (mul == 0: d == 1)
  if (d <= 1) { mul = 0; shr = 0; return; }
  unsigned lg = 0;
  while ((1ull << lg) < d) ++lg;
  const unsigned p = 31 + lg;
  mul = (unsigned)(((1ull << p) + d - 1) / d);
  shr = p - 32;
}

typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcGeom);

template <int TAPS, bool S2, int EPI>
KernelFn pick_ksteps(int ksteps) {
  return ksteps == 4 ? conv_tc_kernel<TAPS, 4, S2, EPI> : ksteps == 2 ? conv_tc_kernel<TAPS, 2, S2, EPI>
                                                                        : conv_tc_kernel<TAPS, 1, S2, EPI>;
}

template <int EPI>
KernelFn pick_taps(int taps, int ksteps, bool stride2) {
  if (taps == 1) return pick_ksteps<1, false, EPI>(ksteps);
  if (taps == 4) return pick_ksteps<4, false, EPI>(ksteps);
  return stride2 ? pick_ksteps<9, true, EPI>(ksteps) : pick_ksteps<9, false, EPI>(ksteps);
}

KernelFn pick_kernel(int taps, int ksteps, bool stride2, int epi) {
  if (epi == 2) return pick_ksteps<1, false, 2>(ksteps);  // the Detect output convs are 1x1
  return epi == 1 ? pick_taps<1>(taps, ksteps, stride2) : pick_taps<0>(taps, ksteps, stride2);
}

}  // namespace

struct ConvTcPlan {
  CUtensorMap tmA, tmB, tmO;
  TcGeom g;
  unsigned long long *trace = nullptr;  // device buffer when tracing (B2Y_TRACE_OP)
  int grid;
  size_t smem;
  KernelFn fn;
};

bool conv_tc_supported(const ConvParams &p, char *why, int why_len) {
  auto no = [&](const char *m) {
    if (why) snprintf(why, why_len, "%s", m);
    return false;
  };
  if (p.kh != p.kw) return no("non-square kernel");
  const bool k1 = p.kh == 1 && p.sh == 1 && p.sw == 1 && p.ph == 0 && p.pw == 0;
  const bool k3s1 = p.kh == 3 && p.sh == 1 && p.sw == 1 && p.ph == 1 && p.pw == 1 && !p.asym;
  const bool k3s2 = p.kh == 3 && p.sh == 2 && p.sw == 2 && p.ph == 1 && p.pw == 1 && p.H % 2 == 0 && p.W % 2 == 0;
  const bool k2 = p.kh == 2 && p.sh == 1 && p.sw == 1 && p.ph == 1 && p.pw == 1 && p.asym == 1;  // space-to-depth stem
  if (!(k1 || k3s1 || k3s2 || k2)) return no("kernel/stride/padding combination");
  if (p.epi_mode && !(k1 && p.out_f32 && p.res_mode == 0 && !p.act)) return no("fused decode needs a plain fp32 1x1 conv");
  if (p.epi_mode == 1 && p.cout != 64) return no("fused DFL decode needs 64 box channels");
  if (p.grouped && !(p.cin == p.cout && p.cout == p.cout_pad && p.cout <= 256 && !p.epi_mode && p.kh == 3))
    return no("grouped (depthwise) mode needs a 3x3 conv with cin == cout <= 256, multiple of 16");
  if (p.res_mode == B2Y_RES_PREADD_UP2 &&
      !(k1 && !p.out_f32 && !p.epi_mode && p.cout % 16 == 0 && p.res_cs % 8 == 0 && p.out_cs % 8 == 0 && p.Ho % 2 == 0 &&
        p.Wo % 2 == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0))
    return no("upsampled pre-activation residual needs an aligned bf16 1x1 conv");
  if (p.cin % 16) return no("cin not a multiple of 16");
  if (p.cout_pad % 16) return no("cout_pad not a multiple of 16");
  if (p.in_cs % 8 || (reinterpret_cast<uintptr_t>(p.in) & 15)) return no("input view not 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(p.w) & 15) return no("weights not 16-byte aligned");
  if (p.Wo > 65535 || p.Ho > 65535) return no("map too large");
  return true;
}

int conv_tc_plan_create(const ConvParams &p_in, int device, const TcConfig &cfg, ConvTcPlan **out) {
  char why[128];
  B2Y_REQUIRE(conv_tc_supported(p_in, why, sizeof(why)), B2Y_ERR_UNSUPPORTED, "tcgen05 conv: %s", why);
  EncodeTiledFn encode = get_encode_fn();
  B2Y_REQUIRE(encode != nullptr, B2Y_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cudaDeviceProp prop;
  B2Y_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  const int num_sms = prop.multiProcessorCount;
  ConvTcPlan *pl = new ConvTcPlan();
  TcGeom &g = pl->g;
  memset(&g, 0, sizeof(g));
  // 1x1 convs see the activations as one flat pixel axis (NHWC pixels are equidistant in memory across rows and
  // images): a tile is Wf consecutive pixels x MB rows of a (npix/Wf, Wf) grid, Wf the largest divisor of npix <= 128,
  // so accumulator blocks are full and every TMA box row is one contiguous run, whatever the map size.
  ConvParams pf = p_in;
  if (p_in.kh == 1) {
    const long long npix = (long long)p_in.B * p_in.Ho * p_in.Wo;
    int wf = 128;
    while (wf > 1 && npix % wf) --wf;
    g.real_hw = p_in.Ho * p_in.Wo;  // (also valid unflattened: pixel index = (b*Ho + ho)*Wo + wo either way)
    g.real_w = p_in.Wo;
    if (wf >= 64 && npix / wf <= 65535) {
      g.flat = 1;
      pf.B = 1;
      pf.H = pf.Ho = (int)(npix / wf);
      pf.W = pf.Wo = wf;
    }
  }
  const ConvParams &p = pf;
  g.B = p.B; g.Ho = p.Ho; g.Wo = p.Wo; g.cin = p.cin; g.cout = p.cout; g.cout_pad = p.cout_pad;
  g.taps = p.kh * p.kw; g.stride2 = (p.sh == 2); g.in_cs = p.in_cs;
  g.nplanes = g.stride2 ? 4 : 1;
  g.grouped = p.grouped ? 1 : 0;
  g.BK = g.grouped ? 16 : (p.cin % 64 == 0) ? 64 : (p.cin % 32 == 0) ? 32 : 16;
  g.nkc = p.cin / g.BK;
  g.row_bytes = g.BK * 2;
  const int halo = (p.kh == 3) ? (g.stride2 ? 1 : 2) : (p.kh == 2 ? 1 : 0);  // extra halo columns/rows per tile
  g.ox = g.oy = (p.kh >= 2) ? 1 : 0;
  g.n_ntiles = ceil_div(p.cout_pad, 256);
  g.BN = ceil_div(ceil_div(p.cout_pad, g.n_ntiles), 16) * 16;
  // bf16 outputs are staged in shared memory and leave through TMA stores (coalesced, clipped, channel-sliced by the
  // tensor map); fp32 head outputs and unaligned views are stored directly from registers.
  g.tma_out = (!p.out_f32 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && p.out_cs % 8 == 0 && p.cout % 8 == 0) ? 1 : 0;
  g.oc = g.BN <= 16 ? 16 : (g.BN <= 32 ? 32 : 64);
  g.n_oboxes = ceil_div(g.BN, g.oc);
  g.o_row_bytes = g.oc * 2;
  g.o_sw_mask = g.oc == 64 ? 7 : (g.oc == 32 ? 3 : 1);
  const int bias_bytes = round_up(p.cout_pad * 4, 1024);
  const int w_tiles = g.taps * g.nkc;
  const int b_rows = g.grouped ? 16 : g.BN;  // weight rows per (tap, chunk) tile
  g.b_tile_bytes = round_up(b_rows * g.row_bytes, g.grouped ? 512 : 1024);
  g.b_tx_bytes = b_rows * g.row_bytes;
  const long long w_res_bytes = round_up(w_tiles * g.b_tile_bytes, 1024);
  g.w_bytes_tx = w_tiles * g.b_tx_bytes;

  // spatial block: accumulator rows m = th*P + tw, P = TW + halo; maximise useful rows per 128-row MMA block
  {
    double best = -1.0;
    int bth = 1, btw = 1;
    for (int tw = 1; tw <= (p.Wo < 128 - halo ? p.Wo : 128 - halo); ++tw) {
      const int P = tw + halo;
      int th = (128 - tw) / P + 1;  // (th-1)*P + tw <= 128
      if (th > p.Ho) th = p.Ho;
      if (th < 1 || P > 256) continue;
      const double eff = (double)p.Ho * p.Wo / ((double)ceil_div(p.Ho, th) * ceil_div(p.Wo, tw) * 128.0);
      const double amp = (double)(th + halo) * P / ((double)th * tw);  // halo read amplification
      const double score = eff - 0.02 * amp + 1e-4 * tw;  // ties: the widest tile (longest contiguous TMA rows)
      if (score > best) { best = score; bth = th; btw = tw; }
    }
    g.TH = bth; g.TW = btw; g.P = btw + halo;
  }
  g.tiles_w = ceil_div(p.Wo, g.TW);

  // blocks per tile (MB), stages and residency under the shared-memory / TMEM / parallelism budgets.
  // cfg.per_sm CTAs are meant to co-reside on an SM: the shared-memory budget is divided accordingly.
  const int want_per_sm = cfg.per_sm > 0 ? cfg.per_sm : 1;
  const int smem_budget = want_per_sm == 1 ? 220 * 1024 : (want_per_sm == 2 ? 110 * 1024 : 72 * 1024);
  int best_mb = 0;
  for (int mb = (cfg.mb > 0 ? cfg.mb : kMaxMB); mb >= (cfg.mb > 0 ? cfg.mb : 1); --mb) {
    if (2 * mb * g.BN * want_per_sm > 512) continue;   // TMEM columns (double buffered, per co-resident CTA)
    if (mb * g.TH + halo > 256) continue;              // TMA box limit
    if (mb > 1 && (mb - 1) * g.TH >= p.Ho) continue;   // would only add empty blocks
    const long long tiles = (long long)p.B * ceil_div(p.Ho, mb * g.TH) * g.tiles_w * g.n_ntiles;
    if (cfg.mb <= 0 && mb > 1 && tiles < 4LL * num_sms * want_per_sm) continue;  // keep every SM busy with a few tiles
    const int box_rows = (mb * g.TH + halo) * g.P;
    const int max_shift = (p.kh == 1) ? 0 : ((g.stride2 || p.kh == 2) ? g.P + 1 : 2 * g.P + 2);
    const int need_rows = (mb - 1) * g.TH * g.P + max_shift + 128;
    // one A stage = one halo box (stride 2: one parity plane; a k-chunk of a tile is nsub consecutive stages)
    const int a_stage = round_up((box_rows > need_rows ? box_rows : need_rows) * g.row_bytes, 1024);
    const int nsub = g.stride2 ? 4 : 1;
    const int o_stage = g.tma_out ? g.n_oboxes * mb * 128 * g.o_row_bytes : 0;
    int o_stages = g.tma_out ? 2 : 0;
    const long long fixed = bias_bytes + 1024;
    if (g.grouped && !((g.n_ntiles == 1) && (w_res_bytes + (long long)(nsub + 1) * a_stage + o_stage + fixed <= smem_budget)))
      continue;  // grouped weights are tiny and must be resident
    const bool res_ok = (g.n_ntiles == 1) && (w_res_bytes + (long long)(nsub + 1) * a_stage + o_stage + fixed <= smem_budget);
    const long long b_bytes = res_ok ? w_res_bytes : 2LL * g.b_tile_bytes;
    if (b_bytes + (long long)nsub * a_stage + o_stage + fixed > smem_budget) continue;  // does not fit even single-buffered
    if (b_bytes + (long long)(nsub + 1) * a_stage + (long long)o_stages * o_stage + fixed > smem_budget)
      o_stages = g.tma_out ? 1 : 0;
    best_mb = mb;
    g.MB = mb;
    g.plane_bytes = a_stage;
    g.a_stage_bytes = a_stage;
    g.a_tx_bytes = box_rows * g.row_bytes;
    g.o_stage_bytes = o_stage;
    g.o_stages = o_stages;
    g.resident = res_ok ? 1 : 0;
    break;
  }
  if (!best_mb) {
    delete pl;
    b2y_set_error("tcgen05 conv: no tile configuration fits (BN %d, BK %d, mb %d, per_sm %d)", g.BN, g.BK, cfg.mb, cfg.per_sm);
    return B2Y_ERR_UNSUPPORTED;
  }
  g.obox_bytes = g.MB * 128 * g.o_row_bytes;
  g.blk_rows_bytes = g.TH * g.P * g.row_bytes;
  g.tiles_h = ceil_div(p.Ho, g.MB * g.TH);
  g.ntiles = (long long)p.B * g.tiles_h * g.tiles_w * g.n_ntiles;
  if (g.ntiles >= (1ll << 31)) {
    delete pl;
    b2y_set_error("tcgen05 conv: too many tiles");
    return B2Y_ERR_UNSUPPORTED;
  }
  // tap sequence: position -> (tap index, operand start inside its A stage = row shift, 16-byte units).  Stride 2: taps
  // grouped by parity plane pl = rowpar*2 + colpar (one stage each): {centre}, {left,right}, {up,down}, {corners}.
  {
    int pos = 0;
    for (int pl = 0; pl < (g.stride2 ? 4 : 1); ++pl)
      for (int r = 0; r < p.kh; ++r)
        for (int s = 0; s < p.kw; ++s) {
          int shift;
          if (g.stride2) {
            if ((((r == 1) ? 0 : 1) * 2 + ((s == 1) ? 0 : 1)) != pl) continue;
            shift = ((r == 0) ? 0 : 1) * g.P + ((s == 0) ? 0 : 1);
          } else {
            shift = (p.kh == 1) ? 0 : r * g.P + s;
          }
          g.tap_seq[pos] = r * p.kw + s;
          g.tap_off16[pos] = (shift * g.row_bytes) >> 4;
          ++pos;
        }
  }
  // stage counts: as many A stages as fit (<= 8): bytes in flight per SM are what hides the HBM latency on these layers
  const int out_bytes = g.o_stages * g.o_stage_bytes;
  const int nsub = g.stride2 ? 4 : 1;
  int b_bytes_total;
  if (g.resident) {
    g.sb = 0;
    b_bytes_total = round_up((int)w_res_bytes, 1024);  // keeps the output staging area 1024-byte aligned
  } else {
    g.sb = 4;
    while (g.sb > 2 && (long long)g.sb * g.b_tile_bytes + (long long)(nsub + 1) * g.a_stage_bytes + out_bytes + bias_bytes + 1024 > smem_budget) --g.sb;
    b_bytes_total = g.sb * g.b_tile_bytes;
  }
  int sa = (smem_budget - b_bytes_total - out_bytes - bias_bytes - 1024) / g.a_stage_bytes;
  if (sa > kMaxStages) sa = kMaxStages;
  {
    // a CTA never has more than its own tiles' stages in flight: thin layers (1-2 tiles per CTA) keep their shared-memory
    // request small, so the next kernel's CTAs can become resident beside them (PDL) instead of queueing for the SM
    const long long max_ctas_ = (long long)num_sms * want_per_sm;
    const long long tiles_per_cta = (g.ntiles + max_ctas_ - 1) / max_ctas_;
    const long long useful = tiles_per_cta * g.nkc * nsub;
    if (sa > useful) sa = (int)(useful < nsub ? nsub : useful);
  }
  if (sa < 1) sa = 1;
  g.sa = sa;
  g.o_off = g.sa * g.a_stage_bytes + b_bytes_total;
  g.bias_off = g.o_off + out_bytes;
  pl->smem = (size_t)g.bias_off + bias_bytes + 1024;
  find_div((unsigned)g.n_ntiles, g.dv_nt_mul, g.dv_nt_shr);
  find_div((unsigned)g.tiles_w, g.dv_tw_mul, g.dv_tw_shr);
  find_div((unsigned)g.tiles_h, g.dv_th_mul, g.dv_th_shr);
  int cols = 32;
  while (cols < 2 * g.MB * g.BN) cols <<= 1;
  g.tmem_cols = cols;
  const unsigned layout = (g.row_bytes == 128) ? 2u : (g.row_bytes == 64) ? 4u : 6u;  // UMMA LayoutType SW128/SW64/SW32
  const unsigned sbo = (unsigned)(8 * g.row_bytes) >> 4;
  g.desc_hi = sbo | (1u << 14) | (layout << 29);  // bits [32,46) SBO, [46,48) version=1, [61,64) layout
  g.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)((g.grouped ? 16 : g.BN) >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  g.out = p.out; g.out_cs = p.out_cs; g.out_f32 = p.out_f32;
  g.res = p.res; g.res_cs = p.res_cs; g.res_mode = p.res_mode;
  g.bias = p.bias; g.act = p.act;
  g.epi_mode = p.epi_mode; g.a_off = p.a_off; g.A = p.A; g.nc = p.nc; g.det_stride = p.det_stride;
  {
    const uintptr_t o = reinterpret_cast<uintptr_t>(p.out);
    bool ok = (o & 15) == 0 && (p.out_f32 ? (p.out_cs % 4 == 0) : (p.out_cs % 8 == 0));
    if (p.res_mode) ok = ok && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0 && (p.res_cs % 8 == 0);
    g.vec_ok = ok ? 1 : 0;
  }
  g.fast_epi = (g.tma_out && !g.epi_mode && p.cout % 16 == 0 && (!p.res_mode || g.vec_ok)) ? 1 : 0;

  const CUtensorMapSwizzle sw = (g.row_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (g.row_bytes == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_32B;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const int box_h = g.MB * g.TH + halo;
  CUresult r;
  if (!g.stride2) {
    const cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.in_cs * 2, (cuuint64_t)p.W * p.in_cs * 2,
                                   (cuuint64_t)p.H * p.W * p.in_cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)g.BK, (cuuint32_t)g.P, (cuuint32_t)box_h, 1};
    r = encode(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16 *>(p.in), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[5] = {(cuuint64_t)p.in_cs + p.cin, (cuuint64_t)p.W / 2, 2, (cuuint64_t)p.H / 2,
                                (cuuint64_t)p.B};
    const cuuint64_t strides[4] = {(cuuint64_t)2 * p.in_cs * 2, (cuuint64_t)p.W * p.in_cs * 2,
                                   (cuuint64_t)2 * p.W * p.in_cs * 2, (cuuint64_t)p.H * p.W * p.in_cs * 2};
    const cuuint32_t box[5] = {(cuuint32_t)g.BK, (cuuint32_t)g.P, 1, (cuuint32_t)box_h, 1};
    r = encode(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16 *>(p.in), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    delete pl;
    b2y_set_error("cuTensorMapEncodeTiled(activations) failed with CUresult %d", (int)r);
    return B2Y_ERR_CUDA;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)g.taps * p.cin, (cuuint64_t)p.cout_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)g.taps * p.cin * 2};
    const cuuint32_t box[2] = {(cuuint32_t)g.BK, (cuuint32_t)(g.grouped ? 16 : g.BN)};
    r = encode(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(p.w), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    delete pl;
    b2y_set_error("cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
    return B2Y_ERR_CUDA;
  }
  if (g.tma_out) {
    const CUtensorMapSwizzle osw = (g.oc == 64) ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (g.oc == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                  : CU_TENSOR_MAP_SWIZZLE_32B;
    const cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.out_cs * 2, (cuuint64_t)p.Wo * p.out_cs * 2,
                                   (cuuint64_t)p.Ho * p.Wo * p.out_cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)g.oc, (cuuint32_t)g.TW, (cuuint32_t)(g.MB * g.TH), 1};
    r = encode(&pl->tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p.out, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               osw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      b2y_set_error("cuTensorMapEncodeTiled(output) failed with CUresult %d", (int)r);
      return B2Y_ERR_CUDA;
    }
  } else {
    pl->tmO = pl->tmB;  // unused
  }
  // without PDL keep the SM to one CTA as configured; with PDL the request stays minimal so that the next kernel's CTA
  // can become resident beside this one (the persistent grid is one CTA per SM, placed breadth-first)
  if (want_per_sm == 1 && pl->smem < 116 * 1024 && !b2y_pdl_enabled()) pl->smem = 116 * 1024;
  if (g.tmem_cols * want_per_sm > 512) {
    delete pl;
    b2y_set_error("tcgen05 conv: TMEM over-subscribed (%d columns x %d CTAs)", g.tmem_cols, want_per_sm);
    return B2Y_ERR_UNSUPPORTED;
  }
  const long long max_ctas = (long long)num_sms * want_per_sm;
  pl->grid = (int)(g.ntiles < max_ctas ? g.ntiles : max_ctas);
  pl->fn = pick_kernel(g.taps, g.BK / 16, g.stride2 != 0, g.epi_mode ? 2 : (g.fast_epi ? 1 : 0));
  B2Y_CUDA_CHECK(cudaFuncSetAttribute(pl->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  *out = pl;
  return B2Y_OK;
}

void conv_tc_plan_destroy(ConvTcPlan *p) {
  if (p && p->trace) cudaFree(p->trace);
  delete p;
}

// Debug aid: log the pipeline timeline of CTA 0 (producer / MMA / epilogue events, ns relative to the first) to stderr.
int conv_tc_plan_trace(ConvTcPlan *pl, cudaStream_t st) {
  if (!pl->trace) B2Y_CUDA_CHECK(cudaMalloc(&pl->trace, 3 * 64 * sizeof(unsigned long long)));
  pl->g.trace = pl->trace;
  B2Y_CUDA_CHECK(cudaMemsetAsync(pl->trace, 0, 3 * 64 * sizeof(unsigned long long), st));
  int rc = launch_conv_tc(pl, st);
  if (rc != B2Y_OK) return rc;
  unsigned long long h[192];
  B2Y_CUDA_CHECK(cudaMemcpyAsync(h, pl->trace, sizeof(h), cudaMemcpyDeviceToHost, st));
  B2Y_CUDA_CHECK(cudaStreamSynchronize(st));
  unsigned long long t0 = ~0ull;
  for (int i = 0; i < 192; ++i)
    if (h[i] && h[i] < t0) t0 = h[i];
  const TcGeom &g = pl->g;
  fprintf(stderr, "[b2y trace] grid %d ntiles %lld MB %d TH %d TW %d BN %d BK %d nkc %d sa %d resident %d o_stages %d fast %d smem %zu\n",
          pl->grid, g.ntiles, g.MB, g.TH, g.TW, g.BN, g.BK, g.nkc, g.sa, g.resident, g.o_stages, g.fast_epi, pl->smem);
  const char *names[3] = {"producer", "mma", "epilogue"};
  for (int r = 0; r < 3; ++r) {
    fprintf(stderr, "[b2y trace] %-8s", names[r]);
    for (int i = 0; i < 64 && h[r * 64 + i]; ++i) fprintf(stderr, " %llu", h[r * 64 + i] - t0);
    fprintf(stderr, "\n");
  }
  pl->g.trace = nullptr;
  return B2Y_OK;
}

int launch_conv_tc(const ConvTcPlan *pl, cudaStream_t st) {
  b2y_launch(pl->fn, dim3(pl->grid), dim3(kNumThreads), pl->smem, st, pl->tmA, pl->tmB, pl->tmO, pl->g);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

// Bit-exact non_max_suppression (ultralytics/utils/ops.py:167-316) + torchvision.ops.nms (third party; greedy
// IoU suppression over a stable descending score order, fp32 IoU by division, strict '>').
//
//   kernel 1  nms_score_kernel   grid (A/256, B): per anchor, the class maximum / first arg-max (ops.py:274, single
//                                label) or every class above conf (ops.py:271, multi label), conf + class filter
//                                -> score per slot (slot = anchor | anchor*nc + class), -1 if not a candidate
//   kernel 2  nms_select_kernel  one 1024-thread CTA per image: order-preserving compaction of the candidate slots,
//                                64-bit bitonic sort of (~score_bits, slot) (descending score, ties by ascending
//                                anchor index == the stable sort torchvision performs on anchor-ordered rows),
//                                max_nms cut (ops.py:285-286), then a blocked greedy scan: 64 candidates at a time are
//                                tested against the kept boxes (shared memory) and against each other through a
//                                64x64 IoU bitmask built with warp ballots; stops at max_det (ops.py:297).
// Every fp32 operation mirrors the reference's op order (xywh2xyxy ops.py:416-433, class offset boxes + cls*max_wh
// ops.py:289-295); this file is compiled with -fmad=false so nothing is contracted into FMAs.
#include <stdlib.h>

#include "common.cuh"

namespace {
constexpr int NT = 1024;
constexpr int CH = 64;            // candidates per greedy step
constexpr int MAX_KEPT = 1024;    // upper bound on max_det
constexpr int SMEM_SORT = 8192;   // keys sorted in shared memory (64 KB); larger sets sort in global memory
constexpr int PRESEL_K = 6144;    // images with more than SMEM_SORT candidates: the top PRESEL_K (+ score ties) are tried first
constexpr size_t SMEM_DYN = (size_t)SMEM_SORT * 8;

struct ClassMask {
  uint32_t bits[32];  // nc <= 1024
  int enabled;
};

struct NmsParams {
  const float *pred;  // B x (4+nc) x A
  int B, nc, A, S;    // S slots per image
  float conf, iou_f, max_wh;
  int max_det, max_nms, agnostic, multi_label;
  int preselect;      // 1: top-PRESEL_K preselection for images with more than SMEM_SORT candidates (validator mode)
  float *slot_score;  // B x S
  int *slot_cls;      // B x A (single label)
  unsigned long long *keys;  // B x P2 (P2 = pow2 >= S)
  int P2;
  float *dets;
  int *counts, *kept_anchor;
  ClassMask cm;
};

__device__ __forceinline__ bool class_ok(const ClassMask &cm, int c) {
  return !cm.enabled || ((cm.bits[c >> 5] >> (c & 31)) & 1u);
}

__global__ void __launch_bounds__(256) nms_score_kernel(NmsParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= p.A) return;
  const float *cls = p.pred + ((size_t)b * (4 + p.nc) + 4) * p.A + a;
  if (!p.multi_label) {
    float best = __ldg(cls);
    int bc = 0, c = 1;
    for (; c + 8 <= p.nc; c += 8) {  // eight independent (coalesced, A-strided) loads in flight per thread
      float s[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = __ldg(cls + (size_t)(c + k) * p.A);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (s[k] > best) { best = s[k]; bc = c + k; }  // strict >: the first index wins ties (ops.py:269, probed)
    }
    for (; c < p.nc; ++c) {
      const float s = __ldg(cls + (size_t)c * p.A);
      if (s > best) { best = s; bc = c; }
    }
    const bool ok = (best > p.conf) && class_ok(p.cm, bc);
    p.slot_score[(size_t)b * p.S + a] = ok ? best : -1.f;
    p.slot_cls[(size_t)b * p.A + a] = bc;
  } else {
    float *out = p.slot_score + (size_t)b * p.S + (size_t)a * p.nc;
    for (int c = 0; c < p.nc; ++c) {
      const float s = cls[(size_t)c * p.A];
      out[c] = (s > p.conf && class_ok(p.cm, c)) ? s : -1.f;
    }
  }
}

struct Box { float x1, y1, x2, y2; };

__device__ __forceinline__ bool iou_gt(const Box &k, float ka, const Box &c, float ca, float thr) {
  const float xx1 = fmaxf(k.x1, c.x1), yy1 = fmaxf(k.y1, c.y1);
  const float xx2 = fminf(k.x2, c.x2), yy2 = fminf(k.y2, c.y2);
  const float w = __fsub_rn(xx2, xx1), h = __fsub_rn(yy2, yy1);
  // disjoint boxes (the vast majority of pairs): torchvision clamps w, h to 0, so inter = 0 and IoU is 0 or NaN (0/0),
  // neither of which exceeds a threshold >= 0 -- same decision without the division
  if (!(w > 0.f && h > 0.f)) return false;
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ka, ca), inter));
  return ovr > thr;  // NaN (0/0 for zero-area duplicates) compares false: not suppressed, like the reference
}

__device__ void bitonic_sort(unsigned long long *k, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += NT) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool up = ((i & size) == 0);
        const unsigned long long a = k[i], c = k[j];
        if ((a > c) == up) { k[i] = c; k[j] = a; }
      }
    }
  }
  __syncthreads();
}

// Exclusive block scan of one count per thread: returns this thread's offset, `total` = the sum.  `scan` (NT/32 ints)
// and `s_total` are shared scratch, reusable when the call returns.
__device__ __forceinline__ int block_scan_excl(int cnt, int *scan, int *s_total, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int v = scan[lane];
    int iv = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += u;
    }
    scan[lane] = iv - v;  // exclusive warp offsets
    if (lane == 31) *s_total = iv;
  }
  __syncthreads();
  total = *s_total;
  const int off = scan[warp] + incl - cnt;
  __syncthreads();
  return off;
}

// The K-th smallest 32-bit score key (~score bits: ascending key = descending score) among the candidate slots
// [s0, s1) of all threads: three radix passes (12 + 12 + 8 bits) with a shared-memory histogram (4096 ints).  Exact.
// Requires at least K candidates in total.
__device__ unsigned kth_score_key(const float *score, int s0, int s1, int K, int *hist, int *s_sel) {
  unsigned prefix = 0, mask = 0;
  const int shifts[3] = {20, 8, 0}, bits[3] = {12, 12, 8};
  for (int r = 0; r < 3; ++r) {
    const int nb = 1 << bits[r], sh = shifts[r];
    for (int i = threadIdx.x; i < nb; i += NT) hist[i] = 0;
    __syncthreads();
    for (int s = s0; s < s1; ++s) {
      const float v = score[s];
      if (v >= 0.f) {
        const unsigned key = ~__float_as_uint(v);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> sh) & (nb - 1)], 1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, d = 0;
      for (; d < nb - 1; ++d) {
        if (cum + hist[d] >= K) break;
        cum += hist[d];
      }
      s_sel[0] = d;
      s_sel[1] = K - cum;
    }
    __syncthreads();
    prefix |= (unsigned)s_sel[0] << sh;
    mask |= (unsigned)(nb - 1) << sh;
    K = s_sel[1];
    __syncthreads();
  }
  return prefix;
}

__global__ void __launch_bounds__(NT) nms_select_kernel(NmsParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ unsigned long long skeys[];  // SMEM_SORT keys
  __shared__ int scan[NT / 32];
  __shared__ int s_total;
  __shared__ Box kbox[MAX_KEPT];
  __shared__ float karea[MAX_KEPT];
  __shared__ Box cbox[CH], craw[CH];
  __shared__ float carea[CH], cscore[CH], ccls[CH];
  __shared__ int canchor[CH];
  __shared__ int csup[CH];
  __shared__ unsigned long long cmask[CH];
  __shared__ int s_newkept[CH], s_nnew, s_nk, s_sel[2];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *score = p.slot_score + (size_t)b * p.S;
  unsigned long long *gkeys = p.keys + (size_t)b * p.P2;

  // ---- 1. order-preserving compaction: thread t owns slots [t*per, (t+1)*per)
  const int per = (p.S + NT - 1) / NT;
  const int s0 = min(tid * per, p.S), s1 = min(s0 + per, p.S);
  int cnt = 0;
  for (int s = s0; s < s1; ++s) cnt += (score[s] >= 0.f);
  int n;
  const int off_all = block_scan_excl(cnt, scan, &s_total, n);
  if (n == 0) {
    if (tid == 0) p.counts[b] = 0;
    for (int i = tid; i < p.max_det * 6; i += NT) p.dets[(size_t)b * p.max_det * 6 + i] = 0.f;  // padded rows are defined
    if (p.kept_anchor)
      for (int i = tid; i < p.max_det; i += NT) p.kept_anchor[(size_t)b * p.max_det + i] = 0;
    return;
  }
  // Images with more candidates than the shared-memory sort holds (validator mode: conf 0.001, multi_label, up to A * nc
  // slots): attempt 0 sorts only the PRESEL_K best candidates (+ everything tied with the last of them), found by an exact
  // radix select.  The greedy scan visits candidates in sorted order and never looks ahead, so whenever it reaches max_det
  // (or max_nms) inside that prefix the result IS the full result; if it runs out of candidates first, attempt 1 redoes
  // the image with every candidate.
  const bool can_preselect = p.preselect && n > SMEM_SORT;
  const float *pb = p.pred + (size_t)b * (4 + p.nc) * p.A;
  for (int attempt = can_preselect ? 0 : 1; attempt < 2; ++attempt) {
    unsigned T = 0xffffffffu;  // candidates with key32 <= T take part
    int n1 = n, pos = off_all;
    if (attempt == 0) {
      T = kth_score_key(score, s0, s1, PRESEL_K, reinterpret_cast<int *>(skeys), s_sel);
      int c1 = 0;
      for (int s = s0; s < s1; ++s) {
        const float v = score[s];
        c1 += (v >= 0.f && ~__float_as_uint(v) <= T);
      }
      pos = block_scan_excl(c1, scan, &s_total, n1);
    }
    int P = 1;
    while (P < n1) P <<= 1;
    unsigned long long *keys = (P <= SMEM_SORT) ? skeys : gkeys;
    for (int s = s0; s < s1; ++s) {
      const float v = score[s];
      if (v >= 0.f && ~__float_as_uint(v) <= T) keys[pos++] = ((unsigned long long)(~__float_as_uint(v)) << 32) | (unsigned)s;
    }
    __syncthreads();
    for (int i = n1 + tid; i < P; i += NT) keys[i] = ~0ull;
    // ---- 2. sort: descending score, ascending slot
    bitonic_sort(keys, P);
    const int m = min(n1, p.max_nms);

    if (tid == 0) s_nk = 0;
    __syncthreads();
    // ---- 3. blocked greedy suppression
    for (int c0 = 0; c0 < m; c0 += CH) {
      const int nc_chunk = min(CH, m - c0);
      const int nk = s_nk;
      if (nk >= p.max_det) break;
      if (tid < CH) {
        csup[tid] = 0;
        if (tid < nc_chunk) {
          const unsigned long long key = keys[c0 + tid];
          const int slot = (int)(key & 0xffffffffu);
          const float sc = __uint_as_float(~(unsigned)(key >> 32));
          int a, c;
          if (p.multi_label) { a = slot / p.nc; c = slot - a * p.nc; }
          else { a = slot; c = p.slot_cls[(size_t)b * p.A + a]; }
          const float cx = pb[a], cy = pb[(size_t)p.A + a], w = pb[(size_t)2 * p.A + a], h = pb[(size_t)3 * p.A + a];
          const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);
          Box r;
          r.x1 = __fsub_rn(cx, hw); r.y1 = __fsub_rn(cy, hh); r.x2 = __fadd_rn(cx, hw); r.y2 = __fadd_rn(cy, hh);
          const float off = __fmul_rn((float)c, p.agnostic ? 0.f : p.max_wh);
          Box o;
          o.x1 = __fadd_rn(r.x1, off); o.y1 = __fadd_rn(r.y1, off); o.x2 = __fadd_rn(r.x2, off); o.y2 = __fadd_rn(r.y2, off);
          craw[tid] = r; cbox[tid] = o;
          carea[tid] = __fmul_rn(__fsub_rn(o.x2, o.x1), __fsub_rn(o.y2, o.y1));
          cscore[tid] = sc; ccls[tid] = (float)c; canchor[tid] = a;
        }
      }
      __syncthreads();
      // (a) against boxes kept in earlier chunks
      for (int t = tid; t < nc_chunk * nk; t += NT) {
        const int i = t % nc_chunk, k = t / nc_chunk;
        if (iou_gt(kbox[k], karea[k], cbox[i], carea[i], p.iou_f)) csup[i] = 1;
      }
      // (b) pairwise mask inside the chunk: warp handles rows {warp, warp+32}, lanes are columns
      for (int i = warp; i < CH; i += 32) {
        unsigned lo = 0, hi = 0;
        const bool vi = i < nc_chunk;
        const int j0 = lane, j1 = lane + 32;
        const bool t0 = vi && j0 > i && j0 < nc_chunk && iou_gt(cbox[i], carea[i], cbox[j0], carea[j0], p.iou_f);
        const bool t1 = vi && j1 > i && j1 < nc_chunk && iou_gt(cbox[i], carea[i], cbox[j1], carea[j1], p.iou_f);
        lo = __ballot_sync(0xffffffffu, t0);
        hi = __ballot_sync(0xffffffffu, t1);
        if (lane == 0) cmask[i] = ((unsigned long long)hi << 32) | lo;
      }
      __syncthreads();
      // (c) serial resolve of the chunk
      if (tid == 0) {
        unsigned long long dead = 0;
        for (int i = 0; i < nc_chunk; ++i)
          if (csup[i]) dead |= (1ull << i);
        int nn = 0, k = nk;
        for (int i = 0; i < nc_chunk && k < p.max_det; ++i) {
          if ((dead >> i) & 1ull) continue;
          s_newkept[nn++] = i;
          ++k;
          dead |= cmask[i];
        }
        s_nnew = nn;
        s_nk = k;
      }
      __syncthreads();
      // (d) publish the newly kept boxes
      if (tid < s_nnew) {
        const int i = s_newkept[tid], k = nk + tid;
        kbox[k] = cbox[i];
        karea[k] = carea[i];
        float *d = p.dets + ((size_t)b * p.max_det + k) * 6;
        d[0] = craw[i].x1; d[1] = craw[i].y1; d[2] = craw[i].x2; d[3] = craw[i].y2; d[4] = cscore[i]; d[5] = ccls[i];
        if (p.kept_anchor) p.kept_anchor[(size_t)b * p.max_det + k] = canchor[i];
      }
      __syncthreads();
    }
    const bool exhausted = attempt == 0 && s_nk < p.max_det && n1 < min(n, p.max_nms);
    __syncthreads();  // every thread has read s_nk before a second attempt resets it
    if (!exhausted) break;
  }
  const int nk_final = s_nk;
  if (tid == 0) p.counts[b] = nk_final;
  for (int i = nk_final * 6 + tid; i < p.max_det * 6; i += NT) p.dets[(size_t)b * p.max_det * 6 + i] = 0.f;
  if (p.kept_anchor)
    for (int i = nk_final + tid; i < p.max_det; i += NT) p.kept_anchor[(size_t)b * p.max_det + i] = 0;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
inline int pow2_ge(long long v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

__global__ void scale_boxes_kernel(float *dets, const int *counts, const float *geom, int B, int max_det) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * max_det) return;
  const int b = idx / max_det, k = idx % max_det;
  if (k >= counts[b]) return;
  const float gain = geom[b * 5 + 0], padx = geom[b * 5 + 1], pady = geom[b * 5 + 2], w0 = geom[b * 5 + 3],
              h0 = geom[b * 5 + 4];
  float *d = dets + (size_t)idx * 6;
  d[0] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[0], padx), gain), 0.f), w0);
  d[1] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[1], pady), gain), 0.f), h0);
  d[2] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[2], padx), gain), 0.f), w0);
  d[3] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[3], pady), gain), 0.f), h0);
}
}  // namespace

extern "C" size_t b2y_nms_workspace_bytes(int32_t B, int32_t nc, int32_t A, int32_t multi_label, int32_t max_nms) {
  (void)max_nms;
  const long long S = multi_label ? (long long)A * nc : A;
  return align256((size_t)B * S * sizeof(float)) + align256((size_t)B * A * sizeof(int)) +
         align256((size_t)B * pow2_ge(S) * sizeof(unsigned long long));
}

extern "C" int b2y_nms(const void *pred, int32_t B, int32_t nc, int32_t A, float conf_thres, double iou_thres,
                       int32_t max_det, int32_t max_nms, float max_wh, int32_t agnostic, int32_t multi_label,
                       const int32_t *classes, int32_t n_classes, void *workspace, size_t workspace_bytes, void *dets,
                       void *counts, void *kept_anchor, void *stream) {
  B2Y_REQUIRE(pred && dets && counts && workspace && B > 0 && nc > 0 && A > 0, B2Y_ERR_INVALID, "nms: bad arguments");
  B2YDeviceGuard guard(b2y_device_of(pred));
  B2Y_REQUIRE(conf_thres >= 0.f && conf_thres <= 1.f, B2Y_ERR_INVALID,
              "Invalid Confidence threshold %g, valid values are between 0.0 and 1.0", (double)conf_thres);
  B2Y_REQUIRE(iou_thres >= 0.0 && iou_thres <= 1.0, B2Y_ERR_INVALID,
              "Invalid IoU %g, valid values are between 0.0 and 1.0", iou_thres);
  B2Y_REQUIRE(max_det > 0 && max_det <= MAX_KEPT, B2Y_ERR_UNSUPPORTED, "nms: max_det must be in [1, %d]", MAX_KEPT);
  B2Y_REQUIRE(nc <= 1024, B2Y_ERR_UNSUPPORTED, "nms: nc > 1024 not supported");
  B2Y_REQUIRE(workspace_bytes >= b2y_nms_workspace_bytes(B, nc, A, multi_label, max_nms), B2Y_ERR_NOMEM,
              "nms: workspace too small");
  NmsParams p;
  p.pred = reinterpret_cast<const float *>(pred);
  p.B = B; p.nc = nc; p.A = A;
  p.multi_label = (multi_label && nc > 1) ? 1 : 0;  // ops.py:239
  const long long S = p.multi_label ? (long long)A * nc : A;
  B2Y_REQUIRE(S < (1ll << 31), B2Y_ERR_UNSUPPORTED, "nms: too many slots");
  p.S = (int)S;
  p.conf = conf_thres;
  // (double)ovr > iou_thres  <=>  ovr > largest float <= iou_thres   (torchvision's CPU kernel compares in double)
  float f = (float)iou_thres;
  if ((double)f > iou_thres) f = nextafterf(f, -INFINITY);
  p.iou_f = f;
  p.max_wh = max_wh;
  p.max_det = max_det; p.max_nms = max_nms > 0 ? max_nms : 30000; p.agnostic = agnostic;
  static const int preselect = [] { const char *e = getenv("B2Y_NMS_PRESELECT"); return e ? atoi(e) : 1; }();  // A/B knob
  p.preselect = preselect;
  char *ws = reinterpret_cast<char *>(workspace);
  p.slot_score = reinterpret_cast<float *>(ws);
  ws += align256((size_t)B * (multi_label ? (size_t)A * nc : (size_t)A) * sizeof(float));
  p.slot_cls = reinterpret_cast<int *>(ws);
  ws += align256((size_t)B * A * sizeof(int));
  p.keys = reinterpret_cast<unsigned long long *>(ws);
  p.P2 = pow2_ge(multi_label ? (long long)A * nc : A);
  p.dets = reinterpret_cast<float *>(dets);
  p.counts = reinterpret_cast<int *>(counts);
  p.kept_anchor = reinterpret_cast<int *>(kept_anchor);
  p.cm.enabled = 0;
  for (int i = 0; i < 32; ++i) p.cm.bits[i] = 0;
  if (classes != nullptr) {
    p.cm.enabled = 1;
    for (int i = 0; i < n_classes; ++i)
      if (classes[i] >= 0 && classes[i] < nc) p.cm.bits[classes[i] >> 5] |= (1u << (classes[i] & 31));
  }
  cudaStream_t st = (cudaStream_t)stream;
  b2y_launch(nms_score_kernel, dim3(dim3(ceil_div(A, 256), B)), dim3(256), 0, st, p);
  B2Y_LAUNCH_CHECK();
  const size_t smem = SMEM_DYN;
  // per device and cheap: set on every call (a once-per-process flag breaks the second GPU of a multi-device process)
  B2Y_CUDA_CHECK(cudaFuncSetAttribute(nms_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  b2y_launch(nms_select_kernel, dim3(B), dim3(NT), smem, st, p);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

extern "C" int b2y_scale_boxes(void *dets, const void *counts, const void *geom, int32_t B, int32_t max_det,
                               void *stream) {
  B2Y_REQUIRE(dets && counts && geom && B > 0 && max_det > 0, B2Y_ERR_INVALID, "scale_boxes: bad arguments");
  B2YDeviceGuard guard(b2y_device_of(dets));
  b2y_launch(scale_boxes_kernel, dim3(ceil_div(B * max_det, 256)), dim3(256), 0, (cudaStream_t)stream, 
      reinterpret_cast<float *>(dets), reinterpret_cast<const int *>(counts), reinterpret_cast<const float *>(geom), B,
      max_det);
  B2Y_LAUNCH_CHECK();
  return B2Y_OK;
}

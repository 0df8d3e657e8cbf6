// Plan object + C ABI glue: resolves the planner's op list (include/b2y.h) against caller-owned device memory,
// builds TMA descriptors for the tcgen05 convolutions once, and replays the whole network either eagerly or as a
// CUDA graph.  Replaces the reference's Python layer interpreter (ultralytics/nn/tasks.py:138-165).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

void b2y_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool b2y_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("B2Y_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

extern "C" const char *b2y_last_error(void) { return g_err; }
extern "C" int b2y_abi_version(void) { return B2Y_ABI_VERSION; }

int decode_launch(const void *const *raw, const int32_t *hs, const int32_t *ws, const float *strides, int32_t n_levels,
                  int32_t B, int32_t nc, void *y, cudaStream_t st);

struct ResolvedOp {
  b2y_op op;
  ConvParams cp;
  ConvTcPlan *tc = nullptr;
  ConvPairPlan *pair = nullptr;
  int impl = 0;
};

struct b2y_plan {
  std::vector<ResolvedOp> ops;
  std::vector<b2y_buf> bufs;
  char *ws = nullptr;
  size_t ws_bytes = 0;
  const char *wts = nullptr;
  size_t wts_bytes = 0;
  int B = 0, device = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  std::vector<cudaStream_t> lane_streams;  // capture-time only (multi-lane graph build)
  std::vector<cudaEvent_t> lane_events;
  // "the wide layers are done" marker: recorded in front of op `mid_op` (graph: external event-record node), so that a
  // caller can park unrelated kernels (pre-/post-processing of neighbouring batches) under the thin, latency-bound layers
  cudaEvent_t mid_event = nullptr;
  int mid_op = -1;
  bool mid_recorded = false;
};

namespace {
inline size_t elem_size(int dtype) { return dtype == B2Y_F32 ? 4 : 2; }

struct Resolver {
  b2y_plan *pl;
  bool ok = true;
  const b2y_buf *buf(int i) {
    if (i < 0 || i >= (int)pl->bufs.size()) {
      b2y_set_error("op references buffer %d of %d", i, (int)pl->bufs.size());
      ok = false;
      return nullptr;
    }
    return &pl->bufs[i];
  }
  void *ptr(int i, int coff) {
    const b2y_buf *b = buf(i);
    if (!b) return nullptr;
    return pl->ws + b->offset + (size_t)coff * elem_size(b->dtype);
  }
  const void *wptr(int64_t off) { return off < 0 ? nullptr : pl->wts + off; }
};

int resolve_conv(b2y_plan *pl, ResolvedOp &r, int idx) {
  Resolver R{pl};
  const b2y_op &o = r.op;
  const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
  if (!R.ok) return B2Y_ERR_INVALID;
  ConvParams &c = r.cp;
  c.in = reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0]));
  c.in_cs = bi->C;
  c.out = R.ptr(o.out, o.out_coff);
  c.out_cs = bo->C;
  c.out_f32 = (bo->dtype == B2Y_F32);
  c.res = nullptr; c.res_cs = 0; c.res_mode = o.res_mode;
  if (o.res_mode) {
    const b2y_buf *br = R.buf(o.res);
    if (!R.ok) return B2Y_ERR_INVALID;
    const int rs = (o.res_mode == B2Y_RES_PREADD_UP2) ? 2 : 1;  // half-resolution residual, upsampled on the fly
    B2Y_REQUIRE(br->dtype == B2Y_BF16 && br->H * rs == bo->H && br->W * rs == bo->W && o.res_coff + o.cout <= br->C,
                B2Y_ERR_INVALID, "op %d: residual view does not match the output", idx);
    c.res = reinterpret_cast<const bf16 *>(R.ptr(o.res, o.res_coff));
    c.res_cs = br->C;
  }
  c.w = R.wptr(o.w_off);
  c.bias = reinterpret_cast<const float *>(R.wptr(o.b_off));
  c.B = pl->B; c.H = bi->H; c.W = bi->W; c.Ho = bo->H; c.Wo = bo->W;
  c.cin = o.cin; c.cout = o.cout; c.cout_pad = o.cout_pad;
  c.kh = o.kh; c.kw = o.kw; c.sh = o.sh; c.sw = o.sw; c.ph = o.ph; c.pw = o.pw; c.act = o.act;
  c.asym = 0; c.grouped = 0; c.epi_mode = 0; c.a_off = 0; c.A = 0; c.nc = 0; c.det_stride = 0.f;
  c.dw_chunk = 0; c.dw_in_stride = 0;
  if (o.kind == B2Y_OP_DWCONV && o.ip[0] > 0) {
    c.dw_chunk = o.ip[0];
    c.dw_in_stride = o.ip[1];
    B2Y_REQUIRE(o.cout % c.dw_chunk == 0 && c.dw_in_stride >= c.dw_chunk &&
                    o.in_coff[0] + (o.cout / c.dw_chunk - 1) * c.dw_in_stride + c.dw_chunk <= bi->C,
                B2Y_ERR_INVALID, "op %d: chunked depthwise input view out of range", idx);
  }
  if (o.kind == B2Y_OP_CONV) {
    c.asym = o.ip[0];
    c.grouped = o.ip[5];
    c.epi_mode = o.ip[1];
    c.a_off = o.ip[2];
    c.A = o.ip[3];
    c.nc = o.ip[4];
    c.det_stride = o.fp[0];
  }
  if (c.epi_mode) {  // output operand is y (B, 4+nc, A) fp32; the conv's own map is the input map (1x1)
    B2Y_REQUIRE(bo->dtype == B2Y_F32 && bo->C == c.A && bo->H == 4 + c.nc && c.a_off + bi->H * bi->W <= c.A,
                B2Y_ERR_INVALID, "op %d: fused decode output must be y (B,4+nc,A) fp32", idx);
    c.Ho = bi->H; c.Wo = bi->W;
    c.out = R.ptr(o.out, 0);
    c.out_cs = 0;
    B2Y_REQUIRE(bi->dtype == B2Y_BF16 && c.w && c.bias && o.in_coff[0] + o.cin <= bi->C, B2Y_ERR_INVALID,
                "op %d: fused decode conv operands", idx);
    return B2Y_OK;
  }
  B2Y_REQUIRE(bi->dtype == B2Y_BF16, B2Y_ERR_INVALID, "op %d: conv input must be bf16", idx);
  B2Y_REQUIRE(c.w && c.bias, B2Y_ERR_INVALID, "op %d: conv needs weights and bias", idx);
  B2Y_REQUIRE((c.dw_chunk || o.in_coff[0] + o.cin <= bi->C) && o.out_coff + o.cout <= bo->C, B2Y_ERR_INVALID,
              "op %d: channel view out of range", idx);
  const int padf = c.asym ? 1 : 2;
  B2Y_REQUIRE(o.sh > 0 && o.sw > 0 && c.Ho == (c.H + padf * c.ph - c.kh) / c.sh + 1 &&
                  c.Wo == (c.W + padf * c.pw - c.kw) / c.sw + 1,
              B2Y_ERR_INVALID, "op %d: output map %dx%d inconsistent with input %dx%d k%dx%d s%d p%d", idx, c.Ho, c.Wo,
              c.H, c.W, c.kh, c.kw, c.sh, c.ph);
  return B2Y_OK;
}

int resolve_pair(b2y_plan *pl, ResolvedOp &r, int idx, PairParams &q) {
  Resolver R{pl};
  const b2y_op &o = r.op;
  const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
  if (!R.ok) return B2Y_ERR_INVALID;
  B2Y_REQUIRE(bi->dtype == B2Y_BF16 && bo->dtype == B2Y_BF16 && bi->H == bo->H && bi->W == bo->W &&
                  o.in_coff[0] + o.cin <= bi->C && o.out_coff + o.cout <= bo->C,
              B2Y_ERR_INVALID, "op %d: fused Bottleneck views", idx);
  q.in = reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0]));
  q.in_cs = bi->C;
  q.out = R.ptr(o.out, o.out_coff);
  q.out_cs = bo->C;
  q.w1 = R.wptr(o.w_off);
  q.b1 = reinterpret_cast<const float *>(R.wptr(o.b_off));
  q.w2 = R.wptr(o.p_off[0]);
  q.b2 = reinterpret_cast<const float *>(R.wptr(o.p_off[1]));
  B2Y_REQUIRE(q.w1 && q.b1 && q.w2 && q.b2, B2Y_ERR_INVALID, "op %d: fused Bottleneck needs both weight sets", idx);
  q.B = pl->B; q.H = bi->H; q.W = bi->W;
  q.cin = o.cin; q.chid = o.ip[0]; q.cout = o.cout; q.add = o.ip[1];
  return B2Y_OK;
}

int run_op(b2y_plan *pl, ResolvedOp &r, int idx, cudaStream_t st) {
  if (idx == pl->mid_op && pl->mid_event) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    B2Y_CUDA_CHECK(cudaEventRecordWithFlags(pl->mid_event, st,
                                            cs == cudaStreamCaptureStatusNone ? cudaEventRecordDefault : cudaEventRecordExternal));
    if (cs == cudaStreamCaptureStatusNone) pl->mid_recorded = true;
  }
  Resolver R{pl};
  const b2y_op &o = r.op;
  switch (o.kind) {
    case B2Y_OP_CONV_PAIR:
      return launch_conv_pair(r.pair, st);
    case B2Y_OP_CONV: {
      if (r.impl == B2Y_IMPL_TCGEN05) {
        static const char *tr = getenv("B2Y_TRACE_OP");  // debug: comma-free single op index, eager launches only
        if (tr && atoi(tr) == idx) {
          cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
          cudaStreamIsCapturing(st, &cs);
          if (cs == cudaStreamCaptureStatusNone) return conv_tc_plan_trace(r.tc, st);
        }
        return launch_conv_tc(r.tc, st);
      }
      return launch_conv_simt(r.cp, st);
    }
    case B2Y_OP_DWCONV:
      return launch_dwconv(r.cp, st);
    case B2Y_OP_SPPF_POOL: {
      const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
      if (!R.ok) return B2Y_ERR_INVALID;
      return launch_sppf_pool(reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])), bi->C,
                              reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)), bo->C, o.cin, pl->B, bi->H, bi->W, st);
    }
    case B2Y_OP_RESAMPLE: {
      const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
      if (!R.ok) return B2Y_ERR_INVALID;
      B2Y_REQUIRE(bo->H == bi->H * o.ip[0] && bo->W == bi->W * o.ip[0], B2Y_ERR_INVALID, "op %d: resample dims", idx);
      return launch_resample(reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])), bi->C,
                             reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)), bo->C, o.cin, pl->B, bo->H, bo->W,
                             o.ip[0], st);
    }
    case B2Y_OP_ATTN: {
      const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
      if (!R.ok) return B2Y_ERR_INVALID;
      return launch_attention(reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])), bi->C,
                              reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)), bo->C, pl->B, bi->H * bi->W, o.ip[0],
                              o.ip[1], o.ip[2], st);
    }
    case B2Y_OP_ROWCOL_MEAN: {
      const b2y_buf *bi = R.buf(o.in[0]);
      if (!R.ok) return B2Y_ERR_INVALID;
      return launch_rowcol_mean(reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])), bi->C, o.cin, pl->B,
                                bi->H, bi->W, o.ip[0], reinterpret_cast<float *>(R.ptr(o.out, 0)),
                                reinterpret_cast<float *>(R.ptr(o.in[1], 0)), st);
    }
    case B2Y_OP_ELA_GATE:
      return launch_ela_gate(reinterpret_cast<const float *>(R.ptr(o.in[0], 0)),
                             reinterpret_cast<const float *>(R.ptr(o.in[1], 0)),
                             reinterpret_cast<const float *>(R.wptr(o.p_off[0])), o.cin, pl->B, o.ip[2], o.ip[3],
                             o.ip[0], o.ip[1], o.fp[0], o.ip[4], reinterpret_cast<float *>(R.ptr(o.out, 0)),
                             reinterpret_cast<float *>(R.ptr(o.in[2], 0)), reinterpret_cast<float *>(R.ptr(o.in[3], 0)),
                             st);
    case B2Y_OP_ELA_APPLY: {
      const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
      if (!R.ok) return B2Y_ERR_INVALID;
      return launch_ela_apply(reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])), bi->C,
                              reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)), bo->C, o.cin, pl->B, bi->H, bi->W,
                              reinterpret_cast<const float *>(R.ptr(o.in[1], 0)),
                              reinterpret_cast<const float *>(R.ptr(o.in[2], 0)),
                              reinterpret_cast<const float *>(R.ptr(o.in[3], 0)), o.fp[0], o.fp[1], o.fp[2], st);
    }
    case B2Y_OP_GAP4:
    case B2Y_OP_WSUM4: {
      const bf16 *br[4];
      int cs[4];
      int HW = 0;
      for (int j = 0; j < 4; ++j) {
        const b2y_buf *b = R.buf(o.in[j]);
        if (!R.ok) return B2Y_ERR_INVALID;
        br[j] = reinterpret_cast<const bf16 *>(R.ptr(o.in[j], o.in_coff[j]));
        cs[j] = b->C;
        HW = b->H * b->W;
      }
      if (o.kind == B2Y_OP_GAP4)
        return launch_gap4(br, cs, o.cin, pl->B, HW, reinterpret_cast<float *>(R.ptr(o.out, 0)), st);
      const b2y_buf *bo = R.buf(o.out);
      if (!R.ok) return B2Y_ERR_INVALID;
      return launch_wsum4(br, cs, reinterpret_cast<const float *>(R.ptr(o.res, 0)),
                          reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)), bo->C, o.cin, pl->B, HW, st);
    }
    case B2Y_OP_MSCA_GATE:
      return launch_msca_gate(reinterpret_cast<const float *>(R.ptr(o.in[0], 0)),
                              reinterpret_cast<const float *>(R.wptr(o.p_off[0])),
                              reinterpret_cast<const float *>(R.wptr(o.p_off[1])), o.cin, pl->B,
                              reinterpret_cast<float *>(R.ptr(o.out, 0)), st);
    case B2Y_OP_PMSFA_DW: {
      const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
      if (!R.ok) return B2Y_ERR_INVALID;
      B2Y_REQUIRE(bi->dtype == B2Y_BF16 && bo->dtype == B2Y_BF16 && bi->H == bo->H && bi->W == bo->W && o.cin == o.cout &&
                      o.in_coff[0] + o.cin <= bi->C && o.out_coff + o.cout <= bo->C && o.in[0] != o.out,
                  B2Y_ERR_INVALID, "op %d: fused PMSFA depthwise views", idx);
      PmsfaParams q;
      q.in = reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])); q.in_cs = bi->C;
      q.out = reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)); q.out_cs = bo->C;
      q.w2 = reinterpret_cast<const float *>(R.wptr(o.w_off)); q.b2 = reinterpret_cast<const float *>(R.wptr(o.b_off));
      q.w3 = reinterpret_cast<const float *>(R.wptr(o.p_off[0])); q.b3 = reinterpret_cast<const float *>(R.wptr(o.p_off[1]));
      B2Y_REQUIRE(q.w2 && q.b2 && q.w3 && q.b3, B2Y_ERR_INVALID, "op %d: fused PMSFA depthwise needs both weight sets", idx);
      q.B = pl->B; q.H = bi->H; q.W = bi->W; q.c = o.cin;
      return launch_pmsfa_dw(q, st);
    }
    case B2Y_OP_MSCA_BRANCH: {
      const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out), *bg = R.buf(o.res);
      if (!R.ok) return B2Y_ERR_INVALID;
      B2Y_REQUIRE(bi->dtype == B2Y_BF16 && bo->dtype == B2Y_BF16 && bg->dtype == B2Y_F32 && bi->H == bo->H && bi->W == bo->W &&
                      o.cout == 4 * o.cin && o.in_coff[0] + o.cin <= bi->C && o.out_coff + o.cout <= bo->C &&
                      bg->C * bg->H * bg->W >= 4 * o.cin,
                  B2Y_ERR_INVALID, "op %d: fused MSCA branch views", idx);
      MscaParams q;
      q.in = reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])); q.in_cs = bi->C;
      q.out = reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)); q.out_cs = bo->C;
      q.gap = reinterpret_cast<float *>(R.ptr(o.res, 0));
      q.params = reinterpret_cast<const float *>(R.wptr(o.w_off));
      B2Y_REQUIRE(q.params, B2Y_ERR_INVALID, "op %d: fused MSCA branch needs its packed parameters", idx);
      q.B = pl->B; q.H = bi->H; q.W = bi->W; q.C = o.cin;
      return launch_msca_branch(q, st);
    }
    case B2Y_OP_D2S: {
      const b2y_buf *bi = R.buf(o.in[0]), *bo = R.buf(o.out);
      if (!R.ok) return B2Y_ERR_INVALID;
      B2Y_REQUIRE(bi->dtype == B2Y_BF16 && bo->dtype == B2Y_BF16 && bo->H == 2 * bi->H && bo->W == 2 * bi->W &&
                      o.cin == 4 * o.cout && o.in_coff[0] + o.cin <= bi->C && o.out_coff + o.cout <= bo->C,
                  B2Y_ERR_INVALID, "op %d: depth-to-space views", idx);
      return launch_d2s(reinterpret_cast<const bf16 *>(R.ptr(o.in[0], o.in_coff[0])), bi->C,
                        reinterpret_cast<bf16 *>(R.ptr(o.out, o.out_coff)), bo->C, o.cout, pl->B, bi->H, bi->W, st);
    }
    case B2Y_OP_DECODE: {
      const void *raw[4];
      int32_t hs[4], wsz[4];
      float strides[4];
      const int nl = o.ip[1];
      B2Y_REQUIRE(nl >= 1 && nl <= 4, B2Y_ERR_INVALID, "op %d: decode levels", idx);
      for (int l = 0; l < nl; ++l) {
        const b2y_buf *b = R.buf(o.in[l]);
        if (!R.ok) return B2Y_ERR_INVALID;
        B2Y_REQUIRE(b->dtype == B2Y_F32 && b->C == 64 + o.ip[0], B2Y_ERR_INVALID, "op %d: head map %d layout", idx, l);
        raw[l] = R.ptr(o.in[l], 0);
        hs[l] = b->H;
        wsz[l] = b->W;
        strides[l] = o.fp[l];
      }
      return decode_launch(raw, hs, wsz, strides, nl, pl->B, o.ip[0], R.ptr(o.out, 0), st);
    }
    default:
      b2y_set_error("op %d: unknown kind %d", idx, o.kind);
      return B2Y_ERR_UNSUPPORTED;
  }
}
}  // namespace

extern "C" int b2y_plan_create(const b2y_op *ops, int32_t n_ops, const b2y_buf *bufs, int32_t n_bufs, void *workspace,
                               size_t workspace_bytes, const void *weights, size_t weights_bytes, int32_t batch,
                               int32_t device, b2y_plan **out) {
  B2Y_REQUIRE(ops && bufs && workspace && out && n_ops > 0 && n_bufs > 0 && batch > 0, B2Y_ERR_INVALID,
              "plan_create: bad arguments");
  B2YDeviceGuard guard(device);  // the caller's current device is restored on return
  b2y_plan *pl = new b2y_plan();
  pl->ws = reinterpret_cast<char *>(workspace);
  pl->ws_bytes = workspace_bytes;
  pl->wts = reinterpret_cast<const char *>(weights);
  pl->wts_bytes = weights_bytes;
  pl->B = batch;
  pl->device = device;
  pl->bufs.assign(bufs, bufs + n_bufs);
  for (int i = 0; i < n_bufs; ++i) {
    const b2y_buf &b = bufs[i];
    const size_t bytes = (size_t)batch * b.H * b.W * b.C * elem_size(b.dtype);
    if (b.offset < 0 || (b.offset & 255) || (size_t)b.offset + bytes > workspace_bytes) {
      b2y_set_error("buffer %d (offset %lld, %zu bytes) does not fit the %zu-byte workspace", i, (long long)b.offset,
                    bytes, workspace_bytes);
      delete pl;
      return B2Y_ERR_NOMEM;
    }
  }
  const char *force = getenv("B2Y_FORCE_SIMT");
  const bool force_simt = force && force[0] == '1';
  pl->ops.resize(n_ops);
  int rc = B2Y_OK;
  for (int i = 0; i < n_ops && rc == B2Y_OK; ++i) {
    ResolvedOp &r = pl->ops[i];
    r.op = ops[i];
    const int k = r.op.kind;
    if (k == B2Y_OP_CONV_PAIR) {
      PairParams q;
      rc = resolve_pair(pl, r, i, q);
      if (rc != B2Y_OK) break;
      rc = conv_pair_plan_create(q, device, &r.pair);
      r.impl = B2Y_IMPL_TCGEN05;
    } else if (k == B2Y_OP_CONV || k == B2Y_OP_DWCONV) {
      rc = resolve_conv(pl, r, i);
      if (rc != B2Y_OK) break;
      if (k == B2Y_OP_CONV) {
        char why[128] = "";
        const bool can_tc = conv_tc_supported(r.cp, why, sizeof(why));
        int impl = r.op.impl;
        if (impl == B2Y_IMPL_AUTO) impl = (can_tc && !force_simt) ? B2Y_IMPL_TCGEN05 : B2Y_IMPL_SIMT;
        if (impl == B2Y_IMPL_TCGEN05) {
          if (!can_tc) {
            b2y_set_error("op %d: tcgen05 conv requested but unsupported: %s", i, why);
            rc = B2Y_ERR_UNSUPPORTED;
            break;
          }
          rc = conv_tc_plan_create(r.cp, device, TcConfig{0, 0}, &r.tc);
        }
        if (impl != B2Y_IMPL_TCGEN05 && r.cp.epi_mode) {
          b2y_set_error("op %d: fused Detect decode exists only in the tcgen05 conv epilogue (%s); lower the plan with "
                        "fuse_decode=False", i, can_tc ? "SIMT forced" : why);
          rc = B2Y_ERR_UNSUPPORTED;
          break;
        }
        r.impl = impl;
      }
    }
  }
  if (rc != B2Y_OK) {
    b2y_plan_destroy(pl);
    return rc;
  }
  *out = pl;
  return B2Y_OK;
}

extern "C" int b2y_plan_destroy(b2y_plan *pl) {
  if (!pl) return B2Y_OK;
  B2YDeviceGuard guard(pl->device);
  for (auto &r : pl->ops) {
    if (r.tc) conv_tc_plan_destroy(r.tc);
    if (r.pair) conv_pair_plan_destroy(r.pair);
  }
  if (pl->graph_exec) cudaGraphExecDestroy(pl->graph_exec);
  if (pl->graph) cudaGraphDestroy(pl->graph);
  if (pl->mid_event) cudaEventDestroy(pl->mid_event);
  delete pl;
  return B2Y_OK;
}

extern "C" int b2y_plan_run(b2y_plan *pl, int32_t first, int32_t last, void *stream) {
  B2Y_REQUIRE(pl, B2Y_ERR_INVALID, "plan_run: null plan");
  B2YDeviceGuard guard(pl->device);
  const int n = (int)pl->ops.size();
  if (last < 0 || last > n) last = n;
  B2Y_REQUIRE(first >= 0 && first <= last, B2Y_ERR_INVALID, "plan_run: bad op range [%d,%d)", first, last);
  for (int i = first; i < last; ++i) {
    const int rc = run_op(pl, pl->ops[i], i, (cudaStream_t)stream);
    if (rc != B2Y_OK) return rc;
  }
  return B2Y_OK;
}

namespace {
// ---- dependency analysis for multi-lane capture ---------------------------------------------------------------------
// The op list is a DAG (Detect branches of the three levels, C3k cv1/cv2, the neck running beside the P3 head ...).
// Captured on one stream it becomes a chain; here every op is put on one of a few capture streams ("lanes") with event
// edges for its true dependencies, so independent thin layers overlap and fill the SMs the persistent kernels of a
// single thin layer leave idle.  Regions are (buffer, channel range); ops other than convs are treated conservatively
// (every buffer they touch is read and written).
struct Region {
  int buf, c0, c1;
  bool write;
};

void op_regions(const b2y_plan *pl, const ResolvedOp &r, std::vector<Region> &out) {
  const b2y_op &o = r.op;
  out.clear();
  auto whole = [&](int b, bool w) {
    if (b >= 0 && b < (int)pl->bufs.size()) out.push_back(Region{b, 0, 1 << 30, w});
  };
  if (o.kind == B2Y_OP_CONV || o.kind == B2Y_OP_DWCONV) {
    const int span = (o.kind == B2Y_OP_DWCONV && o.ip[0] > 0) ? (o.cout / o.ip[0] - 1) * o.ip[1] + o.ip[0] : o.cin;
    out.push_back(Region{o.in[0], o.in_coff[0], o.in_coff[0] + span, false});
    if (o.res_mode) out.push_back(Region{o.res, o.res_coff, o.res_coff + o.cout, false});
    if (o.kind == B2Y_OP_CONV && o.ip[1]) {
      // fused Detect decode: the six output convs write disjoint rows / anchor ranges of y; nothing in the plan reads y
      out.push_back(Region{o.out, o.ip[2] * 8 + o.ip[1], o.ip[2] * 8 + o.ip[1] + 1, true});
    } else {
      out.push_back(Region{o.out, o.out_coff, o.out_coff + o.cout, true});
    }
    return;
  }
  if (o.kind == B2Y_OP_CONV_PAIR) {
    out.push_back(Region{o.in[0], o.in_coff[0], o.in_coff[0] + o.cin, false});
    out.push_back(Region{o.out, o.out_coff, o.out_coff + o.cout, true});
    return;
  }
  if (o.kind == B2Y_OP_PMSFA_DW) {
    out.push_back(Region{o.in[0], o.in_coff[0], o.in_coff[0] + o.cin, false});
    out.push_back(Region{o.out, o.out_coff, o.out_coff + o.cout, true});
    return;
  }
  if (o.kind == B2Y_OP_MSCA_BRANCH) {
    out.push_back(Region{o.in[0], o.in_coff[0], o.in_coff[0] + o.cin, false});
    out.push_back(Region{o.out, o.out_coff, o.out_coff + o.cout, true});
    whole(o.res, true);
    return;
  }
  if (o.kind == B2Y_OP_SPPF_POOL) {
    out.push_back(Region{o.in[0], o.in_coff[0], o.in_coff[0] + o.cin, false});
    out.push_back(Region{o.out, o.out_coff, o.out_coff + o.cout, true});
    return;
  }
  if (o.kind == B2Y_OP_RESAMPLE || o.kind == B2Y_OP_ATTN || o.kind == B2Y_OP_D2S) {
    out.push_back(Region{o.in[0], o.in_coff[0], o.in_coff[0] + o.cin, false});
    out.push_back(Region{o.out, o.out_coff, o.out_coff + o.cout, true});
    return;
  }
  for (int j = 0; j < 4; ++j) whole(o.in[j], true);
  whole(o.res, true);
  whole(o.out, true);
}

inline bool overlaps(const Region &a, const Region &b) { return a.buf == b.buf && a.c0 < b.c1 && b.c0 < a.c1; }

int capture_lanes(b2y_plan *pl, cudaStream_t origin, int max_lanes) {
  const int n = (int)pl->ops.size();
  std::vector<std::vector<Region>> reg(n);
  for (int i = 0; i < n; ++i) op_regions(pl, pl->ops[i], reg[i]);
  std::vector<int> lane(n, 0), lane_last;          // lane of each op; last op issued on each lane
  std::vector<cudaStream_t> streams{origin};
  std::vector<cudaEvent_t> ev(n, nullptr);
  cudaEvent_t fork = nullptr;
  int rc = B2Y_OK;
  auto fail = [&](cudaError_t e, const char *what) {
    b2y_set_error("%s failed during multi-lane capture: %s", what, cudaGetErrorString(e));
    rc = B2Y_ERR_CUDA;
  };
  cudaError_t e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
  if (e != cudaSuccess) { fail(e, "cudaEventCreate"); return rc; }
  e = cudaEventRecord(fork, origin);
  if (e != cudaSuccess) fail(e, "cudaEventRecord");
  lane_last.push_back(-1);
  std::vector<int> preds;
  for (int j = 0; j < n && rc == B2Y_OK; ++j) {
    preds.clear();
    for (int i = j - 1; i >= 0; --i) {  // RAW, WAR and WAW against every earlier op
      bool dep = false;
      for (const Region &a : reg[j]) {
        for (const Region &b : reg[i])
          if ((a.write || b.write) && overlaps(a, b)) { dep = true; break; }
        if (dep) break;
      }
      if (dep) preds.push_back(i);
    }
    // lane: continue the chain of the latest predecessor that is still the tail of its lane, else open a lane
    int L = -1;
    for (int p : preds)
      if (lane_last[lane[p]] == p) { L = lane[p]; break; }  // preds are in descending order: the latest one wins
    if (L < 0) {
      if ((int)streams.size() < max_lanes) {
        cudaStream_t s2 = nullptr;
        e = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
        if (e != cudaSuccess) { fail(e, "cudaStreamCreate"); break; }
        e = cudaStreamWaitEvent(s2, fork, 0);  // joins the capture
        if (e != cudaSuccess) { fail(e, "cudaStreamWaitEvent"); break; }
        streams.push_back(s2);
        lane_last.push_back(-1);
        L = (int)streams.size() - 1;
      } else {
        L = preds.empty() ? 0 : lane[preds[0]];
      }
    }
    for (int p : preds) {
      if (lane[p] == L) continue;  // stream order covers it
      e = cudaStreamWaitEvent(streams[L], ev[p], 0);
      if (e != cudaSuccess) { fail(e, "cudaStreamWaitEvent"); break; }
    }
    if (rc != B2Y_OK) break;
    if (j == pl->mid_op && pl->mid_event) {
      // the mid-point marker promises "every earlier op is done": join the tails of all other lanes before it is recorded
      for (size_t L2 = 0; L2 < streams.size(); ++L2)
        if ((int)L2 != L && lane_last[L2] >= 0 && ev[lane_last[L2]]) cudaStreamWaitEvent(streams[L], ev[lane_last[L2]], 0);
    }
    rc = run_op(pl, pl->ops[j], j, streams[L]);
    if (rc != B2Y_OK) break;
    lane[j] = L;
    lane_last[L] = j;
    e = cudaEventCreateWithFlags(&ev[j], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev[j], streams[L]);
    if (e != cudaSuccess) fail(e, "cudaEventRecord");
  }
  // join every lane back into the origin stream
  for (size_t L = 1; L < streams.size(); ++L)
    if (lane_last[L] >= 0 && ev[lane_last[L]]) cudaStreamWaitEvent(origin, ev[lane_last[L]], 0);
  {
    const char *vb = getenv("B2Y_VERBOSE");
    if (vb && vb[0] == '1') {
      std::vector<int> cnt(streams.size(), 0);
      for (int j = 0; j < n; ++j) ++cnt[lane[j]];
      fprintf(stderr, "[b2y graph] %d ops on %zu lanes:", n, streams.size());
      for (int c : cnt) fprintf(stderr, " %d", c);
      fprintf(stderr, "\n");
    }
  }
  pl->lane_streams.assign(streams.begin() + 1, streams.end());
  pl->lane_events.assign(ev.begin(), ev.end());
  pl->lane_events.push_back(fork);
  return rc;
}
}  // namespace

extern "C" int b2y_plan_set_mid_op(b2y_plan *pl, int32_t op_index) {
  B2Y_REQUIRE(pl, B2Y_ERR_INVALID, "set_mid_op: null plan");
  B2YDeviceGuard guard(pl->device);
  B2Y_REQUIRE(op_index < (int)pl->ops.size(), B2Y_ERR_INVALID, "set_mid_op: op %d of %d", op_index, (int)pl->ops.size());
  B2Y_REQUIRE(!pl->graph_exec, B2Y_ERR_INVALID, "set_mid_op: call before b2y_plan_graph_build");
  if (op_index >= 0 && !pl->mid_event) B2Y_CUDA_CHECK(cudaEventCreateWithFlags(&pl->mid_event, cudaEventDisableTiming));
  pl->mid_op = op_index < 0 ? -1 : op_index;
  pl->mid_recorded = false;
  return B2Y_OK;
}

extern "C" int b2y_plan_wait_mid(b2y_plan *pl, void *stream) {
  B2Y_REQUIRE(pl, B2Y_ERR_INVALID, "wait_mid: null plan");
  B2YDeviceGuard guard(pl->device);
  if (pl->mid_event && pl->mid_op >= 0 && pl->mid_recorded)
    B2Y_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, pl->mid_event, 0));
  return B2Y_OK;
}

extern "C" int b2y_plan_graph_build(b2y_plan *pl, void *stream) {
  B2Y_REQUIRE(pl, B2Y_ERR_INVALID, "graph_build: null plan");
  B2YDeviceGuard guard(pl->device);
  (void)stream;  // capture happens on a private stream: the caller's may be the legacy default stream, which cannot capture
  if (pl->graph_exec) { cudaGraphExecDestroy(pl->graph_exec); pl->graph_exec = nullptr; }
  if (pl->graph) { cudaGraphDestroy(pl->graph); pl->graph = nullptr; }
  cudaStream_t st = nullptr;
  B2Y_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    cudaStreamDestroy(st);
    b2y_set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(e));
    return B2Y_ERR_CUDA;
  }
  int lanes = 4;
  if (const char *le = getenv("B2Y_LANES")) lanes = atoi(le) > 0 ? atoi(le) : 1;
  const int rc = lanes > 1 ? capture_lanes(pl, st, lanes) : b2y_plan_run(pl, 0, -1, st);
  cudaGraph_t g = nullptr;
  e = cudaStreamEndCapture(st, &g);
  for (cudaStream_t ls : pl->lane_streams) cudaStreamDestroy(ls);
  for (cudaEvent_t le : pl->lane_events)
    if (le) cudaEventDestroy(le);
  pl->lane_streams.clear();
  pl->lane_events.clear();
  cudaStreamDestroy(st);
  if (rc != B2Y_OK) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  if (e != cudaSuccess) {
    b2y_set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    return B2Y_ERR_CUDA;
  }
  pl->graph = g;
  B2Y_CUDA_CHECK(cudaGraphInstantiate(&pl->graph_exec, pl->graph, 0));
  {
    const char *vb = getenv("B2Y_VERBOSE");
    if (vb && vb[0] == '1') {  // how many kernel->kernel edges were captured as programmatic (PDL) dependencies
      size_t ne = 0;
      if (cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &ne) == cudaSuccess && ne > 0) {
        std::vector<cudaGraphNode_t> from(ne), to(ne);
        std::vector<cudaGraphEdgeData> ed(ne);
        size_t prog = 0;
        if (cudaGraphGetEdges_v2(g, from.data(), to.data(), ed.data(), &ne) == cudaSuccess)
          for (size_t i = 0; i < ne; ++i) prog += (ed[i].type == cudaGraphDependencyTypeProgrammatic);
        fprintf(stderr, "[b2y graph] %zu edges, %zu programmatic\n", ne, prog);
      }
      (void)cudaGetLastError();
    }
  }
  return B2Y_OK;
}

extern "C" int b2y_plan_graph_launch(b2y_plan *pl, void *stream) {
  B2Y_REQUIRE(pl && pl->graph_exec, B2Y_ERR_INVALID, "graph_launch: graph not built");
  B2YDeviceGuard guard(pl->device);
  B2Y_CUDA_CHECK(cudaGraphLaunch(pl->graph_exec, (cudaStream_t)stream));
  if (pl->mid_event && pl->mid_op >= 0) pl->mid_recorded = true;
  return B2Y_OK;
}

// Per-layer autotuning of the tcgen05 conv tile configuration (blocks per tile x co-resident CTAs): every candidate is
// built and timed on the real buffers (their contents are overwritten by the next full run), the fastest is kept.
extern "C" int b2y_plan_autotune(b2y_plan *pl, void *stream) {
  B2Y_REQUIRE(pl, B2Y_ERR_INVALID, "autotune: null plan");
  B2YDeviceGuard guard(pl->device);
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  B2Y_CUDA_CHECK(cudaEventCreate(&e0));
  B2Y_CUDA_CHECK(cudaEventCreate(&e1));
  const int mbs[4] = {1, 2, 4, 8};
  const char *vb = getenv("B2Y_VERBOSE");
  const bool verbose = vb && vb[0] == '1';
  int op_index = -1;
  for (auto &r : pl->ops) {
    ++op_index;
    if (r.op.kind != B2Y_OP_CONV || r.impl != B2Y_IMPL_TCGEN05) continue;
    float best_ms = 1e30f;
    ConvTcPlan *best = nullptr;
    int best_mb = 0, best_ps = 0;
    if (verbose)
      fprintf(stderr, "[b2y autotune] op %3d %dx%d s%d cin %4d cout %4d map %dx%d:", op_index, r.cp.kh, r.cp.kw, r.cp.sh,
              r.cp.cin, r.cp.cout, r.cp.Ho, r.cp.Wo);
    for (int per_sm = 1; per_sm <= 3; ++per_sm)
      for (int mi = 0; mi < 4; ++mi) {
        ConvTcPlan *cand = nullptr;
        if (conv_tc_plan_create(r.cp, pl->device, TcConfig{mbs[mi], per_sm}, &cand) != B2Y_OK) continue;
        float ms = 1e30f;
        bool ok = launch_conv_tc(cand, st) == B2Y_OK;  // warm-up
        if (ok) {
          // best of two bursts of 4 back-to-back launches (PDL-chained like the real stream): robust to a stray hiccup
          for (int burst = 0; burst < 2 && ok; ++burst) {
            float t = 1e30f;
            cudaEventRecord(e0, st);
            for (int it = 0; it < 4 && ok; ++it) ok = launch_conv_tc(cand, st) == B2Y_OK;
            cudaEventRecord(e1, st);
            ok = ok && cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&t, e0, e1) == cudaSuccess;
            if (ok && t < ms) ms = t;
          }
        }
        if (verbose && ok) fprintf(stderr, " (%d,%d)=%.1fus", mbs[mi], per_sm, ms * 1000.f / 4.f);
        if (ok && ms < best_ms) {
          if (best) conv_tc_plan_destroy(best);
          best = cand;
          best_ms = ms;
          best_mb = mbs[mi];
          best_ps = per_sm;
        } else {
          conv_tc_plan_destroy(cand);
        }
      }
    if (verbose) fprintf(stderr, "  -> mb %d per_sm %d\n", best_mb, best_ps);
    if (best) {
      conv_tc_plan_destroy(r.tc);
      r.tc = best;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  B2Y_CUDA_CHECK(cudaGetLastError());
  if (pl->graph_exec) { cudaGraphExecDestroy(pl->graph_exec); pl->graph_exec = nullptr; }
  if (pl->graph) { cudaGraphDestroy(pl->graph); pl->graph = nullptr; }
  return B2Y_OK;
}

extern "C" int b2y_plan_num_launches(const b2y_plan *pl) { return pl ? (int)pl->ops.size() : 0; }

extern "C" int b2y_plan_op_impl(const b2y_plan *pl, int32_t i) {
  if (!pl || i < 0 || i >= (int)pl->ops.size()) return -1;
  return pl->ops[i].impl;
}

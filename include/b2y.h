/* b2y.h -- C ABI of libb2y.so: the B200 (sm_100a) detection inference path that drops in behind the
 * BS-YOLO / Ultralytics fork's predict() hot path.
 *
 * The reference has no native interface on this path (it is torch eager + torchvision + OpenCV);
 * each entry point therefore cites the reference *Python* interface it replaces.  Plain pointers and sizes
 * only; device pointers are raw `void*` (e.g. torch `tensor.data_ptr()`), streams are `cudaStream_t` passed as
 * `void*`.  Every function returns B2Y_OK (0) or a negative error code, never throws, never synchronises the
 * device unless stated, and keeps no global state besides the last-error string (thread local).
 */
#ifndef B2Y_H_
#define B2Y_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2Y_ABI_VERSION 7
#if defined(__GNUC__)
#define B2Y_API __attribute__((visibility("default")))
#else
#define B2Y_API
#endif

enum {
  B2Y_OK = 0,
  B2Y_ERR_INVALID = -1,     /* bad argument / unsupported shape   -> ValueError in the Python shim        */
  B2Y_ERR_UNSUPPORTED = -2, /* op not on the B200 path            -> NotImplementedError (no CPU fallback) */
  B2Y_ERR_CUDA = -3,        /* CUDA runtime / driver error        -> RuntimeError                          */
  B2Y_ERR_NOMEM = -4        /* workspace too small                -> MemoryError                           */
};

/* ---- plan: the lowered network (replaces BaseModel._predict_once, ultralytics/nn/tasks.py:138-165, and the
 *      module forwards in nn/modules/{conv,block,head}.py + nn/Addmodules/{MSCA,ELA}.py) ------------------- */

enum b2y_dtype { B2Y_BF16 = 0, B2Y_F32 = 1 };

enum b2y_op_kind {
  B2Y_OP_CONV = 1,      /* dense conv k{1,3} s{1,2} + bias (+SiLU) (+residual, b2y_res_mode); conv.py:145-151.  ip[0]=1: 2x2
                           conv padded before only (the 3x3 s2 stem on the space-to-depth input); ip[1]: fused Detect
                           decode epilogue 1 = DFL+dist2bbox, 2 = class sigmoid (head.py:100-131) writing y (B,4+nc,A):
                           ip[2] anchor offset, ip[3] A, ip[4] nc, fp[0] stride; ip[5]=1: the weight matrix is diagonal
                           (a depthwise 3x3 lowered to the tensor cores) and may be evaluated in groups of 16 channels   */
  B2Y_OP_DWCONV = 2,    /* depthwise kh x kw conv + bias (+SiLU) (+residual add); conv.py:224-229, MSCA.py:24-39.  ip[0] > 0:
                           chunked input view -- output channel j reads input channel (j / ip[0]) * ip[1] + j % ip[0] of the view
                           (Attention.pe over the v slices of qkv, block.py:4286; 3x3 only)                             */
  B2Y_OP_SPPF_POOL = 4, /* three chained MaxPool2d(5,1,2) written to three channel slices; block.py:3145-3149   */
  B2Y_OP_RESAMPLE = 5,  /* channel-slice copy, optionally nearest 2x upsample (nn.Upsample, Concat conv.py:455)  */
  B2Y_OP_ATTN = 6,      /* softmax(q^T k * kd^-0.5) v per head; block.py:4280-4286                               */
  B2Y_OP_ROWCOL_MEAN = 7, /* ELA: row / column / global means; ELA.py:74-84                                      */
  B2Y_OP_ELA_GATE = 8,  /* ELA: dilated conv1d + GroupNorm + sigmoid gates; ELA.py:77-86                         */
  B2Y_OP_ELA_APPLY = 9, /* ELA: x * (a*ch + b*h*w) + c*x; ELA.py:94-100                                          */
  B2Y_OP_GAP4 = 10,     /* MSCA: global average pool of the four branches; MSCA.py:71-74                         */
  B2Y_OP_MSCA_GATE = 11,/* MSCA: SE 1x1 convs + softmax(sigmoid()) over branches; MSCA.py:71-77                  */
  B2Y_OP_WSUM4 = 12,    /* MSCA: gate-weighted sum of the four branches; MSCA.py:84                              */
  B2Y_OP_DECODE = 13,   /* Detect._inference: DFL + dist2bbox + stride + sigmoid; head.py:100-131                */
  B2Y_OP_CONV_PAIR = 14,/* Bottleneck in one kernel: out = x + SiLU(conv3x3(SiLU(conv3x3(x))))  (block.py:3405-3419), the
                           hidden map never leaves the SM.  cin = cout in {16,32,64}; ip[0] = hidden channels (padded
                           to 16/32/64), ip[1] = 1: add the input; w_off/b_off = first conv, p_off[0]/p_off[1] = second */
  B2Y_OP_PMSFA_DW = 16, /* depthwise tail of a PMSFA block in one kernel (block.py:3046-3051): in = conv1's output [a | b | h]
                           (c channels: quarters a, b and the half h), out = [SiLU(dw7x7(SiLU(dw5x5([a|b]))_a)) |
                           SiLU(dw5x5([a|b]))_b | h] = the input of conv4.  cin = cout = c in {16,32,64,128};
                           w_off/b_off = conv2 fp32 [25][c/2] / [c/2], p_off[0]/p_off[1] = conv3 fp32 [49][c/4] / [c/4] */
  B2Y_OP_MSCA_BRANCH = 17, /* MSCAAttention up to the branch outputs in one kernel (MSCA.py:52-74): attn = dw5x5(x); branch j =
                           dw(k_j x 1)(dw(1 x k_j)(attn)) for k = 5, 7, 11, 21 (the shared per-channel `dilconv` folded into
                           the second strip of branches 0..2); out = the four branch maps as channel slices [jC, (j+1)C) of
                           one 4C-channel buffer, res = gap (B, 4, C) fp32 = their global average pools.  cin = C,
                           cout = 4C; w_off = packed fp32 parameters (see plan.msca) */
  B2Y_OP_D2S = 18       /* depth-to-space: out[b, 2y+dy, 2x+dx, c] = in[b, y, x, (dy*2+dx)*cout + c], cin = 4*cout, cout % 8 == 0.
                           With the preceding 1x1 conv of 4*cout channels it is nn.ConvTranspose2d(c, c, 2, 2, 0) --
                           Proto.upsample of a Segment head (block.py:91) */
};

enum b2y_conv_impl { B2Y_IMPL_AUTO = 0, B2Y_IMPL_SIMT = 1, B2Y_IMPL_TCGEN05 = 2 };
enum b2y_res_mode {
  B2Y_RES_NONE = 0,
  B2Y_RES_ADD = 1,        /* out = act(conv) + res                                                            */
  B2Y_RES_MUL = 2,        /* out = act(conv) * res                                                            */
  B2Y_RES_PREADD_UP2 = 3  /* out = act(conv + nearest-2x-upsampled res): res is a half-resolution map.  A 1x1 conv over
                             Concat(Upsample(a), b) (yaml head: nn.Upsample + Concat + C3k2.cv1) equals
                             act(up(W_a a) + W_b b + bias), so neither the upsampled nor the concatenated tensor exists */
};

/* A buffer is one NHWC activation tensor (batch B is a plan property): B x H x W x C elements. */
typedef struct {
  int64_t offset; /* byte offset into the workspace (256-B aligned) */
  int32_t C, H, W;
  int32_t dtype;  /* b2y_dtype */
} b2y_buf;

/* One fused op.  Tensor operands are *views*: (buffer index, first channel); the view's channel count is
 * cin / cout.  Unused operands are -1.  Weights live in one device blob, addressed by byte offsets. */
typedef struct {
  int32_t kind;
  int32_t in[4], in_coff[4];
  int32_t out, out_coff;
  int32_t res, res_coff, res_mode;
  int32_t cin, cout;       /* logical channels read / written per pixel                                   */
  int32_t cout_pad;        /* rows in the packed weight matrix (>= cout, multiple of 16 for the tcgen05 path) */
  int32_t kh, kw, sh, sw, ph, pw;
  int32_t act;             /* 0 none, 1 SiLU */
  int32_t impl;            /* b2y_conv_impl */
  int64_t w_off, b_off;    /* weights / bias byte offsets (-1: none); extra parameter blocks follow in p_off */
  int64_t p_off[4];
  int32_t ip[8];           /* op-specific integers (nh, kd, hd | groups, ks | nc, level sizes ...)          */
  float fp[4];             /* op-specific floats                                                            */
} b2y_op;

typedef struct b2y_plan b2y_plan;

B2Y_API int b2y_abi_version(void);
B2Y_API const char *b2y_last_error(void);

/* Build a plan over caller-owned device memory (workspace: activations, weights: packed parameter blob).
 * Creates TMA descriptors and per-op launch geometry; does not launch anything. */
B2Y_API int b2y_plan_create(const b2y_op *ops, int32_t n_ops, const b2y_buf *bufs, int32_t n_bufs, void *workspace,
                    size_t workspace_bytes, const void *weights, size_t weights_bytes, int32_t batch, int32_t device,
                    b2y_plan **out);
B2Y_API int b2y_plan_destroy(b2y_plan *plan);
/* Enqueue ops [first, last) on `stream` (last = -1: to the end).  Replaces AutoBackend.forward -> self.model(im)
 * (ultralytics/nn/autobackend.py:523-524). */
B2Y_API int b2y_plan_run(b2y_plan *plan, int32_t first, int32_t last, void *stream);
/* Capture the whole plan into a CUDA graph (stream must not be capturing) / replay it. */
B2Y_API int b2y_plan_graph_build(b2y_plan *plan, void *stream);
B2Y_API int b2y_plan_graph_launch(b2y_plan *plan, void *stream);
/* Mid-point marker for callers that pipeline batches (no reference counterpart: BasePredictor.stream_inference,
 * ultralytics/engine/predictor.py:241-266, runs preprocess -> inference -> postprocess strictly in sequence).
 * set_mid_op: an event is recorded in front of op `op_index` in every later run / graph launch (-1: none; call before
 * b2y_plan_graph_build).  wait_mid: make `stream` wait for that point of the most recent launch (no-op if the plan has not
 * been launched with a marker yet).  The wide, bandwidth-bound layers come first in a detect network; kernels parked behind
 * the marker (the NMS of the previous batch, the preprocess of the next) run under the thin, latency-bound layers. */
B2Y_API int b2y_plan_set_mid_op(b2y_plan *plan, int32_t op_index);
B2Y_API int b2y_plan_wait_mid(b2y_plan *plan, void *stream);
/* Time the candidate tile configurations of every tcgen05 conv on the plan's own buffers and keep the fastest
 * (synchronises the stream; buffer contents are clobbered; an existing graph is dropped). */
B2Y_API int b2y_plan_autotune(b2y_plan *plan, void *stream);
/* Number of kernels one full run launches; which conv implementation op `i` resolved to (b2y_conv_impl). */
B2Y_API int b2y_plan_num_launches(const b2y_plan *plan);
B2Y_API int b2y_plan_op_impl(const b2y_plan *plan, int32_t i);
/* Planner query, no device needed: 1 if B2Y_OP_CONV_PAIR has a kernel configuration for these channels and map size. */
B2Y_API int b2y_conv_pair_fits(int32_t cin, int32_t chid, int32_t B, int32_t H, int32_t W);

/* ---- preprocess: LetterBox + BGR->RGB + /255 (ultralytics/data/augment.py:1535-1601,
 *      engine/predictor.py:116-134).  src: B x h x w x 3 uint8 (device).  The resized image (uw x uh, OpenCV
 *      INTER_LINEAR fixed-point semantics; identity when uw==w && uh==h) is placed at (top,left) of an H x W canvas
 *      padded with 114.  dst_format 0: the network input, NHWC bf16 space-to-depth: (B, H/2, W/2, 16) with channel
 *      ((y&1)*2 + (x&1))*4 + {R,G,B,0} -- the 3x3 s2 stem becomes a 2x2 tensor-core conv over it; 1: NCHW fp32 (the
 *      reference's tensor); 2: HWC uint8 letterboxed BGR (LetterBox output, for bit-exact tests). */
B2Y_API int b2y_preprocess(const void *src, int32_t B, int32_t h, int32_t w, void *dst, int32_t H, int32_t W, int32_t top,
                   int32_t left, int32_t uh, int32_t uw, int32_t dst_format, void *stream);

/* ---- decode: Detect._inference (ultralytics/nn/modules/head.py:100-131; DFL block.py:73-76; make_anchors /
 *      dist2bbox utils/tal.py:371-395).  raw[l]: B x H_l x W_l x (64+nc) fp32 NHWC head maps.
 *      y: B x (4+nc) x A fp32, rows cx,cy,w,h,p0.. (the reference's layout). */
B2Y_API int b2y_decode(const void *const *raw, const int32_t *hs, const int32_t *ws, const float *strides, int32_t n_levels,
               int32_t B, int32_t nc, void *y, void *stream);

/* ---- NMS: ops.non_max_suppression (ultralytics/utils/ops.py:167-316) + torchvision.ops.nms, bit-exact.
 *      pred: B x (4+nc) x A fp32 (not modified).  dets: B x max_det x 6 fp32 rows [x1,y1,x2,y2,conf,cls] in
 *      descending-score order, counts: B int32.  kept_anchor (optional, may be NULL): B x max_det int32 anchor index
 *      of every kept row.  classes: optional device-side filter list (host pointer, copied at call time).
 *      Documented deviations: no wall-clock abort (ops.py:312-314), input not mutated (ops.py:244). */
B2Y_API size_t b2y_nms_workspace_bytes(int32_t B, int32_t nc, int32_t A, int32_t multi_label, int32_t max_nms);
B2Y_API int b2y_nms(const void *pred, int32_t B, int32_t nc, int32_t A, float conf_thres, double iou_thres, int32_t max_det,
            int32_t max_nms, float max_wh, int32_t agnostic, int32_t multi_label, const int32_t *classes,
            int32_t n_classes, void *workspace, size_t workspace_bytes, void *dets, void *counts, void *kept_anchor,
            void *stream);

/* ---- scale_boxes + clip_boxes (ultralytics/utils/ops.py:92-127, 319-338) applied in place to dets
 *      (B x max_det x 6); one (gain, padx, pady, w0, h0) record of 5 floats per image (device pointer). */
B2Y_API int b2y_scale_boxes(void *dets, const void *counts, const void *geom, int32_t B, int32_t max_det, void *stream);

/* ---- validator path: DetectionValidator.preprocess / update_metrics on the device ---------------------
 *      b2y_preprocess_chw: DetectionValidator.preprocess (ultralytics/models/yolo/detect/val.py:47-50): the dataloader's
 *      letterboxed (B,3,H,W) uint8 RGB batch -> float / 255, written in the network-input layout of b2y_preprocess
 *      (dst_format 0).  H even, W % 4 == 0. */
B2Y_API int b2y_preprocess_chw(const void *src, int32_t B, int32_t H, int32_t W, void *dst, void *stream);
/*      b2y_box_iou: metrics.box_iou (ultralytics/utils/metrics.py:52-71), bit-exact.  box1 (n,4), box2 (m,4) xyxy fp32,
 *      iou (n,m) fp32 row-major. */
B2Y_API int b2y_box_iou(const void *box1, int32_t n, const void *box2, int32_t m, float eps, void *iou, void *stream);
/*      b2y_match_predictions: DetectionValidator._process_batch (val.py:209-228) = box_iou(gt, det) +
 *      BaseValidator.match_predictions (ultralytics/engine/validator.py:222-258, use_scipy=False) for a whole batch, bit-exact.
 *      dets: B x max_det x 6 fp32 [x1,y1,x2,y2,conf,cls] + counts (B int32), as b2y_nms / b2y_scale_boxes leave them;
 *      labels: B x max_labels x 5 fp32 [cls,x1,y1,x2,y2] in the same coordinate space + label_counts (B int32);
 *      iouv: niou (<= 16) positive thresholds (host pointer; torch.linspace(0.5, 0.95, 10), val.py:41);
 *      tp: B x max_det x niou uint8 (rows past counts[b] are zero).  Exact IoU ties between two labels of one detection go
 *      to the higher label index (numpy's order for the short match lists it insertion-sorts). */
B2Y_API int b2y_match_predictions(const void *dets, const void *counts, const void *labels, const void *label_counts,
                          const float *iouv, int32_t niou, int32_t B, int32_t max_det, int32_t max_labels, void *tp,
                          void *stream);

/* ---- Detect-derived heads: the extra per-anchor columns of a Pose / Segment head --------------------------------
 *      b2y_gather_extras replaces, for the rows NMS kept, Pose.kpts_decode (ultralytics/nn/modules/head.py:252-279), the
 *      "mask" columns non_max_suppression carries along (ultralytics/utils/ops.py:245-246, 268-273) and
 *      PosePredictor.postprocess's ops.scale_coords + clip_coords (ultralytics/models/yolo/pose/predict.py:52-56,
 *      utils/ops.py:740-772, 341-358).
 *      raw[l]: B x H_l x W_l x pitch fp32 NHWC maps of the cv4 branch (head.py:245; the plan's extra buffers), levels in
 *      anchor order (P3 | P4 | P5, row-major); kept_anchor / counts: as b2y_nms wrote them.
 *      mode 0: copy the ne channels; mode 1: keypoints in groups of ndim (2 | 3): x, y = (v * 2 + cell) * stride (bit-exact
 *      fp32), visibility = sigmoid(v).  geom (optional device pointer): one (gain, padx, pady, w0, h0) record of 5 floats
 *      per image -- x, y are then mapped to the original image and clipped like ops.scale_coords (pad NOT rounded, unlike
 *      scale_boxes).  round_dets (optional): the (B, max_det, 6) detections whose box corners are rounded half-to-even in
 *      place (pose/predict.py:52 `.round()`).  out: B x max_det x ne fp32, rows past counts[b] are zero. */
B2Y_API int b2y_gather_extras(const void *const *raw, const int32_t *hs, const int32_t *ws, const float *strides,
                      int32_t n_levels, int32_t pitch, int32_t ne, int32_t ndim, int32_t mode, const void *kept_anchor,
                      const void *counts, const void *geom, int32_t B, int32_t max_det, void *out, void *round_dets,
                      void *stream);

/*      b2y_process_mask: ops.process_mask(protos, masks_in, bboxes, shape, upsample=True) (ultralytics/utils/ops.py:663-693,
 *      crop_mask :644-660) for all kept rows of a batch in one launch -- the per-image loop of
 *      SegmentationPredictor.postprocess (ultralytics/models/yolo/segment/predict.py:48-56, retina_masks=False).
 *      protos: B x mh x mw x proto_cs bf16 NHWC mask prototypes (nm channels used; the plan's proto buffer), mh = H / 4,
 *      mw = W / 4; coeffs: B x max_det x nm fp32 (b2y_gather_extras mode 0); dets: B x max_det x 6 fp32 with boxes in
 *      network-input pixels (BEFORE b2y_scale_boxes).  row_img / row_det (device int32, `rows` entries): image and
 *      detection index of every output row (the ragged concatenation of the images' kept rows).
 *      masks: rows x H x W fp32, 1.0 where the upsampled cropped mask is > 0 else 0.0 (the reference's `masks.gt_(0.0)`). */
B2Y_API int b2y_process_mask(const void *protos, int32_t proto_cs, int32_t nm, int32_t mh, int32_t mw, const void *coeffs,
                     const void *dets, int32_t max_det, const void *row_img, const void *row_det, int32_t rows, int32_t H,
                     int32_t W, void *masks, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2Y_H_ */

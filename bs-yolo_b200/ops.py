"""Host-side mirror of the reference's tensor utilities for the detect path, backed by the CUDA kernels.

Same names, argument meaning and error behaviour as ultralytics/utils/ops.py so parity tests read like the
reference's own (`test_utils_ops`, tests/test_python.py:412-438):
  non_max_suppression  <- ultralytics/utils/ops.py:167-316   (one C-ABI call: b2y_nms)
  scale_boxes          <- ultralytics/utils/ops.py:92-127
  letterbox/preprocess <- ultralytics/data/augment.py:1535-1601, engine/predictor.py:116-134 (b2y_preprocess)
There is no CPU fallback: CPU tensors are rejected.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi as A
from .engine import letterbox_geometry

_ws_cache: dict = {}


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _workspace(dev, nbytes: int) -> torch.Tensor:
    # one scratch buffer per (device, stream): NMS calls on different streams (the pipeline's post-process stream beside a
    # caller's own stream) must not share slot scores / sort keys
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _ws_cache[key] = t
    return t


def nms_raw(pred: torch.Tensor, conf=0.25, iou=0.7, max_det=300, classes=None, agnostic=False, multi_label=False,
            max_nms=30000, max_wh=7680.0, return_anchors=False):
    """pred (B, 4+nc, A) fp32 CUDA -> dets (B, max_det, 6) fp32, counts (B,) int32 [, kept anchor index (B, max_det)]."""
    if not pred.is_cuda:
        raise A.B2YError("non_max_suppression: CUDA tensor required (no CPU fallback on the B200 path)")
    if pred.dtype != torch.float32:
        raise TypeError("non_max_suppression: prediction must be float32")
    pred = pred.contiguous()
    B, ch, An = pred.shape
    nc = ch - 4
    L = A.lib()
    nbytes = L.b2y_nms_workspace_bytes(B, nc, An, int(bool(multi_label)), max_nms)
    ws = _workspace(pred.device, nbytes)
    # b2y_nms writes every element (rows past counts[b] are zero-filled by the kernel): no torch fill launches
    dets = torch.empty(B, max_det, 6, dtype=torch.float32, device=pred.device)
    counts = torch.empty(B, dtype=torch.int32, device=pred.device)
    anchors = torch.empty(B, max_det, dtype=torch.int32, device=pred.device) if return_anchors else None
    if classes is not None:
        cl = (C.c_int32 * len(classes))(*[int(c) for c in classes])
        ncl = len(classes)
    else:
        cl, ncl = None, 0
    A.check(L.b2y_nms(pred.data_ptr(), B, nc, An, float(conf), float(iou), int(max_det), int(max_nms), float(max_wh),
                      int(bool(agnostic)), int(bool(multi_label)), cl, ncl, ws.data_ptr(), ws.numel(), dets.data_ptr(),
                      counts.data_ptr(), anchors.data_ptr() if anchors is not None else None, _stream(pred.device)))
    return (dets, counts, anchors) if return_anchors else (dets, counts)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True,
                        rotated=False):
    """Drop-in for ultralytics.utils.ops.non_max_suppression (ops.py:167-316) on CUDA tensors; returns a list of
    (n_i, 6) tensors [x1, y1, x2, y2, conf, cls], bit-identical rows in the same order.

    Documented deviations: the input is never modified (the reference mutates it when in_place, ops.py:244) and there
    is no wall-clock abort (ops.py:312-314, `max_time_img` is ignored)."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if rotated or (labels is not None and len(labels)):
        raise NotImplementedError("rotated boxes / apriori labels are not on the B200 detect path")
    if prediction.shape[-1] == 6:
        raise NotImplementedError("end-to-end (B,N,6) predictions are not on the B200 detect path")
    nc = nc or (prediction.shape[1] - 4)
    if prediction.shape[1] - nc - 4 != 0:
        raise NotImplementedError("mask coefficients (segment models) are not on the B200 detect path")
    dets, counts = nms_raw(prediction, conf_thres, iou_thres, max_det, classes, agnostic, multi_label, max_nms,
                           float(max_wh))
    n = counts.tolist()  # the one device->host sync of the post-process
    return [dets[i, : n[i]] for i in range(len(n))]


def _as_device_batch(images, device) -> torch.Tensor:
    if isinstance(images, torch.Tensor):
        t = images
    else:
        if len({im.shape for im in images}) != 1:
            raise ValueError("all images of a batch must share one shape on the B200 path (shard ragged batches by shape)")
        t = torch.from_numpy(np.ascontiguousarray(np.stack(images)))
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
        raise TypeError("expected uint8 images of shape (B, h, w, 3) in BGR order")
    return t.to(device, non_blocking=True).contiguous()


def _letterbox_call(images, imgsz, auto, stride, device, fmt):
    device = torch.device(device)
    if device.type != "cuda":
        raise A.B2YError("preprocess: CUDA device required (no CPU fallback on the B200 path)")
    src = _as_device_batch(images, device)
    B, h, w = int(src.shape[0]), int(src.shape[1]), int(src.shape[2])
    imgsz = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
    H, W, top, left, uh, uw = letterbox_geometry((h, w), imgsz, auto, stride)
    if fmt == 1:
        dst = torch.empty(B, 3, H, W, dtype=torch.float32, device=device)
    elif fmt == 2:
        dst = torch.empty(B, H, W, 3, dtype=torch.uint8, device=device)
    else:
        dst = torch.empty(B, H // 2, W // 2, 16, dtype=torch.bfloat16, device=device)
    A.check(A.lib().b2y_preprocess(src.data_ptr(), B, h, w, dst.data_ptr(), H, W, top, left, uh, uw, fmt, _stream(device)))
    return dst


def letterbox(images, imgsz=640, auto=False, stride=32, device="cuda"):
    """LetterBox(imgsz, auto, stride)(image=...) for a same-shape batch: uint8 (B,h,w,3) -> uint8 (B,H',W',3)."""
    return _letterbox_call(images, imgsz, auto, stride, device, 2)


def preprocess(images, imgsz=640, auto=True, stride=32, device="cuda"):
    """BasePredictor.preprocess for a same-shape batch of uint8 BGR images: -> (B,3,H',W') fp32 RGB in [0,1]."""
    return _letterbox_call(images, imgsz, auto, stride, device, 1)


def scale_geometry(img1_shape, img0_shapes, device) -> torch.Tensor:
    """Per-image (gain, padx, pady, w0, h0) rows of scale_boxes (ops.py:109-117) as one fp32 device tensor.  The values
    are computed in Python doubles exactly like the reference and rounded to fp32 once; constant per (shape, shape) pair,
    so pipelines build it once (a per-call pageable upload would stall the post-process stream)."""
    geom = []
    for (h0, w0) in img0_shapes:
        gain = min(img1_shape[0] / h0, img1_shape[1] / w0)
        padx = round((img1_shape[1] - w0 * gain) / 2 - 0.1)
        pady = round((img1_shape[0] - h0 * gain) / 2 - 0.1)
        geom.append([gain, padx, pady, w0, h0])
    return torch.tensor(geom, dtype=torch.float32).to(device)


def scale_boxes_batched(dets: torch.Tensor, counts: torch.Tensor, img1_shape=None, img0_shapes=None, geom: torch.Tensor | None = None):
    """scale_boxes + clip_boxes (ops.py:92-127, 319-338) in place on padded dets (B, max_det, 6).  Pass a cached
    ``geom`` (scale_geometry) on hot paths."""
    B, max_det, _ = dets.shape
    if geom is None:
        geom = scale_geometry(img1_shape, img0_shapes, dets.device)
    assert geom.shape == (B, 5) and geom.dtype == torch.float32 and geom.device == dets.device
    A.check(A.lib().b2y_scale_boxes(dets.data_ptr(), counts.data_ptr(), geom.data_ptr(), B, max_det, _stream(dets.device)))
    return dets


# ---- Detect-derived heads --------------------------------------------------------------------------------
def coords_geometry(img1_shape, img0_shapes, device) -> torch.Tensor:
    """Per-image (gain, padx, pady, w0, h0) rows of scale_coords + clip_coords (ops.py:756-758, 341-358) as one fp32 device
    tensor.  Same as scale_geometry except that scale_coords does NOT round the pad (scale_boxes does, ops.py:113-116)."""
    geom = []
    for (h0, w0) in img0_shapes:
        gain = min(img1_shape[0] / h0, img1_shape[1] / w0)
        geom.append([gain, (img1_shape[1] - w0 * gain) / 2, (img1_shape[0] - h0 * gain) / 2, w0, h0])
    return torch.tensor(geom, dtype=torch.float32).to(device)


def gather_extras(raws, strides, ne: int, kept: torch.Tensor, counts: torch.Tensor, mode: int = 0, ndim: int = 3,
                  geom: torch.Tensor | None = None, round_dets: torch.Tensor | None = None) -> torch.Tensor:
    """The extra columns of the rows NMS kept (one C-ABI call: b2y_gather_extras).  raws: per-level (B, H_l, W_l, pitch)
    fp32 NHWC maps of a Detect-derived head's cv4 branch; kept (B, max_det) int32 anchors + counts (B,) int32 from
    ``nms_raw(return_anchors=True)``.  mode 1 = Pose.kpts_decode (head.py:252-279) on those rows, then -- with ``geom``
    (coords_geometry) -- ops.scale_coords to the original image; ``round_dets``: detections whose boxes get the
    `.round()` of pose/predict.py:52.  Returns (B, max_det, ne) fp32, zero past counts."""
    if not all(r.is_cuda for r in raws) or not kept.is_cuda:
        raise A.B2YError("gather_extras: CUDA tensors required (no CPU fallback on the B200 path)")
    nl = len(raws)
    B, max_det = int(kept.shape[0]), int(kept.shape[1])
    pitch = int(raws[0].shape[3])
    assert all(r.dtype == torch.float32 and r.is_contiguous() and r.shape[0] == B and r.shape[3] == pitch for r in raws)
    assert kept.dtype == torch.int32 and kept.is_contiguous() and counts.dtype == torch.int32 and counts.shape[0] == B
    assert geom is None or (geom.dtype == torch.float32 and tuple(geom.shape) == (B, 5) and geom.is_cuda)
    assert round_dets is None or (round_dets.dtype == torch.float32 and round_dets.is_contiguous()
                                  and tuple(round_dets.shape) == (B, max_det, 6))
    out = torch.empty(B, max_det, ne, dtype=torch.float32, device=kept.device)
    A.check(A.lib().b2y_gather_extras((C.c_void_p * nl)(*[r.data_ptr() for r in raws]),
                                      (C.c_int32 * nl)(*[int(r.shape[1]) for r in raws]),
                                      (C.c_int32 * nl)(*[int(r.shape[2]) for r in raws]),
                                      (C.c_float * nl)(*[float(v) for v in strides]), nl, pitch, int(ne), int(ndim), int(mode),
                                      kept.data_ptr(), counts.data_ptr(), geom.data_ptr() if geom is not None else None, B,
                                      max_det, out.data_ptr(), round_dets.data_ptr() if round_dets is not None else None,
                                      _stream(kept.device)))
    return out


def process_mask_batched(protos: torch.Tensor, coeffs: torch.Tensor, dets: torch.Tensor, counts, shape) -> list:
    """ops.process_mask(protos[i], pred[:, 6:], pred[:, :4], shape, upsample=True) (ops.py:663-693) for every image of a batch
    in ONE launch (b2y_process_mask).  protos: (B, mh, mw, pitch) bf16 NHWC (Engine.protos); coeffs (B, max_det, nm) fp32
    (gather_extras mode 0); dets (B, max_det, 6) with boxes in network-input pixels (before scale_boxes); counts: the
    per-image row counts as a host list.  Returns one (n_i, H, W) fp32 0/1 tensor per image (views of one allocation),
    None where n_i == 0 (segment/predict.py:49-50)."""
    if not (protos.is_cuda and coeffs.is_cuda and dets.is_cuda):
        raise A.B2YError("process_mask: CUDA tensors required (no CPU fallback on the B200 path)")
    B, mh, mw, pitch = (int(v) for v in protos.shape)
    H, W = int(shape[0]), int(shape[1])
    max_det, nm = int(coeffs.shape[1]), int(coeffs.shape[2])
    assert protos.dtype == torch.bfloat16 and protos.is_contiguous() and coeffs.dtype == torch.float32 and coeffs.is_contiguous()
    assert dets.dtype == torch.float32 and dets.is_contiguous() and tuple(dets.shape) == (B, max_det, 6) and len(counts) == B
    rows = int(sum(counts))
    if rows == 0:
        return [None] * B
    img = np.repeat(np.arange(B, dtype=np.int32), counts)
    det = np.concatenate([np.arange(n, dtype=np.int32) for n in counts])
    idx = torch.from_numpy(np.stack([img, det])).to(dets.device, non_blocking=True)
    masks = torch.empty(rows, H, W, dtype=torch.float32, device=dets.device)
    A.check(A.lib().b2y_process_mask(protos.data_ptr(), pitch, nm, mh, mw, coeffs.data_ptr(), dets.data_ptr(), max_det,
                                     idx[0].data_ptr(), idx[1].data_ptr(), rows, H, W, masks.data_ptr(), _stream(dets.device)))
    out, lo = [], 0
    for n in counts:
        out.append(masks[lo: lo + n] if n else None)
        lo += n
    return out


# ---- validator path ------------------------------------------------------------------------------------
def box_iou(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """Drop-in for ultralytics.utils.metrics.box_iou (metrics.py:52-71) on CUDA tensors: (N,4) x (M,4) xyxy -> (N,M),
    bit-identical (one C-ABI call: b2y_box_iou)."""
    if not (box1.is_cuda and box2.is_cuda):
        raise A.B2YError("box_iou: CUDA tensors required (no CPU fallback on the B200 path)")
    box1, box2 = box1.float().contiguous(), box2.float().contiguous()
    n, m = int(box1.shape[0]), int(box2.shape[0])
    out = torch.empty(n, m, dtype=torch.float32, device=box1.device)
    A.check(A.lib().b2y_box_iou(box1.data_ptr(), n, box2.data_ptr(), m, float(eps), out.data_ptr(), _stream(box1.device)))
    return out


def match_predictions_batched(dets: torch.Tensor, counts: torch.Tensor, labels: torch.Tensor, label_counts: torch.Tensor,
                              iouv) -> torch.Tensor:
    """DetectionValidator._process_batch (detect/val.py:209-228: box_iou + engine/validator.py:222-258 match_predictions)
    for a whole batch in one launch.  dets (B, max_det, 6) + counts (B,) as b2y_nms / b2y_scale_boxes leave them; labels
    (B, max_labels, 5) fp32 rows [cls, x1, y1, x2, y2] in the same coordinates + label_counts (B,) int32; iouv: the
    thresholds (val.py:41).  Returns tp (B, max_det, niou) bool."""
    if not (dets.is_cuda and labels.is_cuda):
        raise A.B2YError("match_predictions: CUDA tensors required (no CPU fallback on the B200 path)")
    B, max_det, _ = dets.shape
    assert dets.dtype == torch.float32 and dets.is_contiguous() and counts.dtype == torch.int32
    assert labels.dtype == torch.float32 and labels.is_contiguous() and labels.shape[0] == B and labels.shape[2] == 5
    assert label_counts.dtype == torch.int32 and label_counts.shape[0] == B
    th = [float(t) for t in (iouv.tolist() if hasattr(iouv, "tolist") else iouv)]
    tp = torch.empty(B, max_det, len(th), dtype=torch.uint8, device=dets.device)
    A.check(A.lib().b2y_match_predictions(dets.data_ptr(), counts.data_ptr(), labels.data_ptr(), label_counts.data_ptr(),
                                          (C.c_float * len(th))(*th), len(th), B, max_det, int(labels.shape[1]),
                                          tp.data_ptr(), _stream(dets.device)))
    return tp.bool()


def preprocess_chw(img_u8: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """DetectionValidator.preprocess (detect/val.py:47-50): the dataloader's letterboxed (B,3,H,W) uint8 RGB batch ->
    float / 255, written into the network-input buffer `dst` ((B, H/2, W/2, 16) bf16, Engine.input_nhwc)."""
    if not (img_u8.is_cuda and dst.is_cuda):
        raise A.B2YError("preprocess_chw: CUDA tensors required (no CPU fallback on the B200 path)")
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[1] == 3
    img_u8 = img_u8.contiguous()
    B, _, H, W = (int(v) for v in img_u8.shape)
    assert dst.dtype == torch.bfloat16 and dst.is_contiguous() and tuple(dst.shape) == (B, H // 2, W // 2, 16), dst.shape
    A.check(A.lib().b2y_preprocess_chw(img_u8.data_ptr(), B, H, W, dst.data_ptr(), _stream(dst.device)))
    return dst

"""ORACLE (test infrastructure, never shipped or measured as the product).

numpy restatement of the reference's pre- and post-processing for the detect path:
  letterbox ............... ultralytics/data/augment.py:1535-1601 (LetterBox.__call__), engine/predictor.py:145-161
  preprocess .............. ultralytics/engine/predictor.py:116-134
  cv2.resize INTER_LINEAR . OpenCV (third party, `opencv-python>=4.6.0` pyproject.toml:66; 4.13.0 here): 11-bit
                            fixed-point bilinear, recipe in SURVEY §8 "a1 fine print"; pinned against cv2 in tests
  non_max_suppression ..... ultralytics/utils/ops.py:167-316 (+ xywh2xyxy :416-433)
  torchvision.ops.nms ..... third party; restated in oracle/nms_ref.c (C) and ``nms_numpy`` below
  scale_boxes/clip_boxes .. ultralytics/utils/ops.py:92-127, 319-338
  scale_coords/clip_coords  ultralytics/utils/ops.py:740-772, 341-358 (Pose keypoints)
  pose_postprocess ........ ultralytics/models/yolo/pose/predict.py:39-59
  process_mask/crop_mask .. ultralytics/utils/ops.py:663-693, 644-660 (Segment masks; torch fp32 because the
                            reference's own bilinear F.interpolate is the definition)
  segment_postprocess ..... ultralytics/models/yolo/segment/predict.py:29-57 (retina_masks=False)

Deviations from the reference that are documented, not reproduced: the wall-clock abort of NMS
(ops.py:238,312-314) and the in-place mutation of the caller's tensor (ops.py:244).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force: bool = False) -> str:
    """Compile oracle/nms_ref.c -> oracle/_build/liboracle.so (gcc only)."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "nms_ref.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_nms.restype = ctypes.c_int32
        lib.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p]
        _LIB = lib
    return _LIB


# ----------------------------------------------------------------------------- torchvision.ops.nms
def nms_c(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), np.int32)
    k = _lib().oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, float(thr), keep.ctypes.data)
    return keep[:k].astype(np.int64)


def nms_numpy(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """Same algorithm as nms_ref.c, vectorised over the inner loop (fp32 ops, same order of operations)."""
    boxes = np.asarray(boxes, np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros(0, np.int64)
    x1, y1, x2, y2 = (boxes[:, k] for k in range(4))
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-np.asarray(scores, np.float32), kind="stable")
    sup = np.zeros(n, bool)
    keep = []
    zero = np.float32(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        for _i in range(n):
            i = order[_i]
            if sup[i]:
                continue
            keep.append(i)
            js = order[_i + 1:]
            w = np.maximum(zero, np.minimum(x2[i], x2[js]) - np.maximum(x1[i], x1[js]))
            h = np.maximum(zero, np.minimum(y2[i], y2[js]) - np.maximum(y1[i], y1[js]))
            w = np.where(np.isnan(w), zero, w)
            h = np.where(np.isnan(h), zero, h)
            inter = w * h
            ovr = inter / (areas[i] + areas[js] - inter)
            sup[js[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, np.int64)


# ----------------------------------------------------------------------------- ops.non_max_suppression
def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """utils/ops.py:416-433."""
    y = np.empty_like(x)
    wh = x[..., 2:4] / np.float32(2)
    y[..., :2] = x[..., :2] - wh
    y[..., 2:4] = x[..., :2] + wh
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        max_det=300, nc=0, max_nms=30000, max_wh=7680, nms_fn=None, return_indices=False):
    """utils/ops.py:167-316 on a (B, 4+nc, A) fp32 array. Returns list of (n_i, 6) fp32 arrays
    [x1,y1,x2,y2,conf,cls]; with ``return_indices`` also the anchor index of every kept row."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    nms_fn = nms_fn or nms_c
    pred = np.array(prediction, dtype=np.float32, copy=True)
    bs = pred.shape[0]
    nc = nc or (pred.shape[1] - 4)
    assert pred.shape[1] - nc - 4 == 0, "mask channels are not on the detect path"
    conf_t = np.float32(conf_thres)
    xc = pred[:, 4:4 + nc].max(1) > conf_t
    multi_label = bool(multi_label) and nc > 1
    pred = pred.transpose(0, 2, 1)
    pred = np.concatenate((xywh2xyxy(pred[..., :4]), pred[..., 4:]), -1)
    out, out_idx = [], []
    for xi in range(bs):
        x = pred[xi]
        aidx = np.nonzero(xc[xi])[0]
        x = x[aidx]
        if not x.shape[0]:
            out.append(np.zeros((0, 6), np.float32)); out_idx.append(np.zeros(0, np.int64)); continue
        box, cls = x[:, :4], x[:, 4:4 + nc]
        if multi_label:
            i, j = np.nonzero(cls > conf_t)
            x = np.concatenate((box[i], cls[i, j][:, None], j[:, None].astype(np.float32)), 1)
            aidx = aidx[i]
        else:
            j = cls.argmax(1)
            conf = cls[np.arange(cls.shape[0]), j]
            x = np.concatenate((box, conf[:, None], j[:, None].astype(np.float32)), 1)
            m = conf > conf_t
            x, aidx = x[m], aidx[m]
        if classes is not None:
            m = np.isin(x[:, 5], np.asarray(classes, np.float32))
            x, aidx = x[m], aidx[m]
        n = x.shape[0]
        if not n:
            out.append(np.zeros((0, 6), np.float32)); out_idx.append(np.zeros(0, np.int64)); continue
        if n > max_nms:  # reference uses an unstable argsort here; KATs keep scores distinct (SURVEY a11 fine print)
            o = np.argsort(-x[:, 4], kind="stable")[:max_nms]
            x, aidx = x[o], aidx[o]
        c = x[:, 5:6] * np.float32(0 if agnostic else max_wh)
        boxes = x[:, :4] + c
        i = nms_fn(boxes, x[:, 4], iou_thres)[:max_det]
        out.append(x[i]); out_idx.append(aidx[i])
    return (out, out_idx) if return_indices else out


# ----------------------------------------------------------------------------- letterbox / preprocess
def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """Integer emulation of cv2.resize(img, (dw, dh), INTER_LINEAR) on uint8 HWC (SURVEY §8 a1 fine print)."""
    sh, sw = img.shape[:2]

    def coeffs(dst, src):
        scale = 1.0 / (dst / src)
        f = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        f = f - s.astype(np.float32)
        lo = s < 0
        s[lo], f[lo] = 0, 0
        hi = s >= src - 1
        s[hi], f[hi] = src - 1, 0
        a1 = np.rint(f * np.float32(2048)).astype(np.int32)
        return s, 2048 - a1, a1

    xs, xa0, xa1 = coeffs(dw, sw)
    ys, ya0, ya1 = coeffs(dh, sh)
    src = img.astype(np.int32)
    x1 = np.minimum(xs + 1, sw - 1)
    rows = src[:, xs] * xa0[None, :, None] + src[:, x1] * xa1[None, :, None]  # (sh, dw, C) int32
    y1 = np.minimum(ys + 1, sh - 1)
    r0, r1 = rows[ys] >> 4, rows[y1] >> 4
    out = (((ya0[:, None, None] * r0) >> 16) + ((ya1[:, None, None] * r1) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_geometry(shape, new_shape=(640, 640), auto=False, stride=32, scaleup=True):
    """augment.py:1560-1588 geometry only. Returns (new_unpad_w, new_unpad_h, top, bottom, left, right)."""
    h, w = shape
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    uw, uh = int(round(w * r)), int(round(h * r))
    dw, dh = new_shape[1] - uw, new_shape[0] - uh
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return uw, uh, top, bottom, left, right


def letterbox(img: np.ndarray, new_shape=(640, 640), auto=False, stride=32, resize_fn=None) -> np.ndarray:
    """augment.py:1535-1601 for one uint8 HWC image; pad value 114."""
    h, w = img.shape[:2]
    uw, uh, top, bottom, left, right = letterbox_geometry((h, w), new_shape, auto, stride)
    if (w, h) != (uw, uh):
        img = (resize_fn or resize_linear_u8)(img, uw, uh)
    out = np.full((uh + top + bottom, uw + left + right, img.shape[2]), 114, np.uint8)
    out[top:top + uh, left:left + uw] = img
    return out


def preprocess(images, imgsz=(640, 640), stride=32, resize_fn=None) -> np.ndarray:
    """predictor.py:116-134 + :145-161. images: list of uint8 HWC BGR -> (B,3,H',W') fp32 RGB in [0,1]."""
    same = len({im.shape for im in images}) == 1
    lb = [letterbox(im, imgsz, auto=same, stride=stride, resize_fn=resize_fn) for im in images]
    im = np.stack(lb)[..., ::-1].transpose(0, 3, 1, 2)
    return np.ascontiguousarray(im).astype(np.float32) / np.float32(255)


# ----------------------------------------------------------------------------- scale_boxes
def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape) -> np.ndarray:
    """utils/ops.py:92-127 + clip_boxes :319-338 (padding=True, xywh=False). boxes (n,>=4) fp32, returns a copy."""
    boxes = np.array(boxes, np.float32, copy=True)
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    padx = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pady = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    boxes[..., 0] -= np.float32(padx); boxes[..., 2] -= np.float32(padx)
    boxes[..., 1] -= np.float32(pady); boxes[..., 3] -= np.float32(pady)
    boxes[..., :4] /= np.float32(gain)
    boxes[..., 0] = boxes[..., 0].clip(0, img0_shape[1]); boxes[..., 2] = boxes[..., 2].clip(0, img0_shape[1])
    boxes[..., 1] = boxes[..., 1].clip(0, img0_shape[0]); boxes[..., 3] = boxes[..., 3].clip(0, img0_shape[0])
    return boxes


# ----------------------------------------------------------------------------- Pose head post-process
def scale_coords(img1_shape, coords: np.ndarray, img0_shape) -> np.ndarray:
    """utils/ops.py:740-772 + clip_coords :341-358 (ratio_pad=None, normalize=False, padding=True).  coords (..., >=2) fp32,
    returns a copy.  Unlike scale_boxes the pad is NOT rounded; the Python doubles act on the fp32 tensor as fp32 scalars."""
    coords = np.array(coords, np.float32, copy=True)
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    coords[..., 0] -= np.float32(pad[0])
    coords[..., 1] -= np.float32(pad[1])
    coords[..., 0] /= np.float32(gain)
    coords[..., 1] /= np.float32(gain)
    coords[..., 0] = coords[..., 0].clip(0, img0_shape[1])
    coords[..., 1] = coords[..., 1].clip(0, img0_shape[0])
    return coords


def nms_with_extras(prediction, nc, **kw):
    """utils/ops.py:167-316 with nm = rows - 4 - nc extra ("mask") columns: they take no part in any decision and ride
    along with the kept rows (:245-246, 268-273).  prediction (B, 4+nc+nm, A) -> list of (n_i, 6+nm)."""
    pred = np.asarray(prediction, np.float32)
    out, idx = non_max_suppression(pred[:, :4 + nc], nc=nc, return_indices=True, **kw)
    return [np.concatenate((o, pred[b][4 + nc:][:, i].T), 1) for b, (o, i) in enumerate(zip(out, idx))]


def pose_postprocess(prediction, nc, kpt_shape, img1_shape, img0_shapes, conf=0.25, iou=0.7, max_det=300, classes=None,
                     agnostic=False):
    """models/yolo/pose/predict.py:39-59: NMS, boxes scaled to the original image and ROUNDED (:52), keypoints reshaped to
    kpt_shape and mapped with scale_coords (:53-54).  Returns per image (boxes (n,6), keypoints (n, *kpt_shape))."""
    res = []
    rows = nms_with_extras(prediction, nc, conf_thres=conf, iou_thres=iou, max_det=max_det, classes=classes, agnostic=agnostic)
    for r, s0 in zip(rows, img0_shapes):
        boxes = r[:, :6].copy()
        boxes[:, :4] = np.round(scale_boxes(img1_shape, boxes[:, :4], s0))  # np.round == torch.round: half to even
        k = r[:, 6:].reshape(len(r), *kpt_shape)
        res.append((boxes, scale_coords(img1_shape, k, s0)))
    return res


def keypoints_result(kpts: np.ndarray) -> np.ndarray:
    """engine/results.py:1305-1309 (Keypoints.__init__): what ``Results.keypoints.data`` holds -- x, y of points whose
    visibility is < 0.5 are zeroed.  Returns a copy."""
    k = np.array(kpts, np.float32, copy=True)
    if k.ndim == 3 and k.shape[2] == 3:
        k[..., :2][k[..., 2] < 0.5] = 0
    return k


# ----------------------------------------------------------------------------- Segment head post-process
def crop_mask(masks, boxes):
    """utils/ops.py:644-660.  masks (n, h, w), boxes (n, 4) torch tensors."""
    import torch

    _, h, w = masks.shape
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None], 4, 1)
    r = torch.arange(w, dtype=x1.dtype)[None, None, :]
    c = torch.arange(h, dtype=x1.dtype)[None, :, None]
    return masks * ((r >= x1) * (r < x2) * (c >= y1) * (c < y2))


def process_mask(protos, masks_in, bboxes, shape) -> np.ndarray:
    """utils/ops.py:663-693 with upsample=True.  protos (c, mh, mw), masks_in (n, c), bboxes (n, 4) in network-input pixels,
    shape = (H, W) of the network input.  Returns (n, H, W) fp32 of 0 / 1."""
    import torch
    import torch.nn.functional as F

    protos, masks_in = torch.as_tensor(np.asarray(protos, np.float32)), torch.as_tensor(np.asarray(masks_in, np.float32))
    bboxes = torch.as_tensor(np.array(bboxes, np.float32, copy=True))
    c, mh, mw = protos.shape
    ih, iw = shape
    masks = (masks_in @ protos.view(c, -1)).view(-1, mh, mw)
    width_ratio, height_ratio = mw / iw, mh / ih
    bboxes[:, 0] *= width_ratio
    bboxes[:, 2] *= width_ratio
    bboxes[:, 3] *= height_ratio
    bboxes[:, 1] *= height_ratio
    masks = crop_mask(masks, bboxes)
    masks = F.interpolate(masks[None], tuple(shape), mode="bilinear", align_corners=False)[0]
    return masks.gt_(0.0).numpy()


def segment_postprocess(prediction, protos, nc, img1_shape, img0_shapes, conf=0.25, iou=0.7, max_det=300, classes=None,
                        agnostic=False):
    """models/yolo/segment/predict.py:29-57 (retina_masks=False): NMS carrying the nm coefficient rows, masks from the
    UNSCALED boxes (:54), then boxes scaled to the original image (:55; not rounded).  Returns per image (boxes (n,6),
    masks (n, H, W) or None when nothing was kept (:49-50))."""
    res = []
    rows = nms_with_extras(prediction, nc, conf_thres=conf, iou_thres=iou, max_det=max_det, classes=classes, agnostic=agnostic)
    for b, (r, s0) in enumerate(zip(rows, img0_shapes)):
        if not len(r):
            res.append((r[:, :6], None))
            continue
        masks = process_mask(protos[b], r[:, 6:], r[:, :4], img1_shape)
        boxes = r[:, :6].copy()
        boxes[:, :4] = scale_boxes(img1_shape, boxes[:, :4], s0)
        res.append((boxes, masks))
    return res

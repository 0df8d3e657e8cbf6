"""ORACLE (test infrastructure, never shipped or measured as the product).

numpy restatement of the validator-side arithmetic of the detect path:
  box_iou ............... ultralytics/utils/metrics.py:52-71
  match_predictions ..... ultralytics/engine/validator.py:222-258 (use_scipy=False branch)
  process_batch ......... ultralytics/models/yolo/detect/val.py:209-228
  prepare_labels ........ ultralytics/models/yolo/detect/val.py:104-115 (+ xywh2xyxy ops.py:416-433, scale_boxes ops.py:92-127)
  preprocess_chw ........ ultralytics/models/yolo/detect/val.py:47-50 (img.float() / 255), in the network-input layout

Pinned by tests/golden/val_match.npz (outputs of the unmodified reference, tests/golden/make_golden.py) and by the live
comparison in tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import numpy as np

IOUV = np.linspace(0.5, 0.95, 10).astype(np.float32)  # == torch.linspace(0.5, 0.95, 10) (val.py:41); checked in the tests


def box_iou(box1: np.ndarray, box2: np.ndarray, eps: float = 1e-7) -> np.ndarray:
    """(N,4) x (M,4) xyxy -> (N,M) fp32, the reference's op order: inter / (area1 + area2 - inter + eps)."""
    a = np.asarray(box1, np.float32)[:, None, :]
    b = np.asarray(box2, np.float32)[None, :, :]
    wh = np.clip(np.minimum(a[..., 2:], b[..., 2:]) - np.maximum(a[..., :2], b[..., :2]), np.float32(0), None)
    inter = wh[..., 0] * wh[..., 1]
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        return (inter / (area_a + area_b - inter + np.float32(eps))).astype(np.float32)


def match_predictions(pred_classes: np.ndarray, true_classes: np.ndarray, iou: np.ndarray, iouv: np.ndarray = IOUV) -> np.ndarray:
    """(N,) predicted classes, (M,) label classes, iou (M,N) -> correct (N, len(iouv)) bool."""
    pred_classes = np.asarray(pred_classes, np.float32)
    true_classes = np.asarray(true_classes, np.float32)
    correct = np.zeros((pred_classes.shape[0], len(iouv)), bool)
    iou = np.asarray(iou, np.float32) * (true_classes[:, None] == pred_classes[None, :]).astype(np.float32)
    for i, thr in enumerate(np.asarray(iouv, np.float32)):
        m = np.array(np.nonzero(iou >= thr)).T  # (label, detection) pairs, row-major
        if m.shape[0]:
            if m.shape[0] > 1:
                m = m[iou[m[:, 0], m[:, 1]].argsort()[::-1]]          # by IoU, descending
                m = m[np.unique(m[:, 1], return_index=True)[1]]       # best label of every detection (-> detection order)
                m = m[np.unique(m[:, 0], return_index=True)[1]]       # first (lowest-index) detection of every label
            correct[m[:, 1].astype(int), i] = True
    return correct


def process_batch(detections: np.ndarray, gt_bboxes: np.ndarray, gt_cls: np.ndarray, iouv: np.ndarray = IOUV) -> np.ndarray:
    """detections (N,6) [xyxy, conf, cls]; gt (M,4) xyxy + (M,) classes -> tp (N, niou) bool."""
    return match_predictions(detections[:, 5], gt_cls, box_iou(gt_bboxes, detections[:, :4]), iouv)


def match_by_definition(detections: np.ndarray, gt_bboxes: np.ndarray, gt_cls: np.ndarray, iouv: np.ndarray = IOUV) -> np.ndarray:
    """The same result stated the way the CUDA kernel computes it (one best label per detection, then the lowest-index
    detection of every label per threshold); used to check the restatement above against itself on random inputs."""
    det = np.asarray(detections, np.float32)
    n, m = det.shape[0], len(gt_cls)
    tp = np.zeros((n, len(iouv)), bool)
    if n == 0 or m == 0:
        return tp
    iou = box_iou(gt_bboxes, det[:, :4]) * (np.asarray(gt_cls, np.float32)[:, None] == det[None, :, 5])
    best = np.full(n, -1)
    best_iou = np.zeros(n, np.float32)
    for d in range(n):
        for l in range(m):
            v = iou[l, d]
            if v > 0 and v >= best_iou[d]:
                best[d], best_iou[d] = l, v
    for i, thr in enumerate(np.asarray(iouv, np.float32)):
        taken = set()
        for d in range(n):
            if best[d] >= 0 and best_iou[d] >= thr and best[d] not in taken:
                taken.add(best[d])
                tp[d, i] = True
    return tp


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.float32)
    half = x[..., 2:] / np.float32(2)
    return np.concatenate([x[..., :2] - half, x[..., :2] + half], -1).astype(np.float32)


def prepare_labels(bboxes_xywhn: np.ndarray, imgsz, ori_shape, ratio_pad) -> np.ndarray:
    """Normalised xywh labels of one image -> xyxy in the ORIGINAL image (val.py:112-114 with ratio_pad given)."""
    h, w = imgsz
    b = xywh2xyxy(bboxes_xywhn) * np.array([w, h, w, h], np.float32)
    gain, (padx, pady) = ratio_pad[0][0], ratio_pad[1]
    b[:, [0, 2]] -= np.float32(padx)
    b[:, [1, 3]] -= np.float32(pady)
    b /= np.float32(gain)
    b[:, [0, 2]] = b[:, [0, 2]].clip(0, ori_shape[1])
    b[:, [1, 3]] = b[:, [1, 3]].clip(0, ori_shape[0])
    return b.astype(np.float32)


def preprocess_chw(img_u8: np.ndarray) -> np.ndarray:
    """(B,3,H,W) uint8 RGB -> (B,H/2,W/2,16) fp32 values of the bf16 network input: channel ((y&1)*2+(x&1))*4 + {R,G,B,0}."""
    import torch

    x = torch.from_numpy(np.ascontiguousarray(img_u8)).float() / 255
    x = x.to(torch.bfloat16).float()
    B, _, H, W = x.shape
    x4 = torch.cat([x, torch.zeros_like(x[:, :1])], 1)
    return x4.view(B, 4, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, H // 2, W // 2, 16).numpy()

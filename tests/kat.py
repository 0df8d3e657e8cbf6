"""Deterministic known-answer inputs shared by tests/golden/make_golden.py, the CPU oracle tests and the
GPU parity tests.  Inputs are regenerated from seeds (they are random, i.e. incompressible); only the
reference's OUTPUTS are committed under tests/golden/."""
from __future__ import annotations

import numpy as np

NET_CASES = {  # name -> (model, batch, H, W, seed)
    "yolo11n_64x96": ("yolo11n", 2, 64, 96, 11),
    "bsyolo_n_64x96": ("bsyolo_n", 2, 64, 96, 12),
    "yolo11x_64x64": ("yolo11x", 1, 64, 64, 13),
}


POSE_NET_CASES = {  # the stock pose graph (yolo11-pose.yaml), same tuple layout
    "yolo11n-pose_64x96": ("yolo11n-pose", 2, 64, 96, 14),
}


SEG_NET_CASES = {  # the stock segment graph (yolo11-seg.yaml): Proto + mask-coefficient branch
    "yolo11n-seg_64x96": ("yolo11n-seg", 2, 64, 96, 15),
}


def net_input(name):
    _, b, h, w, seed = (NET_CASES.get(name) or POSE_NET_CASES.get(name) or SEG_NET_CASES[name])
    return np.random.default_rng(seed).random((b, 3, h, w), dtype=np.float32)


def anchors_for(h, w):
    return sum((h // s) * (w // s) for s in (8, 16, 32))


def _rand_pred(rng, B, nc, A, img=640.0, wh=(10, 110), score_pow=8.0):
    """(B, 4+nc, A) decoded-prediction-like tensor: cx,cy in U(0,img), w,h in U(wh), scores U(0,1)^pow."""
    p = np.empty((B, 4 + nc, A), np.float32)
    p[:, 0:2] = rng.uniform(0, img, (B, 2, A))
    p[:, 2:4] = rng.uniform(wh[0], wh[1], (B, 2, A))
    p[:, 4:] = rng.random((B, nc, A)) ** score_pow
    return p


def nms_case(name):
    """Returns (prediction (B,4+nc,A) fp32, kwargs for non_max_suppression)."""
    rng = np.random.default_rng(abs(hash_name(name)))
    kw = dict(conf_thres=0.25, iou_thres=0.7, max_det=300)
    if name == "empty":  # zero candidates everywhere
        p = _rand_pred(rng, 2, 80, 2100)
        p[:, 4:] *= 0.2
    elif name == "single":  # exactly one candidate in image 0, none in image 1
        p = _rand_pred(rng, 2, 80, 2100)
        p[:, 4:] *= 0.2
        p[0, 4 + 17, 1234] = 0.9
    elif name == "random_2100":
        p = _rand_pred(rng, 3, 80, 2100, score_pow=3.0)
    elif name == "dense_8400":  # adversarial: nearly every anchor is a candidate, small image -> heavy overlap
        p = _rand_pred(rng, 2, 80, 8400, img=320.0, score_pow=1.0)
    elif name == "many_kept":  # > max_det survivors: exercises the [:max_det] cut
        p = _rand_pred(rng, 1, 80, 8400, img=4000.0, wh=(5, 20), score_pow=1.0)
    elif name == "ties":  # exact score ties -> stable order by anchor index
        p = _rand_pred(rng, 2, 80, 2100, score_pow=2.0)
        p[:, 4:] = np.round(p[:, 4:] * 8) / 8
    elif name == "duplicates_zero_area":  # duplicate boxes; zero-area duplicates must NOT suppress each other (0/0 = NaN)
        p = _rand_pred(rng, 2, 12, 2100, score_pow=1.5)
        p[:, :4, 1::2] = p[:, :4, 0::2]
        p[:, 2:4, 100:400] = 0.0
    elif name == "iou_at_threshold":  # pairs whose IoU sits exactly on / next to the threshold
        p = _rand_pred(rng, 1, 4, 2100, score_pow=1.0)
        p[:, 4:] *= 0.2
        for k in range(200):
            a = 2 * k
            cx, cy = 50.0 + 97 * (k % 6), 50.0 + 97 * (k // 6 % 6)
            w = 40.0
            # box b = box a shifted by d: IoU = (w-d)/(w+d); IoU = 0.5 at d = w/3; sweep around it
            d = w / 3 + (k - 100) * 1e-5
            p[0, :4, a] = (cx, cy, w, w)
            p[0, :4, a + 1] = (cx + d, cy, w, w)
            p[0, 4:, a] = 0
            p[0, 4:, a + 1] = 0
            p[0, 4 + (k // 36) % 4, a] = 0.9 - k * 1e-3
            p[0, 4 + (k // 36) % 4, a + 1] = 0.8 - k * 1e-3
        kw["iou_thres"] = 0.5
    elif name == "class_offset_79":  # large class ids: fp32 rounding of x + cls*7680 must be reproduced
        p = _rand_pred(rng, 2, 80, 2100, img=640.0, score_pow=1.0)
        p[:, 4:4 + 70] *= 0.1
    elif name == "agnostic":
        p = _rand_pred(rng, 2, 80, 2100, score_pow=3.0)
        kw["agnostic"] = True
    elif name == "classes_filter":
        p = _rand_pred(rng, 2, 80, 2100, score_pow=3.0)
        kw["classes"] = [0, 3, 17, 79]
    elif name == "multi_label":  # validator mode (models/yolo/detect/val.py:92-102)
        p = _rand_pred(rng, 2, 12, 600, score_pow=4.0)
        kw.update(multi_label=True, conf_thres=0.001, max_det=300)
    elif name == "nc12_conf_low":
        p = _rand_pred(rng, 2, 12, 8400, score_pow=6.0)
        kw.update(conf_thres=0.1, iou_thres=0.45, max_det=100)
    else:
        raise KeyError(name)
    return p, kw


NMS_CASES = ["empty", "single", "random_2100", "dense_8400", "many_kept", "ties", "duplicates_zero_area",
             "iou_at_threshold", "class_offset_79", "agnostic", "classes_filter", "multi_label", "nc12_conf_low"]


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


LETTERBOX_CASES = {  # name -> (src h, src w, imgsz, seed)
    "identity_160": (160, 160, 160, 1),
    "down_landscape": (270, 480, 160, 2),
    "down_portrait": (481, 321, 160, 3),
    "down_wide": (37, 500, 160, 4),
}


def letterbox_input(name):
    h, w, _, seed = LETTERBOX_CASES[name]
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


PREDICT_CASES = {  # name -> (model, n images, src h, src w, imgsz, seed)
    "yolo11n_160": ("yolo11n", 2, 160, 160, 160, 21),
    "bsyolo_n_120x160": ("bsyolo_n", 2, 240, 320, 160, 22),
}


POSE_PREDICT_CASES = {  # YOLO(pose ckpt).predict(): boxes + keypoints, letterboxed (240x320 -> 128x160 canvas, pad 4 rows)
    "yolo11n-pose_240x320": ("yolo11n-pose", 2, 240, 320, 160, 23),
    "yolo11n-pose_200x150": ("yolo11n-pose", 2, 200, 150, 160, 24),  # unrounded pad 20.0 px left/right
}


SEG_PREDICT_CASES = {  # YOLO(seg ckpt).predict(): boxes + (n, H', W') masks on the letterboxed canvas
    "yolo11n-seg_240x320": ("yolo11n-seg", 2, 240, 320, 160, 25),
    "yolo11n-seg_200x150": ("yolo11n-seg", 2, 200, 150, 160, 26),
}


def predict_input(name):
    _, n, h, w, _, seed = (PREDICT_CASES.get(name) or POSE_PREDICT_CASES.get(name) or SEG_PREDICT_CASES[name])
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(n)]


def state_dict_checksum(sd) -> float:
    return float(sum(float(v.double().abs().sum()) for v in sd.values()))


# ---- validator path: DetectionValidator._process_batch / match_predictions ---------------------------------
VAL_MATCH_CASES = {  # name -> (detections, labels, classes, seed, kind)
    "empty_dets": (0, 7, 3, 41, "random"),
    "empty_labels": (9, 0, 3, 42, "random"),
    "single_hit": (1, 1, 1, 43, "jitter"),
    "jitter_40x12": (40, 12, 4, 44, "jitter"),          # detections = jittered copies of the labels: IoU spread over 0.5..1
    "crowded_300x80": (300, 80, 3, 45, "jitter"),       # max_det detections, several per label, wrong-class look-alikes
    "random_250x150": (250, 150, 2, 46, "random"),
    "one_class_200x150": (200, 150, 1, 47, "jitter"),
    "lowest_index_wins": (6, 2, 1, 48, "quirk"),         # a label's LOWEST-INDEX detection wins, not its best (validator.py:254)
}


def val_match_case(name):
    """detections (n,6) fp32 [xyxy, conf, cls] sorted by descending conf (as NMS leaves them), gt boxes (m,4), gt classes (m,)."""
    n, m, nc, seed, kind = VAL_MATCH_CASES[name]
    rng = np.random.default_rng(seed)

    def boxes(k):
        c = rng.uniform(40, 600, (k, 2))
        wh = rng.uniform(16, 160, (k, 2))
        return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)

    gt = boxes(m)
    gcls = rng.integers(0, nc, m).astype(np.float32)
    if kind == "quirk":
        gt = np.array([[100, 100, 200, 200], [300, 300, 420, 400]], np.float32)
        gcls = np.zeros(2, np.float32)
        d = np.array([[104, 104, 204, 204], [100, 100, 200, 201], [100, 100, 200, 200], [302, 300, 422, 400],
                      [300, 300, 420, 400], [500, 500, 560, 560]], np.float32)
        det = np.concatenate([d, np.linspace(0.9, 0.4, 6, dtype=np.float32)[:, None], np.zeros((6, 1), np.float32)], 1)
        return det, gt, gcls
    if kind == "jitter" and m:
        src = rng.integers(0, m, n)
        wh = (gt[src, 2:] - gt[src, :2])
        det_b = gt[src] + (rng.normal(0, 0.06, (n, 4)) * np.concatenate([wh, wh], 1)).astype(np.float32)
        dcls = np.where(rng.random(n) < 0.8, gcls[src], rng.integers(0, nc, n)).astype(np.float32)
    else:
        det_b = boxes(n)
        dcls = rng.integers(0, nc, n).astype(np.float32)
    conf = np.sort(rng.random(n).astype(np.float32))[::-1]
    det = np.concatenate([det_b.astype(np.float32), conf[:, None], dcls[:, None]], 1).astype(np.float32)
    return det, gt, gcls

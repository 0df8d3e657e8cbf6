"""Tracking hook: the reference's ByteTrack callbacks (trackers/track.py:18-104, registered by
``YOLO.track``, engine/model.py:598-605) consume the final (n,6) boxes of the drop-in predictor exactly as they consume the
reference predictor's -- ``on_predict_start`` reads ``predictor.dataset``, ``on_predict_postprocess_end`` reads and rewrites
``predictor.results`` (CUDA ``Boxes`` here, ``.cpu().numpy()`` in the hook).

``lap`` is not in this image (the reference would try to pip-install it); tests/stubs/lap.py stands in for BOTH arms.
Gate: per frame, at least 90 % of the reference tracks are found among ours (same class, IoU >= 0.9), and the reference-id -> our-id mapping
is one consistent bijection over the whole sequence for at least 90 % of the tracks (ids are handed out in detection order,
so two near-equal confidences may swap a pair of ids; the Kalman-filtered boxes differ by the bf16 budget)."""
import os
import sys

import numpy as np
import pytest
import torch

from bs_yolo_b200 import arch
from oracle import ref_import

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")]


def _iou(a, b):
    iw = np.clip(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]), 0, None)
    ih = np.clip(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]), 0, None)
    inter = iw * ih
    area = lambda x: (x[:, 2] - x[:, 0]) * (x[:, 3] - x[:, 1])  # noqa: E731
    return inter / (area(a)[:, None] + area(b)[None, :] - inter + 1e-9)


def test_bytetrack_hook_on_dropin_predictor_vs_reference(tmp_path):
    stubs = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
    try:
        import lap  # noqa: F401
    except ImportError:
        sys.path.insert(0, stubs)
    ref_import.import_reference()
    from ultralytics import YOLO

    from bs_yolo_b200.predictor import B200DetectionPredictor

    spec = arch.build_spec("yolo11n")
    m = ref_import.reference_model("yolo11n", arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    m.args = {"task": "detect", "imgsz": 160}
    path = str(tmp_path / "yolo11n.pt")
    torch.save({"model": m, "ema": None, "train_args": {"task": "detect", "imgsz": 160}}, path)
    base = np.random.default_rng(9).integers(0, 256, (200, 200, 3), dtype=np.uint8)
    frames = [np.ascontiguousarray(base[4 + 2 * t: 164 + 2 * t, 6 + t: 166 + t]) for t in range(6)]  # a panning camera
    kw = dict(persist=True, tracker="bytetrack.yaml", imgsz=160, verbose=False)
    ref = YOLO(path).track(frames, device="cpu", **kw)
    yolo = YOLO(path)
    ours = yolo.track(frames, predictor=B200DetectionPredictor, device=0, **kw)
    assert type(yolo.predictor).__name__ == "B200DetectionPredictor" and hasattr(yolo.predictor, "trackers")
    assert len(ours) == len(ref) == len(frames)
    mapping, total, found = {}, 0, 0
    votes = {}
    for r, o in zip(ref, ours):
        assert r.boxes.id is not None and o.boxes.id is not None and o.boxes.is_track
        rb, ob = r.boxes.data.cpu().numpy(), o.boxes.data.cpu().numpy()  # [x1,y1,x2,y2,id,conf,cls]
        iou = _iou(rb[:, :4], ob[:, :4]) * (rb[:, None, 6] == ob[None, :, 6])
        for i in range(len(rb)):
            total += 1
            j = int(iou[i].argmax()) if len(ob) else -1
            if j >= 0 and iou[i, j] >= 0.9:
                found += 1
                votes.setdefault(int(rb[i, 4]), []).append(int(ob[j, 4]))
    print(f"tracks: {found} of {total} reference track boxes found among ours")
    assert total > 50 and found >= 0.9 * total, (found, total)  # a detection within the bf16 budget of new_track_thresh may not start a track
    consistent = sum(1 for v in votes.values() if len(set(v)) == 1)
    targets = [v[0] for v in votes.values() if len(set(v)) == 1]
    print(f"tracks: {len(votes)} reference ids, {consistent} mapped to one id of ours over all {len(frames)} frames")
    assert consistent >= 0.9 * len(votes) and len(set(targets)) == len(targets)

"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference, CPU)
on the deterministic inputs of tests/kat.py.  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, these vectors can.  What each file pins:
  net_<case>.npz ....... DetectionModel forward (nn/tasks.py:138-165, nn/modules/*.py, head.py:64-131) on the seeded
                         synthetic checkpoint of bs_yolo_b200.arch: decoded y and the three raw head maps
  nms.npz .............. utils/ops.py:167-316 non_max_suppression (+ torchvision.ops.nms) on every tests/kat.py NMS case
  letterbox.npz ........ data/augment.py:1535-1601 LetterBox (cv2.resize / copyMakeBorder)
  predict_<case>.npz ... YOLO(ckpt).predict(list of ndarrays, device="cpu") end to end: Results.boxes.data per image
  net_<pose case>.npz / predict_<pose case>.npz ... the same two for the stock pose graph: PoseModel forward
                         (head.py:242-279: y incl. the decoded keypoint rows, raw keypoint maps) and
                         YOLO(pose ckpt).predict() -> Results.boxes.data + Results.keypoints.data (pose/predict.py:39-59)
  net_<seg case>.npz / predict_<seg case>.npz ..... likewise for the stock segment graph: SegmentationModel forward (head.py:188-197:
                         y incl. the mask-coefficient rows, prototypes) and YOLO(seg ckpt).predict() -> Results.boxes.data +
                         Results.masks.data (segment/predict.py:29-57; masks bit-packed)
  val_match.npz ........ utils/metrics.py:52-71 box_iou and DetectionValidator._process_batch (detect/val.py:209-228 ->
                         engine/validator.py:222-258 match_predictions) on every tests/kat.py VAL_MATCH case
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kat  # noqa: E402
from bs_yolo_b200 import arch  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref_import.import_reference()
    from ultralytics import YOLO
    from ultralytics.data.augment import LetterBox
    from ultralytics.utils import ops

    torch.manual_seed(0)
    # ---- network forward
    for case, (model, *_rest) in kat.NET_CASES.items():
        spec = arch.build_spec(model)
        sd = arch.synth_state_dict(spec, 0)
        m = ref_import.reference_model(model, sd)
        x = torch.from_numpy(kat.net_input(case))
        with torch.no_grad():
            y, raws = m(x)
        np.savez_compressed(os.path.join(OUT, f"net_{case}.npz"), y=y.numpy(), raw0=raws[0].numpy(), raw1=raws[1].numpy(),
                            raw2=raws[2].numpy(), sd_checksum=np.float64(kat.state_dict_checksum(sd)))
        print("net", case, tuple(y.shape))

    # ---- NMS
    nms = {}
    for case in kat.NMS_CASES:
        p, kw = kat.nms_case(case)
        out = ops.non_max_suppression(torch.from_numpy(p.copy()), **kw)
        nms[f"{case}__n"] = np.array([o.shape[0] for o in out], np.int64)
        nms[f"{case}__rows"] = np.concatenate([o.numpy().reshape(-1, 6) for o in out], 0).astype(np.float32)
        nms[f"{case}__insum"] = np.float64(p.astype(np.float64).sum())
        print("nms", case, nms[f"{case}__n"].tolist())
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **nms)

    # ---- letterbox
    lb = {}
    for case, (h, w, imgsz, _seed) in kat.LETTERBOX_CASES.items():
        img = kat.letterbox_input(case)
        for auto in (False, True):
            out = LetterBox((imgsz, imgsz), auto=auto, stride=32)(image=img)
            lb[f"{case}__auto{int(auto)}"] = out
        print("letterbox", case, out.shape)
    np.savez_compressed(os.path.join(OUT, "letterbox.npz"), **lb)

    # ---- end-to-end predict through the reference's public API and its own checkpoint loader
    for case, (model, n, h, w, imgsz, seed) in kat.PREDICT_CASES.items():
        spec = arch.build_spec(model)
        sd = arch.synth_state_dict(spec, 0)
        m = ref_import.reference_model(model, sd)
        m.names = {k: f"class{k}" for k in range(spec["nc"])}
        m.args = {"task": "detect", "imgsz": imgsz}
        path = f"/tmp/b2y_golden_{case}.pt"
        torch.save({"model": m, "ema": None, "train_args": {"task": "detect", "imgsz": imgsz}}, path)
        yolo = YOLO(path)
        res = yolo.predict(kat.predict_input(case), device="cpu", imgsz=imgsz, conf=0.25, iou=0.7, max_det=300, verbose=False)
        d = {f"boxes{i}": r.boxes.data.numpy().astype(np.float32) for i, r in enumerate(res)}
        d["n"] = np.array([r.boxes.data.shape[0] for r in res], np.int64)
        np.savez_compressed(os.path.join(OUT, f"predict_{case}.npz"), **d)
        print("predict", case, d["n"].tolist())

    val_match()
    pose()
    segment()


def pose_checkpoint(model, imgsz, path):
    """A pose checkpoint the way the reference's trainer writes it (pose/train.py:49 sets model.kpt_shape)."""
    spec = arch.build_spec(model)
    m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    m.kpt_shape = list(m.model[-1].kpt_shape)
    m.args = {"task": "pose", "imgsz": imgsz}
    torch.save({"model": m, "ema": None, "train_args": {"task": "pose", "imgsz": imgsz}}, path)
    return path


def pose():
    """the Pose head through the reference's own PoseModel / PosePredictor."""
    ref_import.import_reference()
    from ultralytics import YOLO

    for case, (model, *_rest) in kat.POSE_NET_CASES.items():
        spec = arch.build_spec(model)
        sd = arch.synth_state_dict(spec, 0)
        m = ref_import.reference_model(model, sd)
        with torch.no_grad():
            y, (raws, kpt) = m(torch.from_numpy(kat.net_input(case)))
        np.savez_compressed(os.path.join(OUT, f"net_{case}.npz"), y=y.numpy(), kpt=kpt.numpy(), raw0=raws[0].numpy(),
                            raw1=raws[1].numpy(), raw2=raws[2].numpy(), sd_checksum=np.float64(kat.state_dict_checksum(sd)))
        print("net", case, tuple(y.shape), tuple(kpt.shape))
    for case, (model, n, h, w, imgsz, seed) in kat.POSE_PREDICT_CASES.items():
        yolo = YOLO(pose_checkpoint(model, imgsz, f"/tmp/b2y_golden_{case}.pt"))
        res = yolo.predict(kat.predict_input(case), device="cpu", imgsz=imgsz, conf=0.25, iou=0.7, max_det=300, verbose=False)
        assert type(yolo.predictor).__name__ == "PosePredictor"
        d = {f"boxes{i}": r.boxes.data.numpy().astype(np.float32) for i, r in enumerate(res)}
        d.update({f"kpts{i}": r.keypoints.data.numpy().astype(np.float32) for i, r in enumerate(res)})
        d["n"] = np.array([r.boxes.data.shape[0] for r in res], np.int64)
        np.savez_compressed(os.path.join(OUT, f"predict_{case}.npz"), **d)
        print("predict", case, d["n"].tolist(), [d[f"kpts{i}"].shape for i in range(n)])


def segment():
    """the Segment head through the reference's own SegmentationModel / SegmentationPredictor."""
    ref_import.import_reference()
    from ultralytics import YOLO

    for case, (model, *_rest) in kat.SEG_NET_CASES.items():
        spec = arch.build_spec(model)
        sd = arch.synth_state_dict(spec, 0)
        m = ref_import.reference_model(model, sd)
        with torch.no_grad():
            y, (raws, mc, p) = m(torch.from_numpy(kat.net_input(case)))
        np.savez_compressed(os.path.join(OUT, f"net_{case}.npz"), y=y.numpy(), mc=mc.numpy(), proto=p.numpy(),
                            sd_checksum=np.float64(kat.state_dict_checksum(sd)))
        print("net", case, tuple(y.shape), tuple(mc.shape), tuple(p.shape))
    for case, (model, n, h, w, imgsz, seed) in kat.SEG_PREDICT_CASES.items():
        spec = arch.build_spec(model)
        m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
        m.names = {k: f"class{k}" for k in range(spec["nc"])}
        m.args = {"task": "segment", "imgsz": imgsz}
        path = f"/tmp/b2y_golden_{case}.pt"
        torch.save({"model": m, "ema": None, "train_args": {"task": "segment", "imgsz": imgsz}}, path)
        yolo = YOLO(path)
        res = yolo.predict(kat.predict_input(case), device="cpu", imgsz=imgsz, conf=0.25, iou=0.7, max_det=300, verbose=False)
        assert type(yolo.predictor).__name__ == "SegmentationPredictor"
        d = {f"boxes{i}": r.boxes.data.numpy().astype(np.float32) for i, r in enumerate(res)}
        for i, r in enumerate(res):
            mk = r.masks.data.numpy()
            assert set(np.unique(mk)) <= {0.0, 1.0}
            d[f"masks{i}"] = np.packbits(mk.astype(np.uint8), axis=None)
            d[f"mshape{i}"] = np.array(mk.shape, np.int64)
        d["n"] = np.array([r.boxes.data.shape[0] for r in res], np.int64)
        np.savez_compressed(os.path.join(OUT, f"predict_{case}.npz"), **d)
        print("predict", case, d["n"].tolist(), [d[f"mshape{i}"].tolist() for i in range(n)])


def val_match():
    """validator: IoU matrix and true-positive matrix of the reference's own validator methods."""
    ref_import.import_reference()
    from ultralytics.models.yolo.detect.val import DetectionValidator
    from ultralytics.utils.metrics import box_iou

    v = object.__new__(DetectionValidator)  # the two methods only read self.iouv (val.py:41)
    v.iouv = torch.linspace(0.5, 0.95, 10)
    vm = {"iouv": v.iouv.numpy()}
    for case in kat.VAL_MATCH_CASES:
        det, gt, gcls = (torch.from_numpy(a) for a in kat.val_match_case(case))
        vm[f"{case}__iou"] = box_iou(gt, det[:, :4]).numpy().astype(np.float32)
        vm[f"{case}__tp"] = v._process_batch(det, gt, gcls).numpy()
        vm[f"{case}__insum"] = np.float64(det.double().sum() + gt.double().sum() + gcls.double().sum())
        print("val_match", case, vm[f"{case}__tp"].sum(0).tolist())
    np.savez_compressed(os.path.join(OUT, "val_match.npz"), **vm)


if __name__ == "__main__":
    # `make_golden.py val_match` / `pose` / `segment` regenerate those files only
    {"val_match": val_match, "pose": pose, "segment": segment}.get((sys.argv[1:] or [""])[0], main)()

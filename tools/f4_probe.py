"""SURVEY §8 f4 measurement: the Pose and Segment heads through the drop-in predictors on one B200.

    python tools/f4_probe.py [--frames 64] [--out gpurun_out/f4_probe.json]

Per head (stock yolo11n-pose / yolo11n-seg graphs, seeded synthetic checkpoints, 640x640 uint8 frames):
  * end to end: ``YOLO(ckpt).predict(list of frames, predictor=B200PosePredictor | B200SegmentationPredictor)`` incl. host
    staging, H2D, ``Results`` construction (wall clock, mean of 3 calls after a warm-up call that builds the engine);
  * device-timed stages on the predictor's own engine (CUDA events, 20 iterations after 3 warm-ups): the CUDA-graphed
    network, and the post-process (NMS + kept-row gather + keypoint decode | mask assembly);
  * the mask kernel alone against the HBM roof (it writes rows x H x W fp32 and reads almost nothing);
  * the reference's own predictor on the host cores (``device="cpu"``, 8 frames, after a warm-up call), LAST because the
    reference's select_device("cpu") hides the GPU from later code in the same process.
The reference package must be importable (baseline/_ref)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bs_yolo_b200 import arch, ops  # noqa: E402
from oracle import ref_import  # noqa: E402


def checkpoint(model, task, path):
    spec = arch.build_spec(model)
    m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    if task == "pose":
        m.kpt_shape = list(m.model[-1].kpt_shape)
    m.args = {"task": task, "imgsz": 640}
    torch.save({"model": m, "ema": None, "train_args": {"task": task, "imgsz": 640}}, path)
    return path


def dev_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "f4_probe.json"))
    args = ap.parse_args()
    ref_import.import_reference()
    from ultralytics import YOLO

    from bs_yolo_b200.predictor import B200PosePredictor, B200SegmentationPredictor

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    rng = np.random.default_rng(0)
    n = args.frames
    frames = [rng.integers(0, 256, (640, 640, 3), dtype=np.uint8) for _ in range(n)]
    out = {"frames": n, "imgsz": 640, "heads": {}}
    ckpts = {}
    for head, model, task, cls in (("pose", "yolo11n-pose", "pose", B200PosePredictor),
                                   ("segment", "yolo11n-seg", "segment", B200SegmentationPredictor)):
        ckpts[head] = checkpoint(model, task, f"/tmp/b2y_f4_{head}.pt")
        yolo = YOLO(ckpts[head])
        kw = dict(predictor=cls, device=0, imgsz=640, conf=0.25, iou=0.7, max_det=300, verbose=False)
        t0 = time.perf_counter()
        res = yolo.predict(frames, **kw)  # builds, autotunes and graph-captures the (n, 640, 640) engine
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(3):
            res = yolo.predict(frames, **kw)
        torch.cuda.synchronize()
        e2e = 3 * n / (time.perf_counter() - t0)
        eng = next(iter(yolo.predictor.model._engines.values()))
        rec = {"model": model, "predictor": cls.__name__, "first_call_s": round(build_s, 2), "e2e_images_per_s": round(e2e, 1),
               "detections_per_image": round(float(np.mean([len(r.boxes) for r in res])), 2), "plan_ops": len(eng.plan.ops),
               "plan_launches": eng.n_launches}
        with torch.cuda.device(eng.device):
            rec["network_graph_ms"] = round(dev_time(eng.forward), 4)
            dets, counts, kept = eng.nms(0.25, 0.7, 300, return_anchors=True)
            nl = counts.tolist()
            rec["nms_ms"] = round(dev_time(lambda: eng.nms(0.25, 0.7, 300, return_anchors=True)), 4)
            if head == "pose":
                geom = ops.coords_geometry((640, 640), [(640, 640)] * n, dets.device)
                rec["keypoints_kernel_ms"] = round(dev_time(lambda: eng.keypoints(kept, counts, geom=geom)), 4)
                rec["extra_bytes_read_if_decoded_for_all_anchors"] = n * 8400 * 51 * 4 * 3
            else:
                coeffs = eng.mask_coeffs(kept, counts)
                rec["coeff_gather_ms"] = round(dev_time(lambda: eng.mask_coeffs(kept, counts)), 4)
                ms = dev_time(lambda: ops.process_mask_batched(eng.protos, coeffs, dets, nl, (640, 640)))
                rows = int(sum(nl))
                nbytes = rows * 640 * 640 * 4
                rec.update(process_mask_ms=round(ms, 4), mask_rows=rows, mask_bytes_written=nbytes,
                           process_mask_gbps=round(nbytes / ms / 1e6, 1))
                if peaks:
                    hbm = float(peaks.get("hbm_gbs") or 0) or None  # measured copy bandwidth, GB/s
                    if hbm:
                        rec["process_mask_frac_of_hbm"] = round(nbytes / ms / 1e6 / hbm, 3)
        rec["images_per_s_device"] = round(n / (rec["network_graph_ms"] + rec["nms_ms"] + rec.get("keypoints_kernel_ms", 0)
                                                + rec.get("coeff_gather_ms", 0) + rec.get("process_mask_ms", 0)) * 1e3, 1)
        out["heads"][head] = rec
        print(head, json.dumps(rec), file=sys.stderr)
        del yolo, eng, res
        torch.cuda.empty_cache()
    # the reference's own CPU predictors, last
    torch.set_num_threads(os.cpu_count() or 1)
    for head in ("pose", "segment"):
        yolo = YOLO(ckpts[head])
        kw = dict(device="cpu", imgsz=640, conf=0.25, iou=0.7, max_det=300, verbose=False)
        yolo.predict(frames[:1], **kw)
        t0 = time.perf_counter()
        yolo.predict(frames[:8], **kw)
        dt = time.perf_counter() - t0
        out["heads"][head]["reference_cpu_images_per_s"] = round(8 / dt, 2)
        out["heads"][head]["reference_cpu_sample"] = f"8 frames, one predict() call, {type(yolo.predictor).__name__}, {torch.get_num_threads()} threads"
        out["heads"][head]["e2e_over_reference"] = round(out["heads"][head]["e2e_images_per_s"] / (8 / dt), 1)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

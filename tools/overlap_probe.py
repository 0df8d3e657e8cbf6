"""How much of the pre-/post-processing leaks into the step: time K pipelined steps with parts of the side work removed.
usage: python tools/overlap_probe.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from bs_yolo_b200.ops import nms_raw  # noqa: E402
from bs_yolo_b200.pipeline import DetectPipeline  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
model, batch, imgsz = bench.MODEL_CFG["yolo11n"]
pipe = DetectPipeline(bench.make_ir(model), batch, imgsz, (imgsz, imgsz))
frames = [torch.from_numpy(bench.synth_frames(batch, imgsz, 7 + i)).cuda() for i in range(4)]
main = torch.cuda.current_stream()


def run(fn, name):
    for j in range(5):
        fn(j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(K):
        fn(j)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / K:.4f} ms/step", flush=True)


def full(j):
    pipe.submit(frames[j % 4], src_ready=True)


def fwd_only(j):
    pipe.engines[j & 1].forward()


def pre_fwd(j):
    k = j & 1
    eng = pipe.engines[k]
    with torch.cuda.stream(pipe._pre_stream):
        pipe._pre_stream.wait_event(pipe._fwd_done[k])
        eng.preprocess(frames[j % 4], auto=pipe.auto)
        pipe._pre_done[k].record(pipe._pre_stream)
    main.wait_event(pipe._pre_done[k])
    eng.forward()
    pipe._fwd_done[k].record(main)


def fwd_nms(j):
    k = j & 1
    eng = pipe.engines[k]
    main.wait_event(pipe._post_done[k])
    eng.forward()
    pipe._fwd_done[k].record(main)
    with torch.cuda.stream(pipe._post_stream):
        pipe._post_stream.wait_event(pipe._fwd_done[k])
        nms_raw(eng.y, pipe.conf, pipe.iou, pipe.max_det, pipe.classes, pipe.agnostic)
        pipe._post_done[k].record(pipe._post_stream)


def serial(j):
    eng = pipe.engines[0]
    eng.preprocess(frames[j % 4], auto=pipe.auto)
    eng.forward()
    nms_raw(eng.y, pipe.conf, pipe.iou, pipe.max_det, pipe.classes, pipe.agnostic)


full(0)
torch.cuda.synchronize()
run(fwd_only, "network graph only (alternating slots)")
run(pre_fwd, "preprocess (side stream) + network")
run(fwd_nms, "network + NMS (side stream)")
run(full, "full pipelined step")
run(serial, "one stream, no overlap")

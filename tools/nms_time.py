"""Time b2y_nms on the decoded output of the bench workload; prints candidate statistics.  Usage: nms_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bs_yolo_b200.pipeline import DetectPipeline
from bs_yolo_b200.ops import nms_raw
pipe = DetectPipeline(bench.make_ir("yolo11n"), 64, 640, (640, 640))
frames = torch.from_numpy(bench.synth_frames(64, 640, 100)).cuda()
pipe(frames)
y = pipe.engine.y.clone()
cand = (y[:, 4:].amax(1) > 0.25).sum(1)
print("candidates per image: min %d median %d max %d" % (int(cand.min()), int(cand.median()), int(cand.max())))
for _ in range(3):
    nms_raw(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    d, c = nms_raw(y)
e1.record()
torch.cuda.synchronize()
print("nms ms/call %.4f  kept median %d" % (e0.elapsed_time(e1) / 20, int(c.median())))

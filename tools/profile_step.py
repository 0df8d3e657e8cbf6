"""One eager (non-graph) pass of the whole hot path at the bench workload, for ncu captures:
preprocess -> every network op -> decode -> NMS.  Usage: python tools/profile_step.py [workload] [batch] [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from bs_yolo_b200.pipeline import DetectPipeline  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "yolo11n"
model, batch, imgsz = bench.MODEL_CFG[workload]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else batch
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pipe = DetectPipeline(bench.make_ir(model), batch, imgsz, (imgsz, imgsz), use_graph=False)
frames = torch.from_numpy(bench.synth_frames(batch, imgsz, 7)).cuda()
for _ in range(passes):
    dets, counts = pipe(frames)
torch.cuda.synchronize()
print("ok", counts[:4].tolist(), pipe.launches_per_batch, "launches per pass")

"""One eager (non-graph) pass of the whole hot path at the bench workload, for ncu captures:
preprocess -> every network op (Detect decode fused) -> NMS.

    python tools/profile_step.py [workload] [batch] [passes]

Warm-up (per-layer autotuning + one full pass) runs outside the profiled range; the measured passes are bracketed by
cudaProfilerStart/Stop, so `ncu --profile-from-start off ...` sees exactly `launches per pass` x passes kernels, in
plan order (the op table printed at the end maps launch index -> layer)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from bs_yolo_b200 import _cabi as A  # noqa: E402
from bs_yolo_b200.pipeline import DetectPipeline  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "yolo11n"
model, batch, imgsz = bench.MODEL_CFG[workload][:3]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else batch
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pipe = DetectPipeline(bench.make_ir(model), batch, imgsz, (imgsz, imgsz), use_graph=False)
frames = torch.from_numpy(bench.synth_frames(batch, imgsz, 7)).cuda()
for _ in range(2):  # autotune + warm-up, not profiled
    pipe(frames)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
for _ in range(passes):
    dets, counts = pipe(frames)
torch.cuda.synchronize()
rt.cudaProfilerStop()
table = [{"launch": 0, "name": "preprocess", "kind": "preprocess"}]
for i, o in enumerate(pipe.engine.plan.ops):
    table.append({"launch": i + 1, "name": o.name, "kind": A.OP_NAMES[o.kind], "impl": pipe.engine.op_impl(i),
                  "bytes": pipe.engine.plan.op_bytes(o, batch), "flops": o.flops * batch})
n = len(table)
table += [{"launch": n, "name": "nms.score", "kind": "nms"}, {"launch": n + 1, "name": "nms.select", "kind": "nms"}]
out = os.environ.get("B2Y_PROFILE_TABLE")
if out:
    with open(out, "w") as f:
        json.dump({"workload": workload, "batch": batch, "launches_per_pass": len(table), "ops": table}, f, indent=1)
print("ok", counts[:4].tolist(), pipe.launches_per_batch, "launches per pass")

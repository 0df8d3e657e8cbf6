"""In-graph time of op ranges of the bench workload (torch CUDA-graph capture around Engine.run_ops)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bs_yolo_b200.pipeline import DetectPipeline
pipe = DetectPipeline(bench.make_ir("yolo11n"), 64, 640, (640, 640))
eng = pipe.engine
frames = torch.from_numpy(bench.synth_frames(64, 640, 100)).cuda()
pipe(frames); pipe(frames)
torch.cuda.synchronize()
names = [o.name for o in eng.plan.ops]
def idx(prefix):
    return [i for i, n in enumerate(names) if n.startswith(prefix)]
def first(layer):  # first op of the first top-level layer >= `layer` that still has ops (folded layers have none)
    return min(i for i, n in enumerate(names) if int(n.split(".")[1]) >= layer)
regions = {"all": (0, len(names)), "stem..model.4": (0, first(5)), "model.5..10": (first(5), first(11)),
           "neck 11..22": (first(11), first(23)), "head": (first(23), len(names))}
# finer: each top-level layer
tops = {}
for i, n in enumerate(names):
    k = ".".join(n.split(".")[:2])
    tops.setdefault(k, [i, i + 1])[1] = i + 1
def time_range(a, b, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.run_ops(a, b)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            eng.run_ops(a, b)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for k, (a, b) in regions.items():
    print(f"{k:16s} ops {a:3d}-{b:3d} ({b-a:2d}): {time_range(a, b):8.1f} us")
for k, (a, b) in tops.items():
    print(f"  {k:12s} ops {a:3d}-{b:3d} ({b-a:2d}): {time_range(a, b):8.1f} us   {(time_range(a,b))/(b-a):6.1f} us/op")

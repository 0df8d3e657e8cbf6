"""Debug probe (1 GPU): where does a SAHI step spend its time?  Resident vs end-to-end variants, CUDA-event phase split."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bs_yolo_b200.pipeline import DetectPipeline
from bs_yolo_b200.shard import SahiSharder, merge_prediction
from bs_yolo_b200.ops import nms_raw

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
sh = SahiSharder(lambda B: DetectPipeline(bench.make_ir("bsyolo_n"), B, 640, (640, 640), device=0), (2160, 3840), 640, frames=F)
host = [torch.from_numpy(bench.synth_frames(F, (2160, 3840), 5 + j)).pin_memory() for j in range(2)]
devf = [h.to(dev) for h in host]
for j in range(3):
    sh.predict_frames(devf[j % 2])
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record()
tiles = sh.slice_tiles(devf[0]).view(sh.B, 640, 640, 3)
ev[1].record()
dets, counts = sh.pipe(tiles)
ev[2].record()
valid, org, order = sh._consts(dev)
d2 = dets.view(1, F, sh.per, 300, 6).transpose(0, 1).reshape(F, sh.per, 300, 6).index_select(1, order)
c2 = counts.view(1, F, sh.per).transpose(0, 1).reshape(F, sh.per).index_select(1, order)
pred = merge_prediction(d2, c2, org, sh.nc)
ev[3].record()
md, mc = nms_raw(pred, sh.conf, sh.iou, sh.max_det)
ev[4].record()
torch.cuda.synchronize()
print("phases ms: slice %.3f pipe %.3f merge %.3f nms %.3f" % tuple(ev[i].elapsed_time(ev[i + 1]) for i in range(4)))
K = 10
torch.cuda.synchronize(); t0 = time.perf_counter()
for j in range(K):
    sh.predict_frames(devf[j % 2])
torch.cuda.synchronize(); print("resident ms/step %.3f" % ((time.perf_counter() - t0) * 1e3 / K))
stage = [torch.empty_like(devf[0]) for _ in range(2)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for j in range(K):
    stage[0].copy_(host[j % 2], non_blocking=True)
torch.cuda.synchronize(); print("h2d only ms/step %.3f (%.1f GB/s)" % ((time.perf_counter() - t0) * 1e3 / K, host[0].numel() * K / (time.perf_counter() - t0) / 1e9))
torch.cuda.synchronize(); t0 = time.perf_counter()
for j in range(K):
    stage[0].copy_(host[j % 2], non_blocking=True)
    sh.predict_frames(stage[0])
torch.cuda.synchronize(); print("e2e serial ms/step %.3f" % ((time.perf_counter() - t0) * 1e3 / K))
cs = torch.cuda.Stream(dev); main = torch.cuda.current_stream(dev)
up, used = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
def upload(j):
    with torch.cuda.stream(cs):
        cs.wait_event(used[j & 1]); stage[j & 1].copy_(host[j % 2], non_blocking=True); up[j & 1].record(cs)
torch.cuda.synchronize(); t0 = time.perf_counter()
for e in used: e.record(main)
upload(0)
for j in range(K):
    if j + 1 < K: upload(j + 1)
    main.wait_event(up[j & 1]); sh.predict_frames(stage[j & 1]); used[j & 1].record(main)
torch.cuda.synchronize(); print("e2e overlapped ms/step %.3f" % ((time.perf_counter() - t0) * 1e3 / K))
t0 = time.perf_counter()
for j in range(K):
    sh.predict_frames(devf[j % 2])
print("host enqueue ms/step %.3f" % ((time.perf_counter() - t0) * 1e3 / K)); torch.cuda.synchronize()
# ---- diagnosis of the overlapped variant
import gc
print("reserved GB before", torch.cuda.memory_reserved() / 1e9)
for K2 in (10, 30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for e in used: e.record(main)
    upload(0)
    for j in range(K2):
        if j + 1 < K2: upload(j + 1)
        main.wait_event(up[j & 1]); sh.predict_frames(stage[j & 1]); used[j & 1].record(main)
    torch.cuda.synchronize(); print("overlapped K=%d ms/step %.3f reserved %.1f GB" % (K2, (time.perf_counter() - t0) * 1e3 / K2, torch.cuda.memory_reserved() / 1e9))
# sync every step: the upload of step j+1 still overlaps the compute of step j
torch.cuda.synchronize()
for e in used: e.record(main)
upload(0)
ts = []
for j in range(12):
    t0 = time.perf_counter()
    if j + 1 < 12: upload(j + 1)
    main.wait_event(up[j & 1]); sh.predict_frames(stage[j & 1]); used[j & 1].record(main)
    main.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("overlapped, main.synchronize() per step:", " ".join("%.2f" % t for t in ts))
# upload on the copy stream but WITHOUT overlap (wait for it right away)
torch.cuda.synchronize(); t0 = time.perf_counter()
for j in range(10):
    with torch.cuda.stream(cs):
        stage[0].copy_(host[j % 2], non_blocking=True); up[0].record(cs)
    main.wait_event(up[0]); sh.predict_frames(stage[0]); torch.cuda.synchronize()
print("copy stream, no overlap ms/step %.3f" % ((time.perf_counter() - t0) * 1e3 / 10))

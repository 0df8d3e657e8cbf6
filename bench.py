#!/usr/bin/env python
"""bench.py -- images/sec @640x640 of the detection hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference arm: CPU port on the host cores)

A *step* is one pass of the whole hot path over one batch of synthetic frames per GPU:
uint8 BGR frames -> fused letterbox/normalise -> backbone+neck+Detect -> DFL/dist2bbox decode -> NMS -> detections.
Workload at N=1: BASELINE.json configs[1] (stock yolo11n, batch 64, bf16, one B200); image batches shard data-parallel
(weak scaling: every rank runs its own batch, only the final detections are gathered -- no collective on the layer path).

`value`  : whole-job images/s, frames already resident in HBM (device-timed, max over ranks).
`e2e`    : same metric through DetectPipeline.predict_stream: pinned host frames -> H2D -> pipeline -> D2H of detections,
            every step's copies inside the timed region (wall clock around K steps), copies overlapped with compute.
`roofline`: the dominant kernel (the tcgen05 implicit-GEMM conv, summed over its launches of one step): algorithmic
            bytes / CUDA-event time against the measured HBM peak (MEASURED_PEAKS.json).
`cpu_baseline`: the oracle port (torch fp32 CPU restatement of the reference path) timed on the host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

MODEL_CFG = {  # workload name -> (arch model, per-GPU batch, imgsz)
    "yolo11n": ("yolo11n", 64, 640),
    "bsyolo_n": ("bsyolo_n", 256, 640),
    "yolo11x": ("yolo11x", 32, 1280),
}
METRIC, UNIT = "images/sec @640x640", "images/s"  # BASELINE.json metric; the workload (model, batch) is named in config


def make_ir(model):
    from bs_yolo_b200 import arch, ir
    spec = arch.build_spec(model)
    return ir.build_ir(spec, arch.synth_state_dict(spec, 0))


def synth_frames(n, hw, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, hw, hw, 3), dtype=np.uint8)


# ----------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU port (oracle) timing
def cpu_port_throughput(model, imgsz, n_images, batch, threads, steps=1, warmup=0):
    """Times the ORACLE port of the reference path (preprocess + fp32 forward + decode + NMS) on the host cores."""
    from oracle import oracle_net, oracle_post
    torch.set_num_threads(threads)
    g = make_ir(model)
    frames = synth_frames(n_images, imgsz, 1234)

    def one_pass():
        for i in range(0, n_images, batch):
            chunk = list(frames[i:i + batch])
            x = torch.from_numpy(oracle_post.preprocess(chunk, (imgsz, imgsz)))
            y, _ = oracle_net.forward(g, x)
            oracle_post.non_max_suppression(y.numpy(), 0.25, 0.7, max_det=300)

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    return steps * n_images / dt, dt / steps


def cpu_reference_throughput(model, imgsz, n_images, batch, threads, steps=1, warmup=1):
    """Times the UNMODIFIED reference (baseline/_ref, pip-installed from /root/reference by __graft_entry__.build()):
    YOLO(<synthetic checkpoint>).predict(list of ndarrays, device="cpu") -- LetterBox + preprocess + fp32 torch forward
    + NMS + Results, through its own public API -- on `threads` host threads.  Returns None if it is not importable."""
    from oracle import ref_import
    if not ref_import.available():
        return None
    import tempfile
    from bs_yolo_b200 import arch
    ref_import.import_reference()
    from ultralytics import YOLO
    spec = arch.build_spec(model)
    m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    m.args = {"task": "detect", "imgsz": imgsz}
    path = os.path.join(tempfile.mkdtemp(prefix="b2y_ref_"), f"{model}.pt")
    torch.save({"model": m, "ema": None, "train_args": {"task": "detect", "imgsz": imgsz}}, path)
    yolo = YOLO(path)
    frames = synth_frames(n_images, imgsz, 1234)
    kw = dict(device="cpu", imgsz=imgsz, conf=0.25, iou=0.7, max_det=300, verbose=False)
    yolo.predict([frames[0]], **kw)   # first call: settings / predictor setup (~9 s in a fresh env)
    torch.set_num_threads(threads)    # the reference caps itself at min(8, ncpu-1) threads (torch_utils.py:229); use them all

    def one_pass():
        for i in range(0, n_images, batch):
            yolo.predict(list(frames[i:i + batch]), **kw)

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    return steps * n_images / dt, dt / steps


def cpu_arm(model, imgsz, n_img, cb, cores, steps, warmup):
    """(value, step seconds, kind): the reference itself when importable, else the oracle port."""
    r = None
    try:
        r = cpu_reference_throughput(model, imgsz, n_img, cb, cores, steps=steps, warmup=warmup)
    except Exception as e:  # noqa: BLE001 -- a broken reference install must not take the bench down
        sys.stderr.write(f"[bench] reference package unusable ({type(e).__name__}: {e}); timing the oracle port instead\n")
    if r is not None:
        return r[0], r[1], "reference"
    v, s_ = cpu_port_throughput(model, imgsz, n_img, cb, cores, steps=steps, warmup=warmup)
    return v, s_, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model, batch, imgsz = MODEL_CFG[args.workload]
    cores = os.cpu_count() or 1
    n_img, cb = 32, 16
    val, step_s, kind = cpu_arm(model, imgsz, n_img, cb, cores, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{model} detect forward + DFL decode + NMS, batch {batch} @ {imgsz}x{imgsz}",
                   "note": ("the unmodified reference (baseline/_ref): YOLO(ckpt).predict(list, device='cpu')" if kind == "reference"
                            else "CPU port (oracle) of the reference torch path") + "; each step = bounded sample of the workload"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n_img} synthetic {imgsz}x{imgsz} frames per step in batches of {cb}, torch fp32, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- ours
def per_op_times(eng, reps=3):
    """CUDA-event duration of every op of the plan (eager launches on the current stream), averaged over reps."""
    n = eng.n_launches
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(reps)]
    for r in range(reps):
        ev[r][0].record()
        for i in range(n):
            eng.run_ops(i, i + 1)
            ev[r][i + 1].record()
    torch.cuda.synchronize()
    return [float(np.mean([ev[r][i].elapsed_time(ev[r][i + 1]) for r in range(reps)])) for i in range(n)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="yolo11n", choices=list(MODEL_CFG))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers-out", default="", help="write the per-op timing table to this JSON file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    from bs_yolo_b200 import _cabi as A
    from bs_yolo_b200.pipeline import DetectPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model, batch, imgsz = MODEL_CFG[args.workload]
    batch = args.batch or batch
    K, Wm = args.steps, args.warmup

    pipe = DetectPipeline(make_ir(model), batch, imgsz, (imgsz, imgsz), device=local)
    eng = pipe.engine
    n_rot = 4  # rotate over 4 distinct input batches (4 x 78.6 MB > 126 MB L2); activations add ~GBs of traffic per step
    host = [torch.from_numpy(synth_frames(batch, imgsz, 100 + rank * 16 + j)).pin_memory() for j in range(n_rot)]
    devf = [h.to(dev) for h in host]
    def step_resident(j):
        # pipelined submit: the NMS (and, multi-GPU, the all-gather of the padded detections -- the only data that crosses
        # GPUs) of this batch runs on the post-process stream underneath the next batch's network
        return pipe.submit(devf[j % n_rot], gather=world > 1, src_ready=True)  # frames are resident before the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prof_range = os.environ.get("B2Y_CUDA_PROFILER_RANGE") == "1"  # `ncu --profile-from-start off`: only the timed steps

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if prof_range:
            torch.cuda.cudart().cudaProfilerStart()
        t0 = time.perf_counter()
        e0.record()
        for j in range(steps):
            fn(j)
        if finish is not None:
            finish()  # e.g. DetectPipeline.drain: the side-stream work of the last batches belongs to the timed region
        e1.record()
        barrier()
        if prof_range:
            torch.cuda.cudart().cudaProfilerStop()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall

    for j in range(Wm):
        step_resident(j)
    pipe.drain()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall = timed(step_resident, K, pipe.drain)
    value = world * batch * K / (ms * 1e-3)

    # ---- end to end through the public streaming API with host buffers: every step copies its frames host->device
    #      (pinned memory) and its detections device->host inside the timed region; copies overlap the previous compute
    def run_e2e(steps):
        n = 0
        for dets_h, counts_h in pipe.predict_stream((host[j % n_rot] for j in range(steps)), gather=world > 1):
            n += int(counts_h.shape[0])
        return n

    run_e2e(3)
    barrier()
    t0 = time.perf_counter()
    n_out = run_e2e(K)
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    assert n_out == K * world * batch
    if world > 1:
        t = torch.tensor([wall_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_e2e = float(t[0])
    ms_e2e = wall_e2e
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * batch * K / (max(ms_e2e, wall_e2e) * 1e-3)

    # ---- stage split of one resident step (rank 0): preprocess | CUDA-graphed network (+ fused decode) | NMS
    stages = None
    if rank == 0:
        from bs_yolo_b200.ops import nms_raw
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
        for j in range(K):
            evs[j][0].record()
            eng.preprocess(devf[j % n_rot], auto=pipe.auto)
            evs[j][1].record()
            eng.forward()
            evs[j][2].record()
            nms_raw(eng.y, pipe.conf, pipe.iou, pipe.max_det)
            evs[j][3].record()
        torch.cuda.synchronize()
        stages = {n: float(np.median([evs[j][i].elapsed_time(evs[j][i + 1]) for j in range(K)]))
                  for i, n in enumerate(("preprocess_ms", "network_graph_ms", "nms_ms"))}

    # ---- per-op timing -> roofline of the dominant kernel (rank 0)
    roof, layers = None, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        eng.forward()
        times = per_op_times(eng, reps=3)
        layers = []
        for i, o in enumerate(eng.plan.ops):
            layers.append({"i": i, "name": o.name, "kind": A.OP_NAMES[o.kind], "impl": eng.op_impl(i), "ms": times[i],
                           "bytes": o.bytes_alg * batch, "flops": o.flops * batch})
        dom_kind = "conv"
        dom = [L for L in layers if L["kind"] == dom_kind and L["impl"] == A.IMPL_TCGEN05] or [L for L in layers if L["kind"] == dom_kind]
        b_dom = sum(L["bytes"] for L in dom)
        f_dom = sum(L["flops"] for L in dom)
        # duration of the kernel's launches inside one step: the CUDA-graphed network time (events, median over K steps)
        # times the kernel's share of the eager per-op times (an eager event pair adds launch overhead to every op, the
        # share is insensitive to it)
        share = sum(L["ms"] for L in dom) / max(sum(times), 1e-9)
        t_dom = stages["network_graph_ms"] * 1e-3 * share
        ach = b_dom / t_dom / 1e9
        traffic = None
        try:  # DRAM bytes of the same kernel's launches of one step, from the committed ncu launch list (profiles/)
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr.get("workload") == args.workload and tr.get("batch") == batch:
                traffic = tr["conv_tc_dram_bytes_per_step"] / max(tr["conv_tc_launches_per_step"], 1)
        except Exception:
            pass
        # which roof bounds the kernel on this workload: its launches' summed per-layer floors (yolo11n / BS-YOLO: HBM;
        # yolo11x@1280: the tensor pipe -- sustained cuBLAS figure, the kernel runs inside a long step)
        tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        t_hbm = sum(L["bytes"] for L in dom) / (hbm * 1e9)
        t_tc = sum(L["flops"] for L in dom) / (tf_peak * 1e12)
        tensor_bound = t_tc > t_hbm
        tfl = f_dom / t_dom / 1e12
        roof = {"bound": "tensor" if tensor_bound else "hbm", "achieved": tfl if tensor_bound else ach,
                "peak": tf_peak if tensor_bound else hbm, "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": (tfl / tf_peak) if tensor_bound else ach / hbm, "traffic": traffic,
                "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv), all its launches of one step" if dom and dom[0]["impl"] == 2 else "conv_simt_kernel",
                "launches_per_step": len(dom), "avg_launch_ms": t_dom * 1e3 / max(len(dom), 1),
                "algorithmic_bytes_per_launch": b_dom / max(len(dom), 1), "tflops": f_dom / t_dom / 1e12,
                "share_of_step": share * stages["network_graph_ms"] / (ms / K), "peak_source": which,
                "net_bytes_per_image": eng.plan.work_per_image()[1], "net_flops_per_image": eng.plan.work_per_image()[0],
                "network_roofline_frac": (sum(max(L["bytes"] / (hbm * 1e9), L["flops"] / (float(peaks.get("bf16_tflops", 1590.0)) * 1e12))
                                              for L in layers) * 1e3) / stages["network_graph_ms"]}
        if args.layers_out:
            with open(args.layers_out, "w") as f:
                json.dump({"batch": batch, "sum_ms": sum(times), "layers": layers}, f, indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = 512 if imgsz <= 640 else 32  # ~10 s of CPU work on the box's host cores
        v, s_, kind = cpu_arm(model, imgsz, n_cpu, 16, cores, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_cpu} synthetic {imgsz}x{imgsz} frames (batches of 16) after a warm-up call, "
                         + ("reference YOLO.predict(device='cpu')" if kind == "reference" else "oracle torch-fp32 port")
                         + f", {cores} threads, {s_:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{model} detect forward + DFL decode + NMS, batch {batch} per GPU @ {imgsz}x{imgsz} "
                                   "(BASELINE.json configs[1])", "parallelism": f"dp{world} (images sharded, detections gathered)",
                       "cache": f"inputs rotate over {n_rot} batches ({n_rot * batch * imgsz * imgsz * 3 / 1e6:.0f} MB > 126 MB L2); "
                                "each step streams GBs of activations", "weights": "seeded synthetic checkpoint (bs_yolo_b200.arch, seed 0)",
                       "conf": 0.25, "iou": 0.7, "max_det": 300},
            # whole-job bytes per step: every rank uploads its shard and reads back the gathered detections of all ranks
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * batch * imgsz * imgsz * 3,
                    "d2h_bytes_per_step": world * world * batch * (pipe.max_det * 6 * 4 + 4), "ms_per_step": max(ms_e2e, wall_e2e) / K},
            "gpu_launches": world * pipe.launches_per_batch * K, "launches_per_step": pipe.launches_per_batch,
            "clocks": clocks, "stages": stages, "roofline": roof, "cpu_baseline": cpu, "wall_ms_per_step": wall / K,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- images/sec @640x640 of the detection hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference arm: its own CPU predict on the host cores)

A *step* is one pass of the whole hot path over one batch of synthetic frames per GPU:
uint8 BGR frames -> fused letterbox/normalise -> backbone+neck+Detect -> DFL/dist2bbox decode -> NMS -> detections.
Headline workload: BASELINE.json configs[1] (stock yolo11n, batch 64 per GPU, bf16); image batches shard data-parallel
(weak scaling: every rank runs its own batch, only the final detections are gathered -- no collective on the layer path).

`value`  : whole-job images/s, frames already resident in HBM (device-timed, max over ranks).
`e2e`    : same metric through DetectPipeline.predict_stream: pinned host frames -> H2D -> pipeline -> D2H of detections,
            every step's copies inside the timed region (wall clock around K steps), copies overlapped with compute.
`roofline`: the dominant kernel (the tcgen05 implicit-GEMM conv, summed over its launches of one step): algorithmic
            bytes (activations x batch + weights once) / CUDA-event time against the measured HBM peak.
`cpu_baseline`: the unmodified reference (baseline/_ref) -- else the oracle port -- timed on the host cores, rank 0, N=1.
`workloads`: the other BASELINE.json configs as sub-records of the same line, each with its own value / e2e / roofline:
            configs[2] BS-YOLO-n ("best.pt" architecture) batch 256 sharded 256/N (strong scaling), configs[3] yolo11x
            @1280 batch 32 (N=1), configs[4] SAHI-style tiled inference of 3840x2160 frames (tiles sharded over the N
            GPUs, global NMS merge), the drop-in `YOLO.predict(list, predictor=B200DetectionPredictor)` (N=1), `ingest`
            (`predict("dir of image files", batch=64)`: prefetching loader vs the reference's) and `val`
            (`YOLO.val(validator=B200DetectionValidator)` beside the reference's CPU validator) -- SURVEY §8 f2 / f1, N=1;
            `pose` / `segment` (SURVEY §8 f4, N=1): the Pose / Segment drop-in predictors end to end, device-timed stages,
            the mask kernel's HBM roofline, the reference's CPU predictor beside them.
            `--workload X` makes X the headline instead; `--no-extras` prints the headline only.
"""
from __future__ import annotations

import argparse
import gc
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

MODEL_CFG = {  # workload name -> (arch model, batch, imgsz, BASELINE.json configs index, batch is "per_gpu" | "total")
    "yolo11n": ("yolo11n", 64, 640, 1, "per_gpu"),
    "bsyolo_n": ("bsyolo_n", 256, 640, 2, "total"),
    "yolo11x": ("yolo11x", 32, 1280, 3, "per_gpu"),
}
UNIT = "images/s"


def metric_name(imgsz):
    return f"images/sec @{imgsz}x{imgsz}"  # BASELINE.json metric; the workload (model, batch) is named in config


def make_ir(model):
    from bs_yolo_b200 import arch, ir
    spec = arch.build_spec(model)
    return ir.build_ir(spec, arch.synth_state_dict(spec, 0))


def synth_frames(n, hw, seed):
    hw = (hw, hw) if isinstance(hw, int) else hw
    return np.random.default_rng(seed).integers(0, 256, (n, *hw, 3), dtype=np.uint8)


# ----------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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
        num = lambda v: v.replace(".", "").isdigit()  # noqa: E731
        rows = [r for r in self.rows if len(r) >= 7 and num(r[0])]
        # samples taken while the GPU was under load (nvidia-smi's utilisation is a 1/6-s average: short timed regions between
        # host-side set-up phases rarely reach 50 %)
        busy = [r for r in rows if num(r[6]) and float(r[6]) >= 10] or rows
        sm = [float(r[0]) for r in busy]
        mx = [float(r[1]) for r in rows if num(r[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[2 + j].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(rows), "samples_under_load": len(busy) if busy is not rows else 0,
                "window": "from the first warm-up step to the end of the last timed region (all workloads)"}


# ----------------------------------------------------------------------------- CPU arms
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


def reference_yolo(model, imgsz, task="detect"):
    """YOLO(<synthetic checkpoint written like the reference's trainer does>) of the UNMODIFIED reference package
    (baseline/_ref, pip-installed from /root/reference by __graft_entry__.build()); None if it is not importable."""
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
    if task == "pose":
        m.kpt_shape = list(m.model[-1].kpt_shape)  # the reference's trainer sets it (models/yolo/pose/train.py:49)
    m.args = {"task": task, "imgsz": imgsz}
    path = os.path.join(tempfile.mkdtemp(prefix="b2y_ref_"), f"{model}.pt")
    torch.save({"model": m, "ema": None, "train_args": {"task": task, "imgsz": imgsz}}, path)
    return YOLO(path)


def cpu_reference_throughput(model, imgsz, n_images, batch, threads, steps=1, warmup=1):
    """Times the reference: YOLO(ckpt).predict(list of ndarrays, device="cpu") -- LetterBox + preprocess + fp32 torch
    forward + NMS + Results, through its own public API -- on `threads` host threads."""
    yolo = reference_yolo(model, imgsz)
    if yolo is None:
        return None
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


def workload_label(name, batch_note, imgsz):
    model, _, _, ci, _ = MODEL_CFG[name]
    return f"{model} detect forward + DFL decode + NMS, {batch_note} @ {imgsz}x{imgsz} (BASELINE.json configs[{ci}])"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model, batch, imgsz, _, mode = MODEL_CFG[args.workload]
    if mode == "total":
        batch = max(1, batch // max(args.gpus, 1))
    cores = os.cpu_count() or 1
    # one step = one batch of the workload's size, in ONE predict() call (a list is one batch, data/loaders.py:486), so the
    # arm runs the labelled configuration; yolo11x@1280 is bounded to 2 frames per step (~10 s per frame per thread)
    n_img = batch if imgsz <= 640 else 2
    val, step_s, kind = cpu_arm(model, imgsz, n_img, n_img, cores, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": metric_name(imgsz), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, f"batch {n_img}", imgsz),
                   "same_config": n_img == batch,
                   "note": ("the unmodified reference (baseline/_ref): YOLO(ckpt).predict(list, device='cpu')" if kind == "reference"
                            else "CPU port (oracle) of the reference torch path") + f"; each step = one predict() call on {n_img} frames"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n_img} synthetic {imgsz}x{imgsz} frames per step in one batch, torch fp32, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- ours
class Ctx:
    """Process-wide state of one bench run (rank, device, process group, peaks)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.bind = ""
        if self.world > 1:
            if os.environ.get("B2Y_BIND", "1") != "0":  # before any pinned allocation: first touch on the GPU's NUMA node
                from bs_yolo_b200.shard import bind_to_gpu_numa
                self.bind = bind_to_gpu_numa(self.local)
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm = float(self.peaks.get("hbm_gbs", 6650.0))
        self.peak_src = "measured (MEASURED_PEAKS.json)" if self.peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def timed(self, fn, steps, finish=None):
        """K steps bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks."""
        prof = os.environ.get("B2Y_CUDA_PROFILER_RANGE") == "1"  # `ncu --profile-from-start off`: only the timed steps
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if prof:
            torch.cuda.cudart().cudaProfilerStart()
        t0 = time.perf_counter()
        e0.record()
        for j in range(steps):
            fn(j)
        if finish is not None:
            finish()  # e.g. DetectPipeline.drain: the side-stream work of the last batches belongs to the timed region
        e1.record()
        self.barrier()
        if prof:
            torch.cuda.cudart().cudaProfilerStop()
        wall = (time.perf_counter() - t0) * 1e3
        return self.max_over_ranks(e0.elapsed_time(e1), wall)


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


def run_detect_workload(ctx: Ctx, name: str, K: int, Wm: int, batch_override=0, layers_out="", want_cpu=False):
    """One detect workload on every rank: resident-input throughput, end-to-end throughput, stage split, roofline of the
    dominant kernel.  Returns the record (rank 0) or None."""
    from bs_yolo_b200 import _cabi as A
    from bs_yolo_b200.pipeline import DetectPipeline

    model, batch, imgsz, _, mode = MODEL_CFG[name]
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    if batch_override:
        batch = batch_override
        note = f"batch {batch} per GPU"
    elif mode == "total":
        total = batch
        batch = max(1, total // world)
        note = f"batch {total} sharded {batch} per GPU"
    else:
        note = f"batch {batch} per GPU"
    scaling = "strong" if (mode == "total" and not batch_override) else "weak"

    pipe = DetectPipeline(make_ir(model), batch, imgsz, (imgsz, imgsz), device=ctx.local)
    eng = pipe.engine
    n_rot = 4  # rotate over 4 distinct input batches (> 126 MB L2 together); activations add GBs of traffic per step
    host = [torch.from_numpy(synth_frames(batch, imgsz, 100 + rank * 16 + j)).pin_memory() for j in range(n_rot)]
    devf = [h.to(dev) for h in host]
    gmode = "root" if world > 1 else False  # only the final detections cross GPUs: one NCCL gather to rank 0 per batch

    def step_resident(j):
        # pipelined submit: the NMS (and, multi-GPU, the gather of the padded detections) of this batch runs on the
        # post-process stream underneath the next batch's network
        return pipe.submit(devf[j % n_rot], gather=gmode, src_ready=True)  # frames are resident before the timed region

    for j in range(Wm):
        step_resident(j)
    pipe.drain()
    ms, wall = ctx.timed(step_resident, K, pipe.drain)
    value = world * batch * K / (ms * 1e-3)

    # ---- end to end through the public streaming API with host buffers: every step copies its frames host->device
    #      (pinned memory) and its detections device->host inside the timed region; copies overlap the previous compute
    d2h = [0]

    def run_e2e(steps):
        n = 0
        for dets_h, counts_h in pipe.predict_stream((host[j % n_rot] for j in range(steps)), gather=gmode):
            n += int(counts_h.shape[0])
            d2h[0] = dets_h.numel() * 4 + counts_h.numel() * 4
        return n

    run_e2e(3)
    ctx.barrier()
    t0 = time.perf_counter()
    n_out = run_e2e(K)
    ctx.barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    assert n_out == K * (world * batch if (world > 1 and rank == 0) else batch)
    wall_e2e, d2h_total = ctx.max_over_ranks(wall_e2e, 0)[0], d2h[0]
    if world > 1:
        t = torch.tensor([float(d2h[0])], dtype=torch.float64, device=dev)
        ctx.dist.all_reduce(t)
        d2h_total = int(t[0])
    e2e_value = world * batch * K / (wall_e2e * 1e-3)

    rec = None
    if rank == 0:
        # ---- stage split of one resident step: preprocess | CUDA-graphed network (+ fused decode) | NMS
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

        # ---- per-op timing -> roofline of the dominant kernel
        hbm, peaks = ctx.hbm, ctx.peaks
        eng.forward()
        times = per_op_times(eng, reps=3)
        layers = []
        for i, o in enumerate(eng.plan.ops):
            # algorithmic bytes of the launch: activations x batch + weights ONCE per batch
            layers.append({"i": i, "name": o.name, "kind": A.OP_NAMES[o.kind], "impl": eng.op_impl(i), "ms": times[i],
                           "bytes": eng.plan.op_bytes(o, batch), "flops": o.flops * batch})
        dom = [L for L in layers if L["kind"] == "conv" and L["impl"] == A.IMPL_TCGEN05] or [L for L in layers if L["kind"] == "conv"]
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
            ent = tr.get(name) if isinstance(tr.get(name), dict) else (tr if tr.get("workload") == name else None)
            if ent and ent.get("batch") == batch:
                traffic = ent["conv_tc_dram_bytes_per_step"] / max(ent["conv_tc_launches_per_step"], 1)
        except Exception:
            pass
        # which roof bounds the kernel on this workload: its launches' summed per-layer floors (yolo11n / BS-YOLO: HBM;
        # yolo11x@1280: the tensor pipe -- sustained cuBLAS figure, the kernel runs inside a long step)
        tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        tf_burst = float(peaks.get("bf16_tflops", 1590.0))
        tensor_bound = f_dom / (tf_peak * 1e12) > b_dom / (hbm * 1e9)
        tfl = f_dom / t_dom / 1e12
        net_floor_ms = sum(max(L["bytes"] / (hbm * 1e9), L["flops"] / (tf_burst * 1e12)) for L in layers) * 1e3
        fl_b, by_b = eng.plan.work_per_batch(batch)
        roof = {"bound": "tensor" if tensor_bound else "hbm", "achieved": tfl if tensor_bound else ach,
                "peak": tf_peak if tensor_bound else hbm, "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": (tfl / tf_peak) if tensor_bound else ach / hbm, "traffic": traffic,
                "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv), all its launches of one step" if dom and dom[0]["impl"] == 2 else "conv_simt_kernel",
                "launches_per_step": len(dom), "avg_launch_ms": t_dom * 1e3 / max(len(dom), 1),
                "algorithmic_bytes_per_launch": b_dom / max(len(dom), 1), "tflops": tfl,
                "share_of_step": share * stages["network_graph_ms"] / (ms / K), "peak_source": ctx.peak_src,
                "net_bytes_per_image": by_b / batch, "net_flops_per_image": fl_b / batch,
                "net_bytes_note": "lowered (fused) plan: activations per image + weights/batch",
                "network_floor_ms": net_floor_ms, "network_roofline_frac": net_floor_ms / stages["network_graph_ms"],
                "step_roofline_frac": net_floor_ms / (ms / K)}
        if layers_out:
            with open(layers_out, "w") as f:
                json.dump({"workload": name, "batch": batch, "sum_ms": sum(times), "graph_ms": stages["network_graph_ms"],
                           "layers": layers}, f, indent=1)
        cpu = None
        if want_cpu:
            cores = os.cpu_count() or 1
            n_cpu = 512 if imgsz <= 640 else 4  # ~10-30 s of CPU work on the box's host cores
            cb = min(batch, 64) if imgsz <= 640 else 2
            v, s_, kind = cpu_arm(model, imgsz, n_cpu, cb, cores, 1, 0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{n_cpu} synthetic {imgsz}x{imgsz} frames (predict() calls of {cb}) after a warm-up call, "
                             + ("reference YOLO.predict(device='cpu')" if kind == "reference" else "oracle torch-fp32 port")
                             + f", {cores} threads, {s_:.1f} s"}
        rec = {
            "metric": metric_name(imgsz), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_label(name, note, imgsz),
                       "parallelism": f"dp{world} (images sharded, detections gathered to rank 0)",
                       "cache": f"inputs rotate over {n_rot} batches ({n_rot * batch * imgsz * imgsz * 3 / 1e6:.0f} MB); "
                                "each step streams GBs of activations (> 126 MB L2)",
                       "weights": "seeded synthetic checkpoint (bs_yolo_b200.arch, seed 0)", "conf": 0.25, "iou": 0.7, "max_det": 300,
                       "cpu_binding": ctx.bind or None},
            # whole-job bytes per step: every rank uploads its shard; rank 0 reads back all ranks' detections (gathered over
            # NCCL), the other ranks their own shard
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * batch * imgsz * imgsz * 3,
                    "d2h_bytes_per_step": d2h_total, "ms_per_step": wall_e2e / K},
            "gpu_launches": world * pipe.launches_per_batch * K, "launches_per_step": pipe.launches_per_batch,
            "stages": stages, "roofline": roof, "cpu_baseline": cpu, "wall_ms_per_step": wall / K,
        }
    del pipe, eng, devf, host
    gc.collect()
    torch.cuda.empty_cache()
    return rec


def run_sahi_workload(ctx: Ctx, K: int, Wm: int, frames_per_rank: int = 4):
    """BASELINE.json configs[4]: SAHI-style tiled inference of synthetic 3840x2160 frames, 640 tiles (24 per frame, overlap
    0, SAHI's edge rule), tiles sharded round-robin over the N GPUs, per-frame global class-aware NMS merge at the frame
    owner.  One step = N x F frames (each rank owns F): every GPU runs a batch of 24 F tiles (24/N of each frame)."""
    from bs_yolo_b200.pipeline import DetectPipeline
    from bs_yolo_b200.shard import SahiSharder

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    Hf, Wf, tile, F = 2160, 3840, 640, frames_per_rank
    sh = SahiSharder(lambda B: DetectPipeline(make_ir("bsyolo_n"), B, tile, (tile, tile), device=ctx.local),
                     (Hf, Wf), tile, world=world, rank=rank, group=None, frames=F)
    n_rot = 2
    host = [torch.from_numpy(synth_frames(F, (Hf, Wf), 500 + rank * 8 + j)).pin_memory() for j in range(n_rot)]
    devf = [h.to(dev) for h in host]

    def step_res(j):
        return sh.predict_frames(devf[j % n_rot])

    for j in range(Wm):
        step_res(j)
    ms, wall = ctx.timed(step_res, K)
    # end to end: the owner's frames come from pinned host memory (the upload of step i+1 overlaps step i on a copy
    # stream), the merged detections go back to the host
    out = (torch.empty(F, sh.max_det, 6).pin_memory(), torch.empty(F, dtype=torch.int32).pin_memory())
    stage = [torch.empty_like(devf[0]) for _ in range(2)]
    cs = torch.cuda.Stream(dev)
    up, used = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream(dev)

    def upload(j):
        with torch.cuda.stream(cs):
            cs.wait_event(used[j & 1])
            stage[j & 1].copy_(host[j % n_rot], non_blocking=True)
            up[j & 1].record(cs)

    def run_e2e(steps):
        for e in used:
            e.record(main)
        upload(0)
        for j in range(steps):
            if j + 1 < steps:
                upload(j + 1)
            main.wait_event(up[j & 1])
            d, c = sh.predict_frames(stage[j & 1])
            used[j & 1].record(main)
            out[0].copy_(d, non_blocking=True)
            out[1].copy_(c, non_blocking=True)
        main.synchronize()

    run_e2e(2)
    ctx.barrier()
    t0 = time.perf_counter()
    run_e2e(K)
    ctx.barrier()
    wall_e2e = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)[0]
    rec = None
    if rank == 0:
        fps = world * F * K / (ms * 1e-3)
        rec = {"metric": "frames/sec @3840x2160 (SAHI 640 tiles)", "value": fps, "unit": "frames/s", "tiles_per_s": fps * sh.T,
               "n_gpus": world, "steps": K, "ms_per_step": ms / K, "scaling": "weak", "dtype": "bf16",
               "config": {"workload": f"SAHI-style tiled inference, BS-YOLO-n, {Hf}x{Wf} frames -> {sh.T} tiles of {tile}, "
                                      f"tiles round-robin over {world} GPU(s) ({sh.per} per GPU and frame), global NMS merge at the "
                                      "frame owner (BASELINE.json configs[4])",
                          "frames_per_step": world * F, "tile_batch_per_gpu": sh.B,
                          "exchange": "none (1 GPU)" if world == 1 else "NCCL all_to_all of tiles out and of padded detections back"},
               "e2e": {"value": world * F * K / (wall_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": world * F * Hf * Wf * 3,
                       "d2h_bytes_per_step": world * F * (sh.max_det * 24 + 4), "ms_per_step": wall_e2e / K},
               "merged_detections_last_frame": int(out[1][0])}
    del sh, devf, host, stage
    gc.collect()
    torch.cuda.empty_cache()
    return rec


def run_dropin_workload(ctx: Ctx, K: int):
    """The real drop-in end to end (rank 0, N=1): YOLO(ckpt).predict(list of K*64 host frames, predictor=
    B200DetectionPredictor) -- the reference's own predict loop, source loader, callbacks and Results objects around our
    kernels -- wall clock around the call, pageable host ndarrays in, Results (boxes on the GPU) out."""
    yolo = reference_yolo("yolo11n", 640)
    if yolo is None:
        return {"unavailable": "reference package (baseline/_ref) not importable"}
    from bs_yolo_b200.predictor import B200DetectionPredictor
    n = 64 * max(K, 8)
    frames = list(synth_frames(n, 640, 4321))
    kw = dict(device=ctx.local, imgsz=640, conf=0.25, iou=0.7, max_det=300, verbose=False)
    yolo.predict(frames[:128], predictor=B200DetectionPredictor, **kw)  # builds, autotunes and captures the pipeline
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = yolo.predict(frames, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert len(res) == n and type(yolo.predictor).__name__ == "B200DetectionPredictor"
    sp = res[0].speed
    # what the host alone can deliver: the same frames copied into the pinned staging ring (all cores), no GPU work --
    # pageable ndarrays must cross host memory once more than predict_stream's pinned batches
    from bs_yolo_b200.predictor import _Staging

    class _Done:
        def synchronize(self):
            pass

    ring = _Staging(64, 640, 640)
    fill = None
    for _ in range(3):
        t0 = time.perf_counter()
        for buf in ring.chunks(frames):
            ring.uploaded(buf, _Done())
        dt = time.perf_counter() - t0
        fill = dt if fill is None else min(fill, dt)
    return {"metric": metric_name(640), "value": n / best, "unit": UNIT, "frames_per_call": n,
            "api": "YOLO(ckpt).predict(list_of_ndarrays, predictor=B200DetectionPredictor, device=0) incl. Results construction",
            "h2d_bytes_per_call": n * 640 * 640 * 3, "ms_per_image": best * 1e3 / n, "speed_ms_per_image": sp,
            "detections_first_image": int(res[0].boxes.data.shape[0]), "best_of": 3,
            "host_staging_ceiling": {"value": n / fill, "unit": UNIT, "gbps": n * 640 * 640 * 3 / fill / 1e9,
                                     "what": "pageable frames -> pinned staging ring only (thread pool, no GPU work)"}}


def _write_val_dataset(root, n, hw, nc, seed=99, labels_per_image=4):
    """<root>/images/val/*.png (uncompressed PNG of uniform noise), YOLO-format labels and data.yaml; returns the yaml path."""
    import cv2
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "images", "val"), exist_ok=True)
    os.makedirs(os.path.join(root, "labels", "val"), exist_ok=True)
    for i in range(n):
        cv2.imwrite(os.path.join(root, "images", "val", f"im{i:04d}.png"), rng.integers(0, 256, (hw, hw, 3), dtype=np.uint8),
                    [cv2.IMWRITE_PNG_COMPRESSION, 0])
        with open(os.path.join(root, "labels", "val", f"im{i:04d}.txt"), "w") as f:
            for _ in range(labels_per_image):
                cx, cy, w, h = rng.uniform(0.2, 0.8), rng.uniform(0.2, 0.8), rng.uniform(0.05, 0.3), rng.uniform(0.05, 0.3)
                f.write(f"{int(rng.integers(0, nc))} {cx:.6f} {cy:.6f} {w:.6f} {h:.6f}\n")
    y = os.path.join(root, "data.yaml")
    with open(y, "w") as f:
        f.write(f"path: {root}\ntrain: images/val\nval: images/val\nnames:\n" + "".join(f"  {k}: class{k}\n" for k in range(nc)))
    return y


def _pin_device_count():
    """The reference's select_device("cpu") sets CUDA_VISIBLE_DEVICES=-1 in this process (utils/torch_utils.py:176); the first
    torch.cuda.device_count() after that would cache 0 for good.  Cache the true count while the environment is intact."""
    if torch.cuda.is_available():
        torch.cuda.init()
        torch.cuda.device_count()


def run_val_workload(ctx: Ctx):
    """SURVEY §8 f1, rank 0 at N=1: YOLO(ckpt).val(data, validator=B200DetectionValidator) -- the reference's dataset,
    dataloader (8 workers: PNG decode + LetterBox on the host) and metric code around our preprocess / network / multi_label
    NMS at conf 0.001 / scale_boxes / match_predictions kernels -- on 256 synthetic 640x640 files, batch 64, wall clock of
    the second call; the reference's own CPU validator on a 16-image subset beside it."""
    yolo = reference_yolo("yolo11n", 640)
    if yolo is None:
        return {"unavailable": "reference package (baseline/_ref) not importable"}
    import tempfile
    from bs_yolo_b200.validator import B200DetectionValidator
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    open(os.path.join(os.environ["YOLO_CONFIG_DIR"], "Arial.ttf"), "a").close()  # check_font would try to download it
    tmp = tempfile.mkdtemp(prefix="b2y_val_")
    n, n_ref = 256, 16
    data = _write_val_dataset(os.path.join(tmp, "ds"), n, 640, 80)
    data_ref = _write_val_dataset(os.path.join(tmp, "ds_ref"), n_ref, 640, 80)
    kw = dict(imgsz=640, plots=False, verbose=False, rect=False, project=os.path.join(tmp, "runs"))
    best, speed = None, None
    for i in range(3):  # the first call builds, autotunes and captures the plan (kept per model) and writes the label cache
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = yolo.val(data=data, validator=B200DetectionValidator, batch=64, device=ctx.local, workers=8, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0 and (best is None or dt < best):
            best, speed = dt, dict(m.speed)
    torch.set_num_threads(os.cpu_count() or 1)
    yolo.val(data=data_ref, batch=16, device="cpu", **kw)
    t0 = time.perf_counter()
    yolo.val(data=data_ref, batch=16, device="cpu", **kw)
    dt_ref = time.perf_counter() - t0
    return {"metric": "images/sec @640x640 through model.val()", "value": n / best, "unit": UNIT, "images": n, "batch": 64,
            "api": "YOLO(ckpt).val(data=yaml, validator=B200DetectionValidator, device=0) incl. dataloader and metrics",
            "speed_ms_per_image": speed, "best_of": 2,
            "cpu_baseline": {"value": n_ref / dt_ref, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                             "sample": f"{n_ref} images, YOLO.val(device='cpu', batch=16), second call"}}


def run_ingest_workload(ctx: Ctx):
    """SURVEY §8 f2, rank 0 at N=1: YOLO(ckpt).predict("dir/of/256 png files", batch=64, predictor=B200DetectionPredictor)
    with the prefetching loader (ingest.py) and, beside it, with the reference's sequential loader (B2Y_INGEST=0)."""
    yolo = reference_yolo("yolo11n", 640)
    if yolo is None:
        return {"unavailable": "reference package (baseline/_ref) not importable"}
    import tempfile
    from bs_yolo_b200.predictor import B200DetectionPredictor
    tmp = tempfile.mkdtemp(prefix="b2y_ingest_")
    n = 256
    t_ds = time.perf_counter()
    _write_val_dataset(tmp, n, 640, 80, seed=5)
    t_ds = time.perf_counter() - t_ds
    src = os.path.join(tmp, "images", "val")
    kw = dict(device=ctx.local, imgsz=640, conf=0.25, iou=0.7, max_det=300, verbose=False, batch=64)
    out, passes = {}, {"dataset_s": round(t_ds, 2)}
    trace = os.environ.get("B2Y_TRACE_INGEST") == "1"  # debug: dump every thread's stack every 15 s (where does a slow pass sit?)
    if trace:
        import faulthandler
        faulthandler.dump_traceback_later(15, repeat=True, file=sys.stderr)
    for name, env in (("prefetch", "1"), ("reference_loader", "0")):
        os.environ["B2Y_INGEST"] = env
        try:
            best = None
            for i in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = yolo.predict(src, predictor=B200DetectionPredictor, **kw) if yolo.predictor is None else yolo.predict(src, **kw)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                passes.setdefault(name + "_s", []).append(round(dt, 3))
                best = dt if (i > 0 and (best is None or dt < best)) else best
            assert len(res) == n
            out[name] = n / best
        finally:
            os.environ.pop("B2Y_INGEST", None)
    if trace:
        faulthandler.cancel_dump_traceback_later()
    return {"metric": "images/sec @640x640 from image files", "value": out["prefetch"], "unit": UNIT, "files": n, "batch": 64,
            "api": "YOLO(ckpt).predict('dir', batch=64, predictor=B200DetectionPredictor) incl. PNG decode and Results",
            "with_reference_loader": out["reference_loader"], "best_of": 2, "wall": passes,
            "what": "uncompressed 640x640 PNG files decoded by cv2 on the host cores; prefetch = ingest.PrefetchImageLoader"}


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: everything else that writes to file descriptor 1 while the bench runs (the
    reference package's logger and settings banner, NCCL, torch warnings) is sent to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_head_workload(ctx: Ctx, head: str, n: int = 32):
    """SURVEY §8 f4, rank 0 at N=1: the Pose / Segment heads of the stock yolo11n-pose / yolo11n-seg graphs through their
    drop-in predictors -- YOLO(ckpt).predict(list of n host frames, predictor=B200PosePredictor | B200SegmentationPredictor)
    incl. Results (keypoints / masks stay on the GPU), wall clock, mean of 3 calls after the call that builds the engine --
    plus device-timed stages on the predictor's own engine (CUDA events) and the reference's own CPU predictor on 8 frames.
    The synthetic checkpoint fills max_det (300 rows per image): the worst case for the mask kernel, which writes
    rows x 640 x 640 fp32 and is measured against the HBM copy peak."""
    model, task = {"pose": ("yolo11n-pose", "pose"), "segment": ("yolo11n-seg", "segment")}[head]
    yolo = reference_yolo(model, 640, task)
    if yolo is None:
        return {"unavailable": "reference package (baseline/_ref) not importable"}
    from bs_yolo_b200 import ops
    from bs_yolo_b200 import predictor as P
    cls = P.B200PosePredictor if head == "pose" else P.B200SegmentationPredictor
    frames = list(synth_frames(n, 640, 777))
    kw = dict(device=ctx.local, imgsz=640, conf=0.25, iou=0.7, max_det=300, verbose=False)
    res = yolo.predict(frames, predictor=cls, **kw)  # builds, autotunes and captures the (n, 640, 640) engine
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        res = yolo.predict(frames, **kw)
    torch.cuda.synchronize()
    e2e = 3 * n / (time.perf_counter() - t0)
    assert len(res) == n and type(yolo.predictor).__name__ == cls.__name__
    eng = next(iter(yolo.predictor.model._engines.values()))

    def dev_ms(fn, iters=10, warm=2):
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

    rec = {"metric": metric_name(640), "value": None, "unit": UNIT, "frames_per_call": n, "model": model,
           "api": f"YOLO(ckpt).predict(list_of_ndarrays, predictor={cls.__name__}, device=0) incl. Results construction",
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n * 640 * 640 * 3, "d2h_bytes_per_step": n * 4},
           "detections_per_image": float(np.mean([len(r.boxes) for r in res])), "plan_launches": eng.n_launches}
    del res
    with torch.cuda.device(eng.device):
        stages = {"network_graph_ms": dev_ms(eng.forward)}
        dets, counts, kept = eng.nms(0.25, 0.7, 300, return_anchors=True)
        nl = counts.tolist()
        stages["nms_ms"] = dev_ms(lambda: eng.nms(0.25, 0.7, 300, return_anchors=True))
        if head == "pose":
            geom = ops.coords_geometry((640, 640), [(640, 640)] * n, dets.device)
            stages["keypoints_ms"] = dev_ms(lambda: eng.keypoints(kept, counts, geom=geom))
        else:
            coeffs = eng.mask_coeffs(kept, counts)
            stages["coeff_gather_ms"] = dev_ms(lambda: eng.mask_coeffs(kept, counts))
            ms = dev_ms(lambda: ops.process_mask_batched(eng.protos, coeffs, dets, nl, (640, 640)))
            stages["process_mask_ms"] = ms
            nbytes = int(sum(nl)) * 640 * 640 * 4
            traffic = None  # DRAM bytes of this launch, scaled from the committed ncu capture (bytes per mask x masks)
            try:
                ent = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("process_mask")
                if ent:
                    traffic = ent["dram_bytes_per_mask"] * int(sum(nl))
            except (OSError, ValueError, KeyError):
                pass
            rec["roofline"] = {"bound": "hbm", "kernel": "process_mask_kernel", "achieved": nbytes / ms / 1e6, "peak": ctx.hbm,
                               "unit": "GB/s", "frac": nbytes / ms / 1e6 / ctx.hbm, "peak_source": ctx.peak_src, "traffic": traffic,
                               "what": f"{int(sum(nl))} masks of 640x640 fp32 written per launch (algorithmic bytes = the output)"}
    rec["stages"] = stages
    # per-layer roofline of the lowered network incl. the head's extra branches (same definition as the headline's)
    tf_burst = float(ctx.peaks.get("bf16_tflops", 1590.0))
    floor_ms = sum(max(eng.plan.op_bytes(o, n) / (ctx.hbm * 1e9), o.flops * n / (tf_burst * 1e12)) for o in eng.plan.ops) * 1e3
    rec["network_floor_ms"], rec["network_roofline_frac"] = floor_ms, floor_ms / stages["network_graph_ms"]
    rec["value"] = n / sum(stages.values()) * 1e3  # device-timed: network graph + NMS + head post-process, frames resident

    def cpu_ref():  # deferred to the end of the bench: the reference's CPU runs slow later engine builds in this process
        torch.set_num_threads(os.cpu_count() or 1)
        ykw = dict(device="cpu", imgsz=640, conf=0.25, iou=0.7, max_det=300, verbose=False)
        ref = reference_yolo(model, 640, task)
        ref.predict(frames[:1], **ykw)
        t0 = time.perf_counter()
        ref.predict(frames[:8], **ykw)
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 8 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
                               "sample": f"8 frames, one predict() call of the reference's {type(ref.predictor).__name__} (device='cpu')"}

    _DEFERRED.append(cpu_ref)
    return rec


_DEFERRED: list = []


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="yolo11n", choices=list(MODEL_CFG))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch of the headline workload (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no `workloads` sub-records)")
    ap.add_argument("--extras", default="bsyolo_n,yolo11x,sahi,dropin,ingest,pose,segment,val", help="comma list of sub-records to run")
    ap.add_argument("--layers-out", default="", help="write the per-op timing table of the headline workload to this JSON file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    ctx = Ctx()
    _pin_device_count()
    K, Wm = args.steps, args.warmup
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
    line = run_detect_workload(ctx, args.workload, K, Wm, args.batch, args.layers_out,
                               want_cpu=(ctx.world == 1 and not args.no_cpu_baseline))
    subs = {}
    if not args.no_extras:
        todo = [w for w in args.extras.split(",") if w and w != args.workload]
        Ks = max(3, min(K, 10))
        for w in todo:
            t0 = time.perf_counter()
            try:
                if w == "bsyolo_n":
                    r = run_detect_workload(ctx, "bsyolo_n", Ks, 3)
                elif w == "yolo11x":
                    # configs[3] is a single-GPU configuration: measured on rank 0's GPU at N = 1 only
                    r = run_detect_workload(ctx, "yolo11x", max(3, Ks // 2), 3) if ctx.world == 1 else None
                elif w == "sahi":
                    r = run_sahi_workload(ctx, Ks, 3)
                elif w == "dropin":
                    r = run_dropin_workload(ctx, Ks) if ctx.world == 1 else None
                elif w == "val":
                    r = run_val_workload(ctx) if ctx.world == 1 else None
                elif w == "ingest":
                    r = run_ingest_workload(ctx) if ctx.world == 1 else None
                elif w in ("pose", "segment"):
                    r = run_head_workload(ctx, w) if ctx.world == 1 else None
                else:
                    continue
            except Exception as e:  # noqa: BLE001 -- a failing sub-record must not take the headline down
                if ctx.world > 1:
                    raise
                r = {"error": f"{type(e).__name__}: {e}"[:300]}
            if r is not None and ctx.rank == 0:
                r["bench_wall_s"] = round(time.perf_counter() - t0, 1)
                subs[w] = r
    if ctx.rank == 0:
        line["clocks"] = sampler.stop()
        for fn in _DEFERRED:  # CPU reference legs of the pose / segment sub-records (after every GPU workload and the clock window)
            try:
                fn()
            except Exception as e:  # noqa: BLE001
                print(f"deferred cpu baseline failed: {type(e).__name__}: {e}", file=sys.stderr)
        if subs:
            line["workloads"] = subs
        emit(line)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

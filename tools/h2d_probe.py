#!/usr/bin/env python
"""Copy-only probe of the box's host->device ceiling: N ranks (torchrun, one per GPU) each upload
pinned uint8 batches of the bench's size back to back, nothing else runs.  Prints one JSON line per configuration:
per-GPU and aggregate GB/s, with and without binding the rank to the GPU's NUMA-local cores before the pinned
allocation, one or two copy streams.  Also prints the topology (nvidia-smi topo -m, lscpu NUMA lines) once.

    python tools/h2d_probe.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        for cmd in (["nvidia-smi", "topo", "-m"], ["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name|^cpu\\(s\\)'"],
                    ["bash", "-c", "cat /proc/self/status | grep -i cpus_allowed_list; nproc; free -g | head -2"]):
            try:
                print(subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout, file=sys.stderr)
            except Exception as e:  # noqa: BLE001
                print(f"[{cmd[0]} failed: {e}]", file=sys.stderr)
    nbytes = 64 * 640 * 640 * 3
    reps = 40

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(tag, bind, streams):
        note = ""
        if bind:
            from bs_yolo_b200.shard import bind_to_gpu_numa
            note = bind_to_gpu_numa(local)
        host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
        for h in host:
            h.fill_(1)  # first touch
        dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        ss = [torch.cuda.Stream(dev) for _ in range(streams)]
        half = nbytes // 2

        def run(n):
            for i in range(n):
                if streams == 1:
                    with torch.cuda.stream(ss[0]):
                        dst[i & 1].copy_(host[i & 3], non_blocking=True)
                else:
                    with torch.cuda.stream(ss[0]):
                        dst[i & 1][:half].copy_(host[i & 3][:half], non_blocking=True)
                    with torch.cuda.stream(ss[1]):
                        dst[i & 1][half:].copy_(host[i & 3][half:], non_blocking=True)
            for s_ in ss:
                s_.synchronize()

        run(5)
        barrier()
        t0 = time.perf_counter()
        run(reps)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        if rank == 0:
            per = nbytes * reps / dt / 1e9
            print(json.dumps({"probe": "h2d", "config": tag, "n_gpus": world, "bytes_per_copy": nbytes, "copies": reps,
                              "gbps_per_gpu": round(per, 2), "gbps_aggregate": round(per * world, 2),
                              "images_per_s_ceiling": round(per * world * 1e9 / (640 * 640 * 3)), "binding_rank0": note}))
        del host, dst

    measure("unbound, 1 copy stream", False, 1)
    measure("unbound, 2 copy streams", False, 2)
    measure("bound to the GPU's NUMA cores, 1 copy stream", True, 1)
    measure("bound to the GPU's NUMA cores, 2 copy streams", True, 2)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

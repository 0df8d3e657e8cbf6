"""Data-parallel sharding of the detect path across the GPUs of one box.

Images and SAHI tiles are independent units end to end (preprocess, forward, decode, per-image NMS), so the plan is
replicated per GPU, units are sharded, and there is NO collective on the layer path: only the final padded detections
(<= max_det x 6 floats + a count per image, 7.2 KB) cross GPUs, in one all-gather.  The reference has no multi-GPU predict
(single device, nn/autobackend.py:146); SAHI tiling lives in the third-party `sahi` package (detect-sahi.py:1-13), whose
slicing rule is restated in `sahi_tiles`.

Everything here is host logic over torch.distributed (backend "nccl" on the GPUs, "gloo" in the CPU tests)."""
from __future__ import annotations

import torch


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of n units for `rank`; the first n % world ranks take one extra unit."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sahi_tiles(height: int, width: int, tile: int = 640, overlap: float = 0.0) -> list[tuple[int, int]]:
    """Top-left corners (x0, y0) of SAHI's slices in its order (rows outer, columns inner): step = tile*(1-overlap); a
    slice that would cross the right/bottom edge is shifted back inside the image (sahi.slicing.get_slice_bboxes).
    3840x2160 with 640 tiles, overlap 0 -> x in {0..3200 step 640}, y in {0, 640, 1280, 1520}: 24 tiles."""
    if tile <= 0 or not (0.0 <= overlap < 1.0):
        raise ValueError("tile must be positive and overlap in [0, 1)")
    step = tile - int(overlap * tile)
    out = []
    y = 0
    while True:
        y0 = max(0, height - tile) if y + tile >= height else y
        x = 0
        while True:
            x0 = max(0, width - tile) if x + tile >= width else x
            out.append((x0, y0))
            if x + tile >= width:
                break
            x += step
        if y + tile >= height:
            break
        y += step
    return out


def tile_owner(t: int, world: int) -> int:
    """Round-robin tile -> GPU map (24 tiles over 8 GPUs = 3 per GPU per frame)."""
    return t % world


def gather_detections(dets: torch.Tensor, counts: torch.Tensor, group=None, dst: int | None = None):
    """Gather equal-sized shards of padded detections, (b, max_det, 6) fp32 + (b,) int32 per rank, rank-major (== unit
    order for contiguous shards).  ``dst=None``: all-gather, every rank receives (world*b, ...).  ``dst=r``: only rank r
    receives the whole set (one consumer: world x less traffic than the all-gather); the other ranks get their own shard
    back unchanged."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if dst is None:
        out_d = torch.empty((world * dets.shape[0], *dets.shape[1:]), dtype=dets.dtype, device=dets.device)
        out_c = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
        dist.all_gather_into_tensor(out_d, dets.contiguous(), group=group)
        dist.all_gather_into_tensor(out_c, counts.contiguous(), group=group)
        return out_d, out_c
    rank = dist.get_rank(group)
    # one packed message per rank: the counts ride in a trailing row block of the fp32 buffer (bit-cast, not converted)
    b, k = dets.shape[0], dets.shape[1] * dets.shape[2]
    packed = torch.empty(b, k + 1, dtype=torch.float32, device=dets.device)
    packed[:, :k] = dets.reshape(b, k)
    packed[:, k] = counts.view(torch.float32) if counts.dtype == torch.int32 else counts.to(torch.int32).view(torch.float32)
    if rank == dst:
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.gather(packed, parts, dst=dst, group=group)
        allp = torch.cat(parts, 0)
        return (allp[:, :k].reshape(world * b, *dets.shape[1:]).contiguous(), allp[:, k].contiguous().view(torch.int32))
    dist.gather(packed, None, dst=dst, group=group)
    return dets, counts


def bind_to_gpu_numa(device_index: int) -> str:
    """Pin the calling process to the CPU cores of the NUMA node the GPU hangs off (NVML's ideal-CPU mask), so that the
    pinned staging rings allocated afterwards are first-touched in local memory and the copy/launch threads do not cross
    sockets.  Returns a short description (also when nothing could be done -- containers often hide the topology)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + j for i, w in enumerate(words) for j in range(64) if (w >> j) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"gpu{phys}: no narrower CPU set than the current {len(allowed)} cores"
        os.sched_setaffinity(0, cpus)
        return f"gpu{phys}: bound to {len(cpus)} cores {min(cpus)}-{max(cpus)}"
    except Exception as e:  # noqa: BLE001 -- binding is an optimisation, never a requirement
        return f"not bound ({type(e).__name__}: {e})"


def merge_prediction(dets: torch.Tensor, counts: torch.Tensor, origins: torch.Tensor, nc: int) -> torch.Tensor:
    """Tile detections -> one decoded-prediction tensor (1, 4+nc, T*max_det) in full-frame coordinates, the input of the
    global class-aware NMS (the same kernel as the per-image NMS).  dets (T, max_det, 6) rows [x1,y1,x2,y2,conf,cls],
    origins (T, 2) = (x0, y0) of each tile.  Rows past counts[t] get score 0 (never candidates).  All fp32 arithmetic is
    elementwise IEEE (shift, xyxy -> xywh), so the CPU restatement in the tests reproduces it bit for bit."""
    T, K, _ = dets.shape
    d = dets.reshape(T * K, 6)
    off = origins.to(dets.dtype).repeat_interleave(K, 0)
    x1, y1, x2, y2 = d[:, 0] + off[:, 0], d[:, 1] + off[:, 1], d[:, 2] + off[:, 0], d[:, 3] + off[:, 1]
    valid = (torch.arange(K, device=dets.device).repeat(T) < counts.to(torch.int64).repeat_interleave(K))
    pred = torch.zeros(1, 4 + nc, T * K, dtype=dets.dtype, device=dets.device)
    pred[0, 0], pred[0, 1], pred[0, 2], pred[0, 3] = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
    cls = d[:, 5].to(torch.int64).clamp_(0, nc - 1)
    pred[0, 4:].scatter_(0, cls.unsqueeze(0), torch.where(valid, d[:, 4], torch.zeros_like(d[:, 4])).unsqueeze(0))
    return pred


def sahi_predict(pipe, frame_u8: torch.Tensor, tile: int = 640, overlap: float = 0.0, iou: float | None = None):
    """SAHI-style tiled inference of ONE frame on ONE GPU (BASELINE.json config 5; multi-GPU = tiles sharded by
    `tile_owner`, detections gathered with `gather_detections`, then the same merge): slice the (H,W,3) uint8 BGR device
    frame into `sahi_tiles`, run the tile batch through `pipe` (a DetectPipeline built for batch = number of tiles and
    tile-sized sources), shift the boxes to frame coordinates and merge them with one global class-aware NMS (the same
    kernel as the per-image NMS).  Returns (dets (max_det,6), count) on the device plus the per-tile results."""
    from .ops import nms_raw

    H, W = int(frame_u8.shape[0]), int(frame_u8.shape[1])
    origins = sahi_tiles(H, W, tile, overlap)
    if len(origins) != pipe.B:
        raise ValueError(f"pipeline batch {pipe.B} != {len(origins)} tiles of a {H}x{W} frame")
    tiles = torch.stack([frame_u8[y:y + tile, x:x + tile] for x, y in origins]).contiguous()
    dets, counts = pipe(tiles)
    org = torch.tensor(origins, dtype=torch.float32, device=dets.device)
    pred = merge_prediction(dets, counts, org, pipe.engine.nc)
    md, mc = nms_raw(pred, pipe.conf, pipe.iou if iou is None else iou, pipe.max_det)
    return md[0], mc[0], (dets, counts, org)

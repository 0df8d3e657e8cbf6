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
    """Tile detections -> decoded-prediction tensors in full-frame coordinates, the input of the global class-aware NMS
    (the same kernel as the per-image NMS).  One frame: dets (T, max_det, 6) rows [x1,y1,x2,y2,conf,cls], counts (T,),
    origins (T, 2) = (x0, y0) of each tile -> (1, 4+nc, T*max_det).  A batch of frames: dets (F, T, max_det, 6), counts
    (F, T), the same origins -> (F, 4+nc, T*max_det).  Rows past counts[t] get score 0 (never candidates).  All fp32
    arithmetic is elementwise IEEE (shift, xyxy -> xywh), so the CPU restatement in the tests reproduces it bit for bit."""
    single = dets.dim() == 3
    if single:
        dets, counts = dets.unsqueeze(0), counts.unsqueeze(0)
    F, T, K, _ = dets.shape
    d = dets.reshape(F, T * K, 6)
    off = origins.to(dets.dtype).repeat_interleave(K, 0).unsqueeze(0)            # (1, T*K, 2)
    x1, y1, x2, y2 = d[..., 0] + off[..., 0], d[..., 1] + off[..., 1], d[..., 2] + off[..., 0], d[..., 3] + off[..., 1]
    valid = torch.arange(K, device=dets.device).repeat(T).unsqueeze(0) < counts.to(torch.int64).repeat_interleave(K, 1)
    pred = torch.zeros(F, 4 + nc, T * K, dtype=dets.dtype, device=dets.device)
    pred[:, 0], pred[:, 1], pred[:, 2], pred[:, 3] = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
    cls = d[..., 5].to(torch.int64).clamp_(0, nc - 1)
    pred[:, 4:].scatter_(1, cls.unsqueeze(1), torch.where(valid, d[..., 4], torch.zeros_like(d[..., 4])).unsqueeze(1))
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


class SahiSharder:
    """BASELINE.json config 5 across the GPUs of one box: every rank owns ``frames`` frames per step, slices them into SAHI
    tiles on its GPU, and the tiles are sharded round-robin (tile t -> GPU t mod world) with ONE all_to_all of uint8 tiles;
    every GPU then runs its tile batch (``per`` tiles of each of the ``world * frames`` frames) through the detect
    pipeline, the padded per-tile detections return to the frame owners with a second all_to_all, and each owner merges
    its frames with one batched global class-aware NMS (`merge_prediction` + the same NMS kernel as the per-image NMS).
    These two exchanges are the only cross-GPU traffic (world == 1: none).  Host logic only -- `make_pipe(B)` builds the
    per-GPU detector (a DetectPipeline on the GPU; any callable ``tiles (B,t,t,3) uint8 -> (dets (B,K,6), counts (B,))``
    in the CPU tests)."""

    def __init__(self, make_pipe, frame_hw, tile: int = 640, world: int = 1, rank: int = 0, group=None, overlap: float = 0.0,
                 iou: float | None = None, nms_fn=None, frames: int = 1):
        self.H, self.W = frame_hw
        self.tile, self.world, self.rank, self.group, self.F = tile, world, rank, group, frames
        self.origins = sahi_tiles(self.H, self.W, tile, overlap)
        self.T = len(self.origins)
        self.per = -(-self.T // world)              # tiles of one frame per GPU (the last GPUs may get a padding tile)
        self.B = self.per * world * frames          # tile batch per GPU and step
        # destination-major send order of ONE frame: block d = its tiles for GPU d (tile_owner), padded with -1
        self.send_order = []
        for d in range(world):
            mine = [t for t in range(self.T) if tile_owner(t, world) == d]
            self.send_order += mine + [-1] * (self.per - len(mine))
        self.pipe = make_pipe(self.B)
        self.max_det = getattr(self.pipe, "max_det", 300)
        self.nc = self.pipe.engine.nc if hasattr(self.pipe, "engine") else getattr(self.pipe, "nc")
        self.conf = getattr(self.pipe, "conf", 0.25)
        self.iou = iou if iou is not None else getattr(self.pipe, "iou", 0.7)
        self._nms = nms_fn
        self._dev_cache = None
        self._inflight = [None, None]   # completion events of the last two steps (host run-ahead is bounded to two steps)
        self._step = 0

    def _consts(self, device):
        if self._dev_cache is None or self._dev_cache[0] != device:
            slots = self.per * self.world
            valid = torch.tensor([t >= 0 for t in self.send_order], device=device)
            # permutation that puts one frame's slots back into tile order (ties in the merge NMS break by index, so the
            # merged result equals the single-GPU one bit for bit); padding slots go last and are dropped
            order = sorted(range(slots), key=lambda i: (self.send_order[i] < 0, self.send_order[i]))[: self.T]
            org = torch.tensor(self.origins, dtype=torch.float32, device=device)
            self._dev_cache = (device, valid, org, torch.tensor(order, device=device))
        return self._dev_cache[1:]

    def slice_tiles(self, frames: torch.Tensor) -> torch.Tensor:
        """(F,H,W,3) -> (world, F, per, t, t, 3): destination-major, so all_to_all_single splits it evenly."""
        t, F = self.tile, frames.shape[0]
        tiles = torch.stack([frames[:, y:y + t, x:x + t] for x, y in (self.origins[max(i, 0)] for i in self.send_order)], 1)
        return tiles.view(F, self.world, self.per, t, t, 3).transpose(0, 1).contiguous()

    def predict_frames(self, frames_u8: torch.Tensor):
        """frames_u8 (F,H,W,3) uint8 BGR on this rank's device -> (dets (F,max_det,6) in frame coordinates, counts (F,))
        for THIS rank's frames.  Collective: every rank of the group calls it once per step."""
        import torch.distributed as dist

        F = self.F
        assert tuple(frames_u8.shape) == (F, self.H, self.W, 3) and frames_u8.dtype == torch.uint8
        # back-pressure: a step allocates ~250 MB of temporaries (tile batch, exchange buffers); a host that runs many steps
        # ahead of the GPU makes the caching allocator grow and cudaMalloc under load (measured: 4.7 -> 6-20 ms per step)
        prev = self._inflight[self._step & 1]
        if prev is not None:
            prev.synchronize()
        tiles = self.slice_tiles(frames_u8).view(self.B, self.tile, self.tile, 3)
        if self.world > 1:
            recv = torch.empty_like(tiles)
            dist.all_to_all_single(recv, tiles, group=self.group)   # block s of recv: my share of rank s's frames
            tiles = recv
        dets, counts = self.pipe(tiles)
        K = dets.shape[1]
        if self.world > 1:
            packed = torch.empty(self.B, K * 6 + 1, dtype=torch.float32, device=dets.device)
            packed[:, :K * 6] = dets.reshape(self.B, K * 6)
            packed[:, K * 6] = counts.to(torch.int32).view(torch.float32)   # bit-cast: the count rides in the fp32 message
            back = torch.empty_like(packed)
            dist.all_to_all_single(back, packed, group=self.group)  # block d of back: GPU d's results for my frames' tiles
            dets = back[:, :K * 6].reshape(self.B, K, 6)
            counts = back[:, K * 6].contiguous().view(torch.int32)
        valid, org, order = self._consts(dets.device)
        # (world, F, per) -> (F, world*per) slots per frame -> tile order
        slots = self.world * self.per
        dets = dets.view(self.world, F, self.per, K, 6).transpose(0, 1).reshape(F, slots, K, 6).index_select(1, order)
        counts = counts.to(torch.int32).view(self.world, F, self.per).transpose(0, 1).reshape(F, slots).index_select(1, order)
        pred = merge_prediction(dets, counts, org, self.nc)
        nms = self._nms
        if nms is None:
            from .ops import nms_raw as nms
        out = nms(pred, self.conf, self.iou, self.max_det)
        if dets.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dets.device))
            self._inflight[self._step & 1] = ev
        self._step += 1
        return out

    def predict_frame(self, frame_u8: torch.Tensor):
        """One frame per rank and step (frames == 1): (H,W,3) -> (dets (max_det,6), count)."""
        md, mc = self.predict_frames(frame_u8.unsqueeze(0))
        return md[0], mc[0]

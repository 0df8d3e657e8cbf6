"""Host-side multi-GPU logic on CPU: shard maps, SAHI tile enumeration, the detection all-gather over a world_size-2 gloo
group, and the tile-merge prediction fed to the global NMS (checked against the reference's NMS when importable)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from bs_yolo_b200 import shard
from oracle import oracle_post, ref_import


def test_shard_ranges_partition_the_units():
    for n in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard.shard_range(256, 8, r) for r in (0, 7)] == [(0, 32), (224, 256)]  # BASELINE config 3: 256/8 per GPU
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_sahi_tiles_config5():
    t = shard.sahi_tiles(2160, 3840, 640, 0.0)
    assert len(t) == 24
    assert sorted({x for x, _ in t}) == [0, 640, 1280, 1920, 2560, 3200]
    assert sorted({y for _, y in t}) == [0, 640, 1280, 1520]  # last row shifted back inside the frame
    assert t[:7] == [(0, 0), (640, 0), (1280, 0), (1920, 0), (2560, 0), (3200, 0), (0, 640)]
    owners = [shard.tile_owner(i, 8) for i in range(24)]
    assert all(owners.count(g) == 3 for g in range(8))
    # every pixel is covered; tiles stay inside the frame
    cover = np.zeros((2160, 3840), bool)
    for x, y in t:
        assert 0 <= x <= 3840 - 640 and 0 <= y <= 2160 - 640
        cover[y:y + 640, x:x + 640] = True
    assert cover.all()
    assert shard.sahi_tiles(480, 600, 640) == [(0, 0)]  # frame smaller than a tile
    assert len(shard.sahi_tiles(1080, 1920, 800, 0.2)) == 6  # detect-sahi.py's own 800-px / 0.2 setting: 3 x 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(n_units, world, rank)
        g = torch.Generator().manual_seed(1234)
        all_d = torch.rand(n_units, 5, 6, generator=g)
        all_c = torch.randint(0, 6, (n_units,), generator=g, dtype=torch.int32)
        d, c = shard.gather_detections(all_d[lo:hi], all_c[lo:hi])  # each rank contributes only its shard
        ok = torch.equal(d, all_d) and torch.equal(c, all_c)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks timing reduction
        q.put((rank, bool(ok), float(t)))
    finally:
        dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 2.0), (1, True, 2.0)]


def _tile_dets(rng, T, K, nc, tile=640):
    dets = np.zeros((T, K, 6), np.float32)
    counts = rng.integers(0, K + 1, T).astype(np.int32)
    for t in range(T):
        n = counts[t]
        xy = rng.uniform(0, tile - 60, (n, 2))
        wh = rng.uniform(8, 60, (n, 2))
        dets[t, :n, 0:2] = xy
        dets[t, :n, 2:4] = xy + wh
        dets[t, :n, 4] = np.sort(rng.uniform(0.25, 1.0, n))[::-1]
        dets[t, :n, 5] = rng.integers(0, nc, n)
    return dets, counts


def test_tile_merge_prediction_feeds_global_nms():
    """merge_prediction + NMS == shifting the tile boxes by hand and running NMS on them (oracle; reference if present)."""
    rng = np.random.default_rng(5)
    nc, K = 12, 40
    origins = np.array(shard.sahi_tiles(1000, 1500, 640, 0.0), np.float32)  # overlapping edge tiles -> duplicates to merge
    T = len(origins)
    dets, counts = _tile_dets(rng, T, K, nc)
    dets[1, 0, :4] = dets[0, 0, :4] + np.array([origins[0, 0] - origins[1, 0], origins[0, 1] - origins[1, 1]] * 2, np.float32)
    dets[1, 0, 5] = dets[0, 0, 5]  # the same object seen by two tiles
    counts[:2] = np.maximum(counts[:2], 1)
    pred = shard.merge_prediction(torch.from_numpy(dets), torch.from_numpy(counts), torch.from_numpy(origins), nc)
    assert pred.shape == (1, 4 + nc, T * K)
    out = oracle_post.non_max_suppression(pred.numpy(), 0.25, 0.5, max_det=300)[0]
    # by hand: rows of all tiles in tile order, shifted, as (cx,cy,w,h,one-hot score)
    rows = []
    for t in range(T):
        for k in range(counts[t]):
            x1, y1, x2, y2, s, c = dets[t, k]
            ox, oy = origins[t]
            rows.append((np.float32(x1 + ox), np.float32(y1 + oy), np.float32(x2 + ox), np.float32(y2 + oy), s, int(c)))
    p2 = np.zeros((1, 4 + nc, len(rows)), np.float32)
    for i, (x1, y1, x2, y2, s, c) in enumerate(rows):
        p2[0, :4, i] = ((x1 + x2) / np.float32(2), (y1 + y2) / np.float32(2), x2 - x1, y2 - y1)
        p2[0, 4 + c, i] = s
    ref = oracle_post.non_max_suppression(p2, 0.25, 0.5, max_det=300)[0]
    assert out.tobytes() == ref.tobytes() and 0 < len(out) < sum(counts)
    if ref_import.available():
        ref_import.import_reference()
        from ultralytics.utils import ops
        r = ops.non_max_suppression(torch.from_numpy(pred.numpy().copy()), 0.25, 0.5, max_det=300)[0].numpy()
        assert out.tobytes() == r.tobytes()


def _worker_root(rank, world, port, n_units, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(n_units, world, rank)
        g = torch.Generator().manual_seed(99)
        all_d = torch.rand(n_units, 5, 6, generator=g)
        all_c = torch.randint(0, 6, (n_units,), generator=g, dtype=torch.int32)
        d, c = shard.gather_detections(all_d[lo:hi], all_c[lo:hi], dst=0)  # one consumer: only rank 0 gets everything
        if rank == 0:
            ok = torch.equal(d, all_d) and torch.equal(c, all_c) and c.dtype == torch.int32
        else:
            ok = torch.equal(d, all_d[lo:hi]) and torch.equal(c, all_c[lo:hi])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_detections_to_root_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_root, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


class _FakeTileDetector:
    """Stands in for a DetectPipeline in the CPU tests of SahiSharder: deterministic detections derived from the tile's
    pixels only (so a tile yields the same boxes whichever rank and batch slot runs it)."""

    max_det, nc, conf, iou = 6, 3, 0.25, 0.5

    def __init__(self, B):
        self.B = B

    def __call__(self, tiles):
        assert tiles.shape[0] == self.B and tiles.dtype == torch.uint8
        B, K = self.B, self.max_det
        dets = torch.zeros(B, K, 6)
        counts = torch.zeros(B, dtype=torch.int32)
        for b in range(B):
            h = int(tiles[b].to(torch.int64).sum()) % 1000003
            g = torch.Generator().manual_seed(h)
            n = int(torch.randint(1, K + 1, (1,), generator=g))
            xy = torch.rand(n, 2, generator=g) * 40
            wh = torch.rand(n, 2, generator=g) * 20 + 4
            dets[b, :n, :2], dets[b, :n, 2:4] = xy, xy + wh
            dets[b, :n, 4] = torch.rand(n, generator=g).sort(descending=True).values * 0.7 + 0.3
            dets[b, :n, 5] = torch.randint(0, self.nc, (n,), generator=g).float()
            counts[b] = n
        return dets, counts


def _oracle_nms(pred, conf, iou, max_det):
    outs = oracle_post.non_max_suppression(pred.numpy().copy(), conf, iou, max_det=max_det)
    d = torch.zeros(len(outs), max_det, 6)
    for i, out in enumerate(outs):
        d[i, : len(out)] = torch.from_numpy(out)
    return d, torch.tensor([len(out) for out in outs], dtype=torch.int32)


def _sahi_frames(world):
    return [torch.from_numpy(np.random.default_rng(70 + r).integers(0, 256, (150, 230, 3), dtype=np.uint8)) for r in range(world)]


def _worker_sahi(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = shard.SahiSharder(_FakeTileDetector, (150, 230), 64, world=world, rank=rank, nms_fn=_oracle_nms)
        d, c = sh.predict_frame(_sahi_frames(world)[rank])
        q.put((rank, d.numpy().tobytes(), int(c), sh.B, sh.per))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sahi_sharder_exchange_equals_single_gpu_merge(world):
    """Tiles sharded over `world` ranks (all_to_all out, detections all_to_all back, merge at the frame owner) give, bit
    for bit, the single-process merge of every rank's frame (a 150x230 frame has 3 x 4 = 12 tiles of 64: 6 or 4 per rank;
    tile counts that do not divide are covered by test_sahi_sharder_pads_uneven_tile_counts)."""
    single = shard.SahiSharder(_FakeTileDetector, (150, 230), 64, world=1, rank=0, nms_fn=_oracle_nms)
    assert single.T == 12 and single.B == 12
    want = [single.predict_frame(f) for f in _sahi_frames(world)]
    assert all(int(c) > 0 for _, c in want)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sahi, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, blob, cnt, B, per in res:
        assert (B, per) == (12, 12 // world)
        assert cnt == int(want[rank][1]) and blob == want[rank][0].numpy().tobytes(), f"rank {rank}"


def test_sahi_sharder_pads_uneven_tile_counts():
    sh = shard.SahiSharder(_FakeTileDetector, (100, 170), 64, world=1, rank=0, nms_fn=_oracle_nms)
    assert sh.T == 6
    # 6 tiles over 4 ranks: 2 per rank, ranks 2 and 3 carry one padding slot each (never merged)
    sh4 = shard.SahiSharder(lambda B: _FakeTileDetector(B), (100, 170), 64, world=4, rank=0, nms_fn=_oracle_nms)
    assert (sh4.per, sh4.B) == (2, 8) and sh4.send_order == [0, 4, 1, 5, 2, -1, 3, -1]
    valid, org, order = sh4._consts(torch.device("cpu"))
    assert valid.tolist() == [True, True, True, True, True, False, True, False]
    assert [sh4.send_order[i] for i in order.tolist()] == [0, 1, 2, 3, 4, 5]   # tile order, padding slots dropped


def test_sahi_sharder_several_frames_per_step_equal_single_frames():
    frames = torch.stack(_sahi_frames(3))
    one = shard.SahiSharder(_FakeTileDetector, (150, 230), 64, world=1, rank=0, nms_fn=_oracle_nms)
    many = shard.SahiSharder(_FakeTileDetector, (150, 230), 64, world=1, rank=0, nms_fn=_oracle_nms, frames=3)
    assert many.B == 36
    d, c = many.predict_frames(frames)
    for f in range(3):
        d1, c1 = one.predict_frame(frames[f])
        assert int(c[f]) == int(c1) and torch.equal(d[f], d1)

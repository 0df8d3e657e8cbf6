"""Multi-GPU tests over NCCL (need >= 2 visible GPUs; skipped on a single-GPU box): the data-parallel paths really cross
GPUs only with final detections / SAHI tiles, and give bit-identical results to the single-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run with gpurun --gpus 2)")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph(model="yolo11n"):
    from bs_yolo_b200 import arch, ir
    spec = arch.build_spec(model)
    return ir.build_ir(spec, arch.synth_state_dict(spec, 0))


def _frames(world, hw=(300, 400)):
    return [torch.from_numpy(np.random.default_rng(80 + r).integers(0, 256, (*hw, 3), dtype=np.uint8)) for r in range(world)]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from bs_yolo_b200 import shard
    from bs_yolo_b200.pipeline import DetectPipeline
    from oracle import oracle_post

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _worker_body(rank, world, q, dist, shard, DetectPipeline, oracle_post)
    except BaseException as e:  # noqa: BLE001 -- report instead of leaving the parent waiting for the queue
        q.put((rank, f"{type(e).__name__}: {e}"))
        raise
    finally:
        dist.destroy_process_group()


def _worker_body(rank, world, q, dist, shard, DetectPipeline, oracle_post):
    if True:
        g = _graph()
        frames = _frames(world)
        mk = lambda B: DetectPipeline(g, B, 160, (160, 160), device=rank)  # noqa: E731
        # --- SAHI config 5 path: tiles sharded over the ranks, merge at the frame owner
        sh = shard.SahiSharder(mk, (300, 400), 160, world=world, rank=rank, iou=0.5)
        d, c = sh.predict_frame(frames[rank].cuda())
        torch.cuda.synchronize()
        single = shard.SahiSharder(mk, (300, 400), 160, world=1, rank=0, iou=0.5)
        tiles = single.slice_tiles(frames[rank].cuda().unsqueeze(0)).view(single.B, 160, 160, 3)
        tdets, tcounts = single.pipe(tiles)
        d1, c1 = single.predict_frame(frames[rank].cuda())
        torch.cuda.synchronize()
        ok_single = torch.equal(d, d1) and int(c) == int(c1)
        # ... and the merge equals, bit for bit, the oracle's NMS over the same per-tile detections
        org = torch.tensor(single.origins, dtype=torch.float32)
        pred = shard.merge_prediction(tdets.cpu(), tcounts.cpu(), org, single.nc)
        ref = oracle_post.non_max_suppression(pred.numpy(), single.conf, 0.5, max_det=single.max_det)[0]
        ok_oracle = int(c) == ref.shape[0] and d[: int(c)].cpu().numpy().tobytes() == ref.tobytes()
        # --- batch path: predict_stream with the detections gathered to rank 0 only
        pipe = mk(2)
        rng = np.random.default_rng(90 + rank)
        host = [torch.from_numpy(rng.integers(0, 256, (2, 160, 160, 3), dtype=np.uint8)).pin_memory() for _ in range(3)]
        own = [tuple(t.cpu().clone() for t in pipe(h.cuda())) for h in host]
        got = [(dd.clone(), cc.clone()) for dd, cc in pipe.predict_stream(iter(host), gather="root")]
        rows = [g_[0].shape[0] for g_ in got]
        ok_rows = rows == ([2 * world] * 3 if rank == 0 else [2] * 3)
        lo = 2 * rank if rank == 0 else 0
        ok_own = all(torch.equal(g_[0][lo:lo + 2], o[0]) and torch.equal(g_[1][lo:lo + 2], o[1]) for g_, o in zip(got, own))
        # rank 0 holds every rank's shard: compare against what the ranks computed themselves
        mine = torch.stack([o[0] for o in own]).cuda()
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        ok_all = True
        if rank == 0:
            for r in range(world):
                ok_all = ok_all and all(torch.equal(got[i][0][2 * r:2 * r + 2], allm[r][i].cpu()) for i in range(3))
        q.put((rank, bool(ok_single), bool(ok_oracle), bool(ok_rows), bool(ok_own), bool(ok_all), int(c), sh.B))


def test_sahi_and_gather_over_nccl_match_single_gpu():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=300))
            if len(res[-1]) != 8:   # a worker that raised reports (rank, message): its peer may be stuck in a collective
                break
    finally:
        if len(res) != world or any(len(r) != 8 for r in res):
            for p in procs:
                p.terminate()
    res = sorted(res)
    assert len(res) == world and all(len(r) == 8 for r in res), res
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok_single, ok_oracle, ok_rows, ok_own, ok_all, cnt, B in res:
        assert ok_single and ok_oracle and ok_rows and ok_own and ok_all, res
        assert cnt > 0 and B == 6   # a 300x400 frame has 2 x 3 tiles of 160: 3 per GPU and frame, 6 per step

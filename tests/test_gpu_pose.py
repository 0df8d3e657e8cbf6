"""Pose head on the B200 path, all through the C ABI:
  * b2y_gather_extras against the numpy restatement of Pose.kpts_decode + scale_coords (x / y bit-exact);
  * the keypoint branch of the CUDA-graphed plan against the reference's committed raw keypoint maps (bf16 gate 1e-2);
  * the engine's NMS + keypoint gather against the oracle's post-process run on the engine's OWN decoded tensors (exact);
  * the drop-in ``YOLO(pose ckpt).predict(predictor=B200PosePredictor)`` against the reference's PosePredictor results."""
import os

import numpy as np
import pytest
import torch

import kat
from bs_yolo_b200 import arch, ir, ops
from oracle import oracle_net, oracle_post, ref_import

pytestmark = pytest.mark.gpu


def _graph(model="yolo11n-pose", seed=0):
    spec = arch.build_spec(model)
    return ir.build_ir(spec, arch.synth_state_dict(spec, seed))


def _engine(g, x, **kw):
    from bs_yolo_b200.engine import Engine
    B, _, H, W = x.shape
    eng = Engine(g, B, H, W, **kw)
    eng.set_input(x)
    eng.forward()
    torch.cuda.synchronize()
    return eng


@pytest.mark.parametrize("ndim,with_geom,pitch", [(3, True, 52), (3, False, 51), (2, True, 36)])
def test_gather_extras_vs_oracle(ndim, with_geom, pitch):
    rng = np.random.default_rng(5 + ndim)
    B, max_det, nkpt = 3, 40, 17
    ne = nkpt * ndim
    shapes, strides = [(8, 12), (4, 6), (2, 3)], [8.0, 16.0, 32.0]
    An = sum(h * w for h, w in shapes)
    raws = [rng.standard_normal((B, h, w, pitch)).astype(np.float32) * 3 for h, w in shapes]
    counts = np.array([max_det, 0, 17], np.int32)
    kept = rng.integers(0, An, (B, max_det)).astype(np.int32)
    kept[0, :3] = [0, An - 1, shapes[0][0] * shapes[0][1]]  # first / last anchor, first anchor of level 1
    dets = (rng.uniform(-5, 300, (B, max_det, 6)) * 1.0).astype(np.float32)
    dets[0, 0, :4] = [0.5, 1.5, 2.5, -0.5]  # half-to-even: 0, 2, 2, -0
    img1, img0 = (64, 96), [(100, 150), (64, 96), (33, 200)]
    geom = ops.coords_geometry(img1, img0, "cuda") if with_geom else None
    d_dev = torch.from_numpy(dets.copy()).cuda()
    out = ops.gather_extras([torch.from_numpy(r).cuda() for r in raws], strides, ne, torch.from_numpy(kept).cuda(),
                            torch.from_numpy(counts).cuda(), mode=1, ndim=ndim, geom=geom, round_dets=d_dev).cpu().numpy()
    # oracle: decode every anchor like the reference, gather the kept rows, scale_coords
    kp = torch.cat([torch.from_numpy(r[..., :ne]).reshape(B, -1, ne).transpose(1, 2) for r in raws], 2)  # (B, ne, A)
    dec = oracle_net.kpts_decode(kp, shapes, strides, ndim).numpy()
    for b in range(B):
        n = int(counts[b])
        assert not out[b, n:].any()
        ref = dec[b][:, kept[b, :n]].T.reshape(n, nkpt, ndim)
        if with_geom:
            ref = oracle_post.scale_coords(img1, ref, img0[b])
        got = out[b, :n].reshape(n, nkpt, ndim)
        assert got[..., :2].tobytes() == ref[..., :2].tobytes()  # bit-exact x, y
        if ndim == 3:
            assert np.abs(got[..., 2] - ref[..., 2]).max(initial=0) < 1e-6
        exp = dets[b].copy()
        exp[:n, :4] = np.round(exp[:n, :4])
        assert np.array_equal(d_dev[b].cpu().numpy(), exp)
    # mode 0: plain gather of the ne channels
    cp = ops.gather_extras([torch.from_numpy(r).cuda() for r in raws], strides, ne, torch.from_numpy(kept).cuda(),
                           torch.from_numpy(counts).cuda(), mode=0).cpu().numpy()
    for b in range(B):
        n = int(counts[b])
        assert np.array_equal(cp[b, :n], kp.numpy()[b][:, kept[b, :n]].T) and not cp[b, n:].any()


def test_gather_extras_rejects_bad_arguments():
    raws = [torch.zeros(1, 2, 2, 6, device="cuda")]
    kept, counts = torch.zeros(1, 4, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        ops.gather_extras(raws, [8.0], 6, kept, counts, mode=1, ndim=4)
    with pytest.raises(ValueError):
        ops.gather_extras(raws, [8.0], 7, kept, counts, mode=0)  # ne > pitch
    with pytest.raises(Exception, match="CUDA"):
        ops.gather_extras([r.cpu() for r in raws], [8.0], 6, kept, counts)


@pytest.mark.parametrize("case", list(kat.POSE_NET_CASES))
def test_pose_branch_vs_reference_golden(case, golden_dir):
    gld = np.load(os.path.join(golden_dir, f"net_{case}.npz"))
    g = _graph(kat.POSE_NET_CASES[case][0])
    nc = g["nc"]
    eng = _engine(g, torch.from_numpy(kat.net_input(case)))
    assert eng.plan.extra_kind == "pose" and all(eng.op_impl(i) == 2 for i, o in enumerate(eng.plan.ops) if ".cv4." in o.name)
    got = torch.cat([eng.extra(l)[..., :51].reshape(eng.B, -1, 51).transpose(1, 2) for l in range(3)], 2).float().cpu()
    ref = torch.from_numpy(gld["kpt"])
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 1e-2, rel  # raw keypoint maps (B, nk, A), the second element of Pose.forward's tuple (head.py:250)
    assert tuple(eng.y.shape) == (eng.B, 4 + nc, ref.shape[2])  # the Detect part stays (B, 4+nc, A)


def test_engine_keypoints_equal_oracle_postprocess_on_own_tensors():
    """Given the engine's own decoded y and raw keypoint maps, NMS + kept-row gather + kpts_decode + scale_coords + box
    rounding must equal the oracle's pose post-process (which decodes every anchor and drags 51 rows through NMS) exactly:
    the only tolerance left is the visibility sigmoid."""
    g = _graph()
    nc = g["nc"]
    rng = np.random.default_rng(12)
    x = torch.from_numpy(rng.random((3, 3, 160, 128), dtype=np.float32))
    eng = _engine(g, x)
    img1, img0 = (160, 128), [(200, 150), (160, 128), (500, 300)]
    dets, counts, kept = eng.nms(0.25, 0.7, 300, return_anchors=True)
    ops.scale_boxes_batched(dets, counts, img1, img0)
    kp = eng.keypoints(kept, counts, geom=ops.coords_geometry(img1, img0, dets.device), round_dets=dets)
    torch.cuda.synchronize()
    raw = torch.cat([eng.extra(l)[..., :51].reshape(eng.B, -1, 51).transpose(1, 2) for l in range(3)], 2).float().cpu()
    dec = oracle_net.kpts_decode(raw, [(160 // s, 128 // s) for s in (8, 16, 32)], [8.0, 16.0, 32.0], 3)
    full = np.concatenate([eng.y.float().cpu().numpy(), dec.numpy()], 1)
    ref = oracle_post.pose_postprocess(full, nc, (17, 3), img1, img0)
    n = counts.tolist()
    assert n == [len(b) for b, _ in ref] and sum(n) > 3
    for i, (rb, rk) in enumerate(ref):
        assert np.array_equal(dets[i, : n[i]].cpu().numpy(), rb)
        got = kp[i, : n[i]].cpu().numpy()
        assert got.shape == rk.shape and got[..., :2].tobytes() == rk[..., :2].tobytes()
        assert np.abs(got[..., 2] - rk[..., 2]).max(initial=0) < 1e-6


def _pose_checkpoint(model, imgsz, tmp_path):
    ref_import.import_reference()
    spec = arch.build_spec(model)
    m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    m.kpt_shape = list(m.model[-1].kpt_shape)  # pose/train.py:49
    m.args = {"task": "pose", "imgsz": imgsz}
    path = str(tmp_path / f"{model}.pt")
    torch.save({"model": m, "ema": None, "train_args": {"task": "pose", "imgsz": imgsz}}, path)
    return path


@pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")
@pytest.mark.parametrize("case", list(kat.POSE_PREDICT_CASES))
def test_pose_dropin_vs_reference_golden(case, golden_dir, tmp_path):
    """``YOLO(pose ckpt).predict(frames, predictor=B200PosePredictor)`` vs the reference's PosePredictor on the CPU
    (committed golden).  Boxes are ROUNDED pixels on both sides (pose/predict.py:52), so the box gate is |d| <= 1 px and
    |dconf| <= 1e-2 per matched pair; keypoints of matched pairs: |dx|, |dy| <= 0.5 px, |dvis| <= 2e-2, and the
    "visibility < 0.5 -> (0, 0)" masking of Results.keypoints (results.py:1305-1309) must agree wherever the reference's
    visibility is not within 2e-2 of 0.5.  At least 90 % of the reference's rows must be matched outright; an unmatched
    row must sit within 1e-2 of the confidence threshold or be an NMS near-tie (counted, printed)."""
    ref_import.import_reference()
    from ultralytics import YOLO
    from bs_yolo_b200.predictor import B200PosePredictor

    model, n, h, w, imgsz, seed = kat.POSE_PREDICT_CASES[case]
    gld = np.load(os.path.join(golden_dir, f"predict_{case}.npz"))
    yolo = YOLO(_pose_checkpoint(model, imgsz, tmp_path))
    res = yolo.predict(kat.predict_input(case), predictor=B200PosePredictor, device=0, imgsz=imgsz, conf=0.25, iou=0.7,
                       max_det=300, verbose=False)
    assert type(yolo.predictor).__name__ == "B200PosePredictor" and len(res) == n
    total = matched = 0
    for i, r in enumerate(res):
        assert r.boxes.data.is_cuda and r.keypoints.data.is_cuda and r.orig_shape == (h, w)
        ours_b, ours_k = r.boxes.data.cpu().numpy(), r.keypoints.data.cpu().numpy()
        ref_b, ref_k = gld[f"boxes{i}"], gld[f"kpts{i}"]
        assert ours_k.shape[1:] == (17, 3) and ours_k.shape[0] == ours_b.shape[0]
        assert np.array_equal(ours_b[:, :4], np.round(ours_b[:, :4]))
        used = set()
        for j, rb in enumerate(ref_b):
            total += 1
            # the synthetic checkpoint emits many image-sized boxes that clip to the same rounded corners: pair by class,
            # corners within the rounding step, then the closest confidence
            best = min((k for k in range(len(ours_b)) if k not in used and ours_b[k, 5] == rb[5]
                        and np.abs(ours_b[k, :4] - rb[:4]).max() <= 1.0), key=lambda k: abs(ours_b[k, 4] - rb[4]), default=None)
            if best is None or abs(ours_b[best, 4] - rb[4]) > 1e-2:
                continue
            used.add(best)
            matched += 1
            ok, rk = ours_k[best], ref_k[j]
            assert np.abs(ok[:, 2] - rk[:, 2]).max(initial=0) <= 2e-2
            sure = np.abs(rk[:, 2] - 0.5) > 2e-2
            vis = sure & (rk[:, 2] >= 0.5)
            assert np.abs(ok[vis, :2] - rk[vis, :2]).max(initial=0) <= 0.5
            assert not ok[sure & ~vis, :2].any()
    print(f"pose drop-in {case}: {matched}/{total} reference rows matched")
    assert total > 0 and matched >= 0.9 * total


@pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")
def test_pose_dropin_empty_and_ragged(tmp_path):
    """No detections (conf 0.999) -> Results with (0, 6) boxes and empty keypoints like the reference's (predict.py:51);
    a ragged list takes LetterBox(auto=False) per image."""
    ref_import.import_reference()
    from ultralytics import YOLO
    from bs_yolo_b200.predictor import B200PosePredictor

    yolo = YOLO(_pose_checkpoint("yolo11n-pose", 160, tmp_path))
    rng = np.random.default_rng(2)
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((160, 160, 3), (120, 200, 3))]
    res = yolo.predict(imgs, predictor=B200PosePredictor, device=0, imgsz=160, conf=0.999, verbose=False)
    assert all(r.boxes.data.shape == (0, 6) and r.keypoints.data.numel() == 0 for r in res)
    res = yolo.predict(imgs, device=0, imgsz=160, conf=0.25, verbose=False)
    assert type(yolo.predictor).__name__ == "B200PosePredictor"
    assert all(r.keypoints.data.shape[0] == r.boxes.data.shape[0] for r in res) and sum(len(r.boxes) for r in res) > 0


@pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")
def test_head_and_predictor_class_must_match(tmp_path):
    """A pose checkpoint behind the detection predictor would silently drop the keypoints (the reference's own
    DetectionPredictor would read the keypoint rows as classes): it raises and names the class to use."""
    ref_import.import_reference()
    from ultralytics import YOLO
    from bs_yolo_b200.predictor import B200DetectionPredictor, B200SegmentationPredictor

    path = _pose_checkpoint("yolo11n-pose", 160, tmp_path)
    img = [np.zeros((160, 160, 3), np.uint8)]
    for cls in (B200DetectionPredictor, B200SegmentationPredictor):
        with pytest.raises(NotImplementedError, match="B200PosePredictor"):
            YOLO(path).predict(img, predictor=cls, device=0, imgsz=160, verbose=False)

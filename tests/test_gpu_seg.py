"""Segment head on the B200 path, all through the C ABI:
  * b2y_process_mask against the oracle's restatement of ops.process_mask (F.interpolate-based), incl. partial tiles;
  * Proto (ConvTranspose2d as 1x1 conv + depth-to-space) and the coefficient branch of the CUDA-graphed plan against the
    reference's committed prototypes / coefficients (bf16 gate 1e-2);
  * NMS + coefficient gather + masks against the oracle's post-process run on the engine's OWN tensors;
  * the drop-in ``YOLO(seg ckpt).predict(predictor=B200SegmentationPredictor)`` against the reference's results."""
import os

import numpy as np
import pytest
import torch

import kat
from bs_yolo_b200 import _cabi as A
from bs_yolo_b200 import arch, ir, ops
from oracle import oracle_post, ref_import

pytestmark = pytest.mark.gpu


def _graph(model="yolo11n-seg", seed=0):
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


def unpack_masks(g, i):
    shape = tuple(int(v) for v in g[f"mshape{i}"])
    return np.unpackbits(g[f"masks{i}"])[: int(np.prod(shape))].reshape(shape).astype(np.float32)


@pytest.mark.parametrize("H,W,nm,pitch", [(96, 160, 32, 32), (64, 64, 8, 16), (160, 128, 32, 40)])
def test_process_mask_vs_oracle(H, W, nm, pitch):
    rng = np.random.default_rng(H + W)
    B, max_det = 3, 12
    mh, mw = H // 4, W // 4
    protos = torch.from_numpy(rng.standard_normal((B, mh, mw, pitch)).astype(np.float32)).to(torch.bfloat16)
    coeffs = rng.standard_normal((B, max_det, nm)).astype(np.float32)
    counts = [max_det, 0, 5]
    dets = np.zeros((B, max_det, 6), np.float32)
    c, wh = rng.uniform(0, [W, H], (B, max_det, 2)), rng.uniform(8, [W, H], (B, max_det, 2))
    dets[..., :2], dets[..., 2:4] = c - wh / 2, c + wh / 2
    dets[0, 0, :4] = [0, 0, W, H]          # the whole canvas
    dets[0, 1, :4] = [-30, -30, 63.9, 500]  # leaves the canvas on three sides; x2 just short of a tile edge
    dets[0, 2, :4] = [40.0, 20.0, 44.0, 24.0]  # exactly one prototype cell
    out = ops.process_mask_batched(protos.cuda(), torch.from_numpy(coeffs).cuda(), torch.from_numpy(dets).cuda(), counts, (H, W))
    assert out[1] is None and out[0].shape == (max_det, H, W) and out[2].shape == (5, H, W)
    pf = protos.float().numpy()[..., :nm].transpose(0, 3, 1, 2)  # the kernel reads the bf16-rounded prototypes
    bad = tot = 0
    for b in (0, 2):
        ref = oracle_post.process_mask(pf[b], coeffs[b, : counts[b]], dets[b, : counts[b], :4], (H, W))
        got = out[b].cpu().numpy()
        assert set(np.unique(got)) <= {0.0, 1.0}
        bad += int((got != ref).sum())
        tot += ref.size
        assert np.array_equal(got[2], ref[2]) if b == 0 else True
    assert bad <= 1e-4 * tot, (bad, tot)  # only exact-zero crossings may differ (fp32 summation order)


def test_process_mask_rejects_bad_arguments():
    protos = torch.zeros(1, 8, 8, 12, dtype=torch.bfloat16, device="cuda")  # pitch 12: not a multiple of 8
    coeffs, dets = torch.zeros(1, 2, 8, device="cuda"), torch.zeros(1, 2, 6, device="cuda")
    with pytest.raises(NotImplementedError):
        ops.process_mask_batched(protos, coeffs, dets, [1], (32, 32))
    with pytest.raises(NotImplementedError):  # prototypes must be the network input / 4
        ops.process_mask_batched(torch.zeros(1, 8, 8, 8, dtype=torch.bfloat16, device="cuda"), coeffs, dets, [1], (64, 32))
    with pytest.raises(Exception, match="CUDA"):
        ops.process_mask_batched(protos.cpu(), coeffs, dets, [1], (32, 32))


@pytest.mark.parametrize("case", list(kat.SEG_NET_CASES))
def test_segment_branches_vs_reference_golden(case, golden_dir):
    gld = np.load(os.path.join(golden_dir, f"net_{case}.npz"))
    g = _graph(kat.SEG_NET_CASES[case][0])
    eng = _engine(g, torch.from_numpy(kat.net_input(case)))
    assert eng.plan.extra_kind == "segment" and any(o.kind == A.OP_D2S for o in eng.plan.ops)
    assert all(eng.op_impl(i) == 2 for i, o in enumerate(eng.plan.ops) if o.kind == A.OP_CONV and (".cv4." in o.name or ".proto." in o.name))
    got = eng.protos.permute(0, 3, 1, 2).float().cpu()
    ref = torch.from_numpy(gld["proto"])
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 1e-2, rel  # prototypes (B, nm, H/4, W/4): Proto incl. the ConvTranspose2d lowering
    mc = torch.cat([eng.extra(l)[..., :32].reshape(eng.B, -1, 32).transpose(1, 2) for l in range(3)], 2).float().cpu()
    rmc = torch.from_numpy(gld["mc"])
    assert float((mc - rmc).norm() / rmc.norm()) < 1e-2


def test_engine_masks_equal_oracle_postprocess_on_own_tensors():
    """Given the engine's own y, coefficient maps and prototypes: NMS + kept-row gather + b2y_process_mask + scale_boxes
    must equal the oracle's segment post-process (boxes bit-exact; masks up to exact-zero crossings)."""
    g = _graph()
    nc = g["nc"]
    x = torch.from_numpy(np.random.default_rng(13).random((3, 3, 160, 128), dtype=np.float32))
    eng = _engine(g, x)
    img1, img0 = (160, 128), [(200, 150), (160, 128), (500, 300)]
    dets, counts, kept = eng.nms(0.25, 0.7, 300, return_anchors=True)
    coeffs = eng.mask_coeffs(kept, counts)
    n = counts.tolist()
    masks = ops.process_mask_batched(eng.protos, coeffs, dets, n, img1)
    ops.scale_boxes_batched(dets, counts, img1, img0)
    torch.cuda.synchronize()
    mc = torch.cat([eng.extra(l)[..., :32].reshape(eng.B, -1, 32).transpose(1, 2) for l in range(3)], 2).float().cpu().numpy()
    full = np.concatenate([eng.y.float().cpu().numpy(), mc], 1)
    ref = oracle_post.segment_postprocess(full, eng.protos.permute(0, 3, 1, 2).float().cpu().numpy(), nc, img1, img0)
    assert n == [len(b) for b, _ in ref] and sum(n) > 3
    bad = tot = 0
    for i, (rb, rm) in enumerate(ref):
        assert np.array_equal(dets[i, : n[i]].cpu().numpy(), rb)
        assert np.array_equal(coeffs[i, : n[i]].cpu().numpy(), mc[i][:, kept[i, : n[i]].cpu().numpy()].T)
        if n[i]:
            got = masks[i].cpu().numpy()
            bad += int((got != rm).sum())
            tot += rm.size
        else:
            assert masks[i] is None and rm is None
    assert bad <= 1e-4 * tot, (bad, tot)


def _seg_checkpoint(model, imgsz, tmp_path):
    ref_import.import_reference()
    spec = arch.build_spec(model)
    m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    m.args = {"task": "segment", "imgsz": imgsz}
    path = str(tmp_path / f"{model}.pt")
    torch.save({"model": m, "ema": None, "train_args": {"task": "segment", "imgsz": imgsz}}, path)
    return path


@pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")
@pytest.mark.parametrize("case", list(kat.SEG_PREDICT_CASES))
def test_segment_dropin_vs_reference_golden(case, golden_dir, tmp_path):
    """``YOLO(seg ckpt).predict(frames, predictor=B200SegmentationPredictor)`` vs the reference's SegmentationPredictor on
    the CPU (committed golden).  Rows are paired by class, corners within 1 px and the closest confidence (|dconf| <=
    1e-2); at least 90 % of the reference's rows must pair.  Masks of paired rows: IoU >= 0.95 each and >= 0.98 on average
    -- the synthetic checkpoint's masks are sign patterns of noise, so bf16 moves a band of zero crossings (the bf16
    emulation of the plan on the CPU gives IoU 0.979 .. 1.0 against the same golden)."""
    ref_import.import_reference()
    from ultralytics import YOLO
    from bs_yolo_b200.predictor import B200SegmentationPredictor

    model, n, h, w, imgsz, seed = kat.SEG_PREDICT_CASES[case]
    gld = np.load(os.path.join(golden_dir, f"predict_{case}.npz"))
    yolo = YOLO(_seg_checkpoint(model, imgsz, tmp_path))
    res = yolo.predict(kat.predict_input(case), predictor=B200SegmentationPredictor, device=0, imgsz=imgsz, conf=0.25, iou=0.7,
                       max_det=300, verbose=False)
    assert type(yolo.predictor).__name__ == "B200SegmentationPredictor" and len(res) == n
    total = matched = 0
    ious = []
    for i, r in enumerate(res):
        assert r.boxes.data.is_cuda and r.masks.data.is_cuda and r.orig_shape == (h, w)
        ob, om = r.boxes.data.cpu().numpy(), r.masks.data.cpu().numpy()
        rb, rm = gld[f"boxes{i}"], unpack_masks(gld, i)
        assert om.shape[1:] == rm.shape[1:] and om.shape[0] == ob.shape[0] and set(np.unique(om)) <= {0.0, 1.0}
        used = set()
        for j, ref in enumerate(rb):
            total += 1
            k = min((k for k in range(len(ob)) if k not in used and ob[k, 5] == ref[5] and np.abs(ob[k, :4] - ref[:4]).max() <= 1.0),
                    key=lambda k: abs(ob[k, 4] - ref[4]), default=None)
            if k is None or abs(ob[k, 4] - ref[4]) > 1e-2:
                continue
            used.add(k)
            matched += 1
            inter, union = float((om[k] * rm[j]).sum()), float(((om[k] + rm[j]) > 0).sum())
            ious.append(inter / max(union, 1.0))
    print(f"segment drop-in {case}: {matched}/{total} reference rows paired, mask IoU min {min(ious):.4f} mean {np.mean(ious):.4f}")
    assert total > 0 and matched >= 0.9 * total
    assert min(ious) >= 0.95 and float(np.mean(ious)) >= 0.98


@pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")
def test_segment_dropin_empty_and_rejections(tmp_path):
    ref_import.import_reference()
    from ultralytics import YOLO
    from bs_yolo_b200.predictor import B200SegmentationPredictor

    yolo = YOLO(_seg_checkpoint("yolo11n-seg", 160, tmp_path))
    rng = np.random.default_rng(2)
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((160, 160, 3), (120, 200, 3))]
    res = yolo.predict(imgs, predictor=B200SegmentationPredictor, device=0, imgsz=160, conf=0.999, verbose=False)
    assert all(r.boxes.data.shape == (0, 6) and r.masks is None for r in res)  # segment/predict.py:49-50
    res = yolo.predict(imgs, device=0, imgsz=160, conf=0.25, verbose=False)  # ragged list: LetterBox(auto=False) per image
    assert all(r.masks.data.shape == (len(r.boxes), 160, 160) for r in res) and sum(len(r.boxes) for r in res) > 0
    with pytest.raises(NotImplementedError):
        yolo.predict(imgs, device=0, imgsz=160, retina_masks=True, verbose=False)

"""Pin the ORACLE against the reference's own outputs (tests/golden, made by make_golden.py from the
unmodified reference).  Runs on CPU everywhere; no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

import kat
from bs_yolo_b200 import arch, ir
from oracle import oracle_net, oracle_post


@pytest.mark.parametrize("case", list(kat.NET_CASES))
def test_net_forward_matches_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, f"net_{case}.npz"))
    model = kat.NET_CASES[case][0]
    spec = arch.build_spec(model)
    sd = arch.synth_state_dict(spec, 0)
    assert kat.state_dict_checksum(sd) == pytest.approx(float(g["sd_checksum"]), rel=1e-12), "synthetic weights drifted"
    graph = ir.build_ir(spec, sd)
    y, raws = oracle_net.forward(graph, torch.from_numpy(kat.net_input(case)))
    for i, r in enumerate(raws):
        ref = torch.from_numpy(g[f"raw{i}"])
        rel = float((r - ref).norm() / ref.norm())
        assert rel < 1e-5, (i, rel)  # fp32; only the BN-fold association order differs
    ref = torch.from_numpy(g["y"])
    assert float((y[:, :4] - ref[:, :4]).abs().max()) < 2e-3  # pixels
    assert float((y[:, 4:] - ref[:, 4:]).abs().max()) < 1e-5


@pytest.mark.parametrize("nms_fn", ["c", "numpy"])
@pytest.mark.parametrize("case", kat.NMS_CASES)
def test_nms_bit_exact_vs_reference(case, nms_fn, golden_dir):
    if nms_fn == "numpy" and case in ("dense_8400", "many_kept", "nc12_conf_low"):
        pytest.skip("numpy variant is O(n) python iterations; covered by the C variant")
    g = np.load(os.path.join(golden_dir, "nms.npz"))
    p, kw = kat.nms_case(case)
    assert float(p.astype(np.float64).sum()) == float(g[f"{case}__insum"]), "KAT input drifted"
    fn = oracle_post.nms_c if nms_fn == "c" else oracle_post.nms_numpy
    out = oracle_post.non_max_suppression(p, nms_fn=fn, **kw)
    assert [o.shape[0] for o in out] == g[f"{case}__n"].tolist()
    rows = np.concatenate([o.reshape(-1, 6) for o in out], 0).astype(np.float32)
    assert rows.tobytes() == g[f"{case}__rows"].tobytes()  # bit-exact


@pytest.mark.parametrize("case", list(kat.LETTERBOX_CASES))
@pytest.mark.parametrize("auto", [False, True])
def test_letterbox_bit_exact_vs_reference(case, auto, golden_dir):
    g = np.load(os.path.join(golden_dir, "letterbox.npz"))
    imgsz = kat.LETTERBOX_CASES[case][2]
    out = oracle_post.letterbox(kat.letterbox_input(case), (imgsz, imgsz), auto=auto, stride=32)
    ref = g[f"{case}__auto{int(auto)}"]
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)


def test_resize_emulation_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for (h, w, dw, dh) in [(540, 960, 640, 360), (720, 1280, 640, 360), (481, 641, 640, 480), (37, 1000, 640, 24)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(oracle_post.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("case", list(kat.PREDICT_CASES))
def test_predict_end_to_end_matches_reference(case, golden_dir):
    """oracle preprocess -> forward -> NMS -> scale_boxes == YOLO(ckpt).predict() of the reference."""
    g = np.load(os.path.join(golden_dir, f"predict_{case}.npz"))
    model, n, h, w, imgsz, _ = kat.PREDICT_CASES[case]
    spec = arch.build_spec(model)
    graph = ir.build_ir(spec, arch.synth_state_dict(spec, 0))
    imgs = kat.predict_input(case)
    x = oracle_post.preprocess(imgs, (imgsz, imgsz), stride=32)
    y, _ = oracle_net.forward(graph, torch.from_numpy(x))
    dets = oracle_post.non_max_suppression(y.numpy(), 0.25, 0.7, max_det=300)
    assert [d.shape[0] for d in dets] == g["n"].tolist()
    for i, d in enumerate(dets):
        d = oracle_post.scale_boxes(x.shape[2:], d, imgs[i].shape[:2])
        ref = g[f"boxes{i}"]
        assert np.array_equal(d[:, 5], ref[:, 5])
        assert np.abs(d[:, :4] - ref[:, :4]).max() < 2e-2  # pixels (fp32 fold-order noise through the net)
        assert np.abs(d[:, 4] - ref[:, 4]).max() < 1e-4


# ---- validator path ------------------------------------------------------------------------------------
from oracle import oracle_val  # noqa: E402


@pytest.mark.parametrize("case", list(kat.VAL_MATCH_CASES))
def test_val_box_iou_and_match_predictions_bit_exact_vs_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, "val_match.npz"))
    det, gt, gcls = kat.val_match_case(case)
    insum = float(det.astype(np.float64).sum() + gt.astype(np.float64).sum() + gcls.astype(np.float64).sum())
    assert insum == float(g[f"{case}__insum"]), "KAT input drifted"
    assert np.array_equal(oracle_val.IOUV, g["iouv"])  # == torch.linspace(0.5, 0.95, 10), val.py:41
    iou = oracle_val.box_iou(gt, det[:, :4])
    assert iou.shape == g[f"{case}__iou"].shape and iou.tobytes() == g[f"{case}__iou"].tobytes()  # bit-exact
    tp = oracle_val.process_batch(det, gt, gcls)
    assert np.array_equal(tp, g[f"{case}__tp"])
    assert np.array_equal(oracle_val.match_by_definition(det, gt, gcls), g[f"{case}__tp"])  # the kernel's formulation


def test_val_match_formulations_agree_on_random_inputs():
    """match_predictions as the reference runs it (sort + two np.unique) == best label per detection, then the lowest-index
    detection of every label -- the form the CUDA kernel evaluates -- on many random crowded scenes."""
    rng = np.random.default_rng(7)
    for _ in range(60):
        m, n, nc = int(rng.integers(1, 30)), int(rng.integers(1, 120)), int(rng.integers(1, 4))
        c, wh = rng.uniform(50, 300, (m, 2)), rng.uniform(20, 90, (m, 2))
        gt = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        gcls = rng.integers(0, nc, m).astype(np.float32)
        src = rng.integers(0, m, n)
        box = gt[src] + rng.normal(0, 6, (n, 4)).astype(np.float32)
        det = np.concatenate([box, np.sort(rng.random((n, 1)).astype(np.float32), 0)[::-1],
                              np.where(rng.random(n) < 0.85, gcls[src], 0).astype(np.float32)[:, None]], 1).astype(np.float32)
        assert np.array_equal(oracle_val.process_batch(det, gt, gcls), oracle_val.match_by_definition(det, gt, gcls))

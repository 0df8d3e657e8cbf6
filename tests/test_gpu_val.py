"""Validator path on the B200, through the C ABI: b2y_box_iou and b2y_match_predictions bit-exact against
the reference's own outputs (tests/golden/val_match.npz) and the oracle; b2y_preprocess_chw exact; and the drop-in
``YOLO.val(validator=B200DetectionValidator)`` against the reference validating on the CPU in this process."""
import os

import numpy as np
import pytest
import torch

import kat
from bs_yolo_b200 import _cabi as A
from bs_yolo_b200 import arch, ops
from oracle import oracle_val, ref_import

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", list(kat.VAL_MATCH_CASES))
def test_box_iou_bit_exact_vs_reference_golden(case, golden_dir):
    g = np.load(os.path.join(golden_dir, "val_match.npz"))
    det, gt, _ = kat.val_match_case(case)
    iou = ops.box_iou(_dev(gt), _dev(det[:, :4])).cpu().numpy()
    assert iou.shape == g[f"{case}__iou"].shape and iou.tobytes() == g[f"{case}__iou"].tobytes()


def _pack(cases, max_det, max_labels):
    B = len(cases)
    dets = np.zeros((B, max_det, 6), np.float32)
    counts = np.zeros(B, np.int32)
    rows = np.zeros((B, max_labels, 5), np.float32)
    nl = np.zeros(B, np.int32)
    for b, (det, gt, gcls) in enumerate(cases):
        dets[b, : len(det)] = det
        counts[b] = len(det)
        rows[b, : len(gt), 0] = gcls
        rows[b, : len(gt), 1:] = gt
        nl[b] = len(gt)
    return dets, counts, rows, nl


def test_match_predictions_bit_exact_vs_reference_golden(golden_dir):
    """All KAT cases as ONE padded batch (how the validator calls it) and one by one at their exact sizes."""
    g = np.load(os.path.join(golden_dir, "val_match.npz"))
    names = list(kat.VAL_MATCH_CASES)
    cases = [kat.val_match_case(c) for c in names]
    dets, counts, rows, nl = _pack(cases, 300, 150)
    tp = ops.match_predictions_batched(_dev(dets), _dev(counts), _dev(rows), _dev(nl), g["iouv"]).cpu().numpy()
    for b, c in enumerate(names):
        ref = g[f"{c}__tp"]
        assert np.array_equal(tp[b, : counts[b]], ref), c
        assert not tp[b, counts[b]:].any(), c
    for c, (det, gt, gcls) in zip(names, cases):
        d1, c1, r1, n1 = _pack([(det, gt, gcls)], max(1, len(det)), len(gt))
        tp1 = ops.match_predictions_batched(_dev(d1), _dev(c1), _dev(r1.reshape(1, len(gt), 5)), _dev(n1), g["iouv"])
        assert np.array_equal(tp1.cpu().numpy()[0, : len(det)], g[f"{c}__tp"]), c


def test_match_predictions_random_batch_vs_oracle():
    """A validator-sized batch (64 images x 300 detections x up to 200 labels) against the oracle's restatement of
    box_iou + match_predictions, bit-exact."""
    rng = np.random.default_rng(5)
    cases = []
    for _ in range(64):
        m, n, nc = int(rng.integers(0, 200)), int(rng.integers(0, 301)), int(rng.integers(1, 5))
        c, wh = rng.uniform(40, 600, (m, 2)), rng.uniform(10, 200, (m, 2))
        gt = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        gcls = rng.integers(0, nc, m).astype(np.float32)
        if m:
            src = rng.integers(0, m, n)
            box = gt[src] + (rng.normal(0, 0.05, (n, 4)) * np.tile(wh[src], 2)).astype(np.float32)
            dcls = np.where(rng.random(n) < 0.8, gcls[src], rng.integers(0, nc, n)).astype(np.float32)
        else:
            box, dcls = rng.uniform(0, 640, (n, 4)).astype(np.float32), np.zeros(n, np.float32)
        conf = np.sort(rng.random(n).astype(np.float32))[::-1]
        cases.append((np.concatenate([box, conf[:, None], dcls[:, None]], 1).astype(np.float32), gt, gcls))
    dets, counts, rows, nl = _pack(cases, 300, 200)
    tp = ops.match_predictions_batched(_dev(dets), _dev(counts), _dev(rows), _dev(nl), oracle_val.IOUV).cpu().numpy()
    total = 0
    for b, (det, gt, gcls) in enumerate(cases):
        ref = oracle_val.process_batch(det, gt, gcls)
        assert np.array_equal(tp[b, : len(det)], ref), b
        total += int(ref.sum())
    assert total > 2000  # the scenes do produce matches at many thresholds


def test_match_predictions_rejects_bad_arguments():
    d, c = torch.zeros(1, 4, 6, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    r, n = torch.zeros(1, 2, 5, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        ops.match_predictions_batched(d, c, r, n, [0.5, 0.0])
    with pytest.raises(NotImplementedError):
        ops.match_predictions_batched(d, c, r, n, [0.5] * 17)
    with pytest.raises(A.B2YError, match="no CPU fallback"):
        ops.box_iou(torch.zeros(1, 4), torch.zeros(1, 4))


@pytest.mark.parametrize("shape", [(3, 64, 96), (2, 160, 128), (1, 32, 4)])
def test_preprocess_chw_exact(shape):
    B, H, W = shape
    img = np.random.default_rng(H).integers(0, 256, (B, 3, H, W), dtype=np.uint8)
    img[0, :, 0, :4] = [[0, 1, 254, 255]] * 3
    dst = torch.empty(B, H // 2, W // 2, 16, dtype=torch.bfloat16, device="cuda")
    ops.preprocess_chw(_dev(img), dst)
    assert np.array_equal(dst.float().cpu().numpy(), oracle_val.preprocess_chw(img))
    with pytest.raises(NotImplementedError):
        ops.preprocess_chw(_dev(img[..., :-2]), torch.empty(B, H // 2, (W - 2) // 2, 16, dtype=torch.bfloat16, device="cuda"))


@pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")
@pytest.mark.parametrize("model", ["yolo11n", "bsyolo_n"])
def test_val_dropin_vs_reference(model, tmp_path):
    """``YOLO.val(validator=B200DetectionValidator)`` against the reference's CPU validator on a synthetic dataset with
    rect batches.  Gates: (1) the true-positive matrix is bit-exact given OUR predictions (oracle on the captured
    original-image boxes and labels), (2) the label bookkeeping (instances per class, images seen) is identical, (3) the
    metrics agree with the reference's within the bf16 budget (the two sides' predictions differ by bf16 rounding)."""
    ref_import.import_reference()
    from ultralytics import YOLO

    import valdata
    from bs_yolo_b200.validator import B200DetectionValidator

    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    open(os.path.join(os.environ["YOLO_CONFIG_DIR"], "Arial.ttf"), "a").close()  # check_font would try to download it
    spec = arch.build_spec(model)
    m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    m.args = {"task": "detect", "imgsz": 160}
    path = str(tmp_path / f"{model}.pt")
    torch.save({"model": m, "ema": None, "train_args": {"task": "detect", "imgsz": 160}}, path)
    data = valdata.make_dataset(str(tmp_path / "ds"), YOLO(path), spec["nc"], [(160, 160), (120, 200), (240, 180)], 10, 160, seed=3)
    kw = dict(data=data, imgsz=160, batch=4, workers=0, plots=False, verbose=False, project=str(tmp_path / "runs"))
    ref = YOLO(path).val(device="cpu", **kw)

    captured = []

    class Capturing(B200DetectionValidator):
        capture = captured

    yolo = YOLO(path)
    ours = yolo.val(validator=Capturing, device=0, **kw)
    assert len(captured) == 3 and sum(len(c[2]) for c in captured) == 10
    n_tp = 0
    for predn, tp, counts, per_image in captured:
        for si, n in enumerate(counts):
            cls, bbox = per_image[si]
            ref_tp = oracle_val.process_batch(predn[si, :n].numpy(), bbox.numpy().reshape(-1, 4), cls.numpy())
            assert np.array_equal(tp[si, :n].numpy(), ref_tp)
            assert not tp[si, n:].any()
            n_tp += int(ref_tp[:, 0].sum())
    assert n_tp > 0
    assert ours.box.nc == ref.box.nc and list(ours.ap_class_index) == list(ref.ap_class_index)
    print(f"val {model}: reference {ref.results_dict}\n          ours      {ours.results_dict}")
    for k in ("metrics/mAP50(B)", "metrics/mAP50-95(B)", "metrics/recall(B)"):
        assert abs(ours.results_dict[k] - ref.results_dict[k]) <= 0.03, (k, ours.results_dict[k], ref.results_dict[k])
    assert ref.results_dict["metrics/mAP50(B)"] > 0.05  # the dataset does exercise the matcher


@pytest.mark.skipif(not ref_import.available(), reason="reference package not importable")
def test_val_dropin_rejects_what_it_does_not_run(tmp_path):
    ref_import.import_reference()
    from bs_yolo_b200.validator import B200DetectionValidator

    v = B200DetectionValidator(args=dict(model="x.pt", data="x.yaml", half=True, save_dir=str(tmp_path), project=str(tmp_path)))
    with pytest.raises(NotImplementedError):
        v(model=None)
    with pytest.raises(NotImplementedError):
        B200DetectionValidator(args=dict(model="x.pt", data="x.yaml", project=str(tmp_path)))(trainer=object())

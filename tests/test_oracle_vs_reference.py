"""In-process check of spec/IR/oracle against the live reference (build container only; skipped on the GPU box)."""
import numpy as np
import pytest
import torch

from bs_yolo_b200 import arch, ir
from oracle import oracle_net, oracle_post, ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


@pytest.mark.parametrize("model", ["yolo11n", "bsyolo_n"])
def test_spec_and_forward_match_live_reference(model):
    spec = arch.build_spec(model)
    sd = arch.synth_state_dict(spec, 3)
    m = ref_import.reference_model(model, sd)  # strict load: parameter names/shapes identical to the reference's
    assert ir.strip_tensors(ir.spec_from_model(m)) == ir.strip_tensors(spec)
    g = ir.ir_from_model(m)
    x = torch.rand(1, 3, 96, 64)
    with torch.no_grad():
        yr, rr = m(x)
    y, raws = oracle_net.forward(g, x)
    for a, b in zip(raws, rr):
        assert float((a - b).norm() / b.norm()) < 1e-5
    assert float((y - yr).abs().max()) < 2e-3


def test_fused_reference_model_is_handled():
    """model.fuse() removes .bn from the modules it recognises (tasks.py:200-228); the IR must be unchanged."""
    spec = arch.build_spec("yolo11n")
    m = ref_import.reference_model("yolo11n", arch.synth_state_dict(spec, 4))
    g0 = ir.ir_from_model(m)
    m.fuse(verbose=False)
    g1 = ir.ir_from_model(m)
    x = torch.rand(1, 3, 64, 64)
    y0, _ = oracle_net.forward(g0, x)
    y1, _ = oracle_net.forward(g1, x)
    assert float((y0 - y1).abs().max()) < 1e-3


def test_scale_boxes_matches_reference():
    ref_import.import_reference()
    from ultralytics.utils import ops

    rng = np.random.default_rng(0)
    b = (rng.random((50, 6)) * 700 - 30).astype(np.float32)
    for s1, s0 in [((640, 640), (640, 640)), ((384, 640), (1080, 1920)), ((160, 128), (481, 321))]:
        ref = ops.scale_boxes(s1, torch.from_numpy(b.copy()), s0).numpy()
        assert np.array_equal(oracle_post.scale_boxes(s1, b, s0), ref)


def test_dropin_predictor_plugs_into_predict_and_refuses_cpu(tmp_path):
    """The plug point (engine/model.py:549-551) accepts the class; without a CUDA device it raises instead of falling back."""
    ref_import.import_reference()
    from ultralytics import YOLO
    from ultralytics.models.yolo.detect.predict import DetectionPredictor

    from bs_yolo_b200 import _cabi as A
    from bs_yolo_b200.predictor import B200DetectionPredictor

    assert issubclass(B200DetectionPredictor, DetectionPredictor)
    spec = arch.build_spec("yolo11n")
    m = ref_import.reference_model("yolo11n", arch.synth_state_dict(spec, 0))
    m.names = {k: f"class{k}" for k in range(spec["nc"])}
    m.args = {"task": "detect", "imgsz": 64}
    path = str(tmp_path / "y.pt")
    torch.save({"model": m, "ema": None, "train_args": {"task": "detect", "imgsz": 64}}, path)
    with pytest.raises(A.B2YError, match="no CPU fallback"):
        YOLO(path).predict([np.zeros((64, 64, 3), np.uint8)], predictor=B200DetectionPredictor, device="cpu", imgsz=64,
                           verbose=False)


def test_validator_label_rows_match_reference_prepare_batch():
    """Host logic of the drop-in validator (bs_yolo_b200/validator.py:_label_rows) and the oracle's prepare_labels against
    the reference's own DetectionValidator._prepare_batch (detect/val.py:104-115) on a rect-style batch: bit-exact."""
    ref_import.import_reference()
    from ultralytics.models.yolo.detect.val import DetectionValidator

    from bs_yolo_b200.validator import _label_rows
    from oracle import oracle_val

    rng = np.random.default_rng(11)
    nb, H, W = 5, 128, 160
    ori = [(240, 320), (97, 133), (128, 160), (300, 333), (64, 400)]
    ratio_pad, counts = [], [7, 0, 3, 12, 1]
    for (h0, w0) in ori:
        r = min(H / h0, W / w0)
        nh, nw = int(round(h0 * r)), int(round(w0 * r))
        ratio_pad.append(((nh / h0, nw / w0), (round((W - nw) / 2 - 0.1), round((H - nh) / 2 - 0.1))))
    idx = np.concatenate([np.full(c, i) for i, c in enumerate(counts)]).astype(np.float32)
    n = int(sum(counts))
    cxy, wh = rng.uniform(0.05, 0.95, (n, 2)), rng.uniform(0.02, 0.6, (n, 2))
    batch = {"img": torch.zeros(nb, 3, H, W, dtype=torch.uint8), "batch_idx": torch.from_numpy(idx),
             "cls": torch.from_numpy(rng.integers(0, 12, (n, 1)).astype(np.float32)),
             "bboxes": torch.from_numpy(np.concatenate([cxy, wh], 1).astype(np.float32)),
             "ori_shape": ori, "ratio_pad": ratio_pad}
    v = object.__new__(DetectionValidator)
    v.device = torch.device("cpu")
    rows, nl, per_image = _label_rows(batch, nb, (H, W), torch.device("cpu"))
    assert nl.tolist() == counts and rows.shape == (nb, max(counts), 5)
    for si in range(nb):
        ref = v._prepare_batch(si, batch)
        cls, bbox = per_image[si]
        assert torch.equal(cls, ref["cls"]) and bbox.shape == ref["bbox"].shape
        if counts[si]:
            assert bbox.numpy().tobytes() == ref["bbox"].numpy().tobytes()
            assert torch.equal(rows[si, : counts[si], 1:], ref["bbox"]) and torch.equal(rows[si, : counts[si], 0], ref["cls"])
            sel = idx == si
            mine = oracle_val.prepare_labels(batch["bboxes"].numpy()[sel], (H, W), ori[si], ratio_pad[si])
            assert mine.tobytes() == ref["bbox"].numpy().tobytes()
        assert float(rows[si, counts[si]:].abs().sum()) == 0.0


def test_prefetch_image_loader_matches_reference_loader(tmp_path):
    """ingest.PrefetchImageLoader against the reference's LoadImagesAndVideos (data/loaders.py:284-449) on a
    directory with an unreadable file and one odd-shaped image: same batches, paths, info strings, pixels and counts; the
    pinned block of a same-shape batch holds the same pixels in order, a ragged batch has none."""
    cv2 = pytest.importorskip("cv2")
    ref_import.import_reference()
    from ultralytics.data.loaders import LoadImagesAndVideos

    from bs_yolo_b200.ingest import PinnedFrames, PrefetchImageLoader

    rng = np.random.default_rng(2)
    for i in range(11):
        shape = (50, 70, 3) if i == 9 else (48, 64, 3)
        cv2.imwrite(str(tmp_path / f"im{i:02d}.png"), rng.integers(0, 256, shape, dtype=np.uint8))
    (tmp_path / "im03b.png").write_bytes(b"not a png at all")
    ref = LoadImagesAndVideos(str(tmp_path), batch=4)
    ours = PrefetchImageLoader(LoadImagesAndVideos(str(tmp_path), batch=4))
    assert len(ours) == len(ref) and ours.bs == ref.bs and ours.files == ref.files and ours.mode == "image"
    for _ in range(2):  # iterable twice, like the reference's
        it_r, it_o, nb = iter(ref), iter(ours), 0
        while True:
            try:
                pr, ir_, sr = next(it_r)
            except StopIteration:
                with pytest.raises(StopIteration):
                    next(it_o)
                break
            po, io, so = next(it_o)
            nb += 1
            assert po == pr and so == sr and ours.count == ref.count and isinstance(io, PinnedFrames) and len(io) == len(ir_)
            assert all(np.array_equal(a, b) for a, b in zip(io, ir_))
            if len({a.shape for a in ir_}) == 1:
                assert io.pinned is not None and np.array_equal(io.pinned.numpy()[: len(io)], np.stack(ir_))
            else:
                assert io.pinned is None
        assert nb == 3
    class VideoLoader:  # a loader over a video file keeps the reference's sequential decoder
        video_flag, files, nf, ni, bs, vid_stride = [True], ["a.mp4"], 1, 0, 2, 1

    with pytest.raises(ValueError):
        PrefetchImageLoader(VideoLoader())


# ---- Pose / Segment heads ----------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["yolo11n-pose", "yolo11n-seg"])
def test_head_spec_and_forward_match_live_reference(model):
    spec = arch.build_spec(model)
    sd = arch.synth_state_dict(spec, 5)
    m = ref_import.reference_model(model, sd)  # PoseModel / SegmentationModel, strict load
    assert ir.strip_tensors(ir.spec_from_model(m)) == ir.strip_tensors(spec)
    g = ir.ir_from_model(m)
    x = torch.rand(1, 3, 96, 64)
    with torch.no_grad():
        yr, extra_r = m(x)
    y, extra = oracle_net.forward(g, x)
    assert y.shape == yr.shape and float((y[:, :4] - yr[:, :4]).abs().max()) < 2e-3
    assert float((y[:, 4:84] - yr[:, 4:84]).abs().max()) < 1e-5
    for a, b in zip(extra[1:], extra_r[1:]):  # raw keypoint maps | (mask coefficients, prototypes)
        assert a.shape == b.shape and float((a - b).norm() / b.norm()) < 1e-5


def test_scale_coords_and_process_mask_match_reference_bit_for_bit():
    ref_import.import_reference()
    from ultralytics.utils import ops

    rng = np.random.default_rng(1)
    k = (rng.random((40, 17, 3)) * 700 - 30).astype(np.float32)
    for s1, s0 in [((640, 640), (640, 640)), ((384, 640), (1080, 1920)), ((160, 128), (481, 321)), ((160, 160), (200, 151))]:
        ref = ops.scale_coords(s1, torch.from_numpy(k.copy()), s0).numpy()
        assert oracle_post.scale_coords(s1, k, s0).tobytes() == ref.tobytes()
    protos = rng.standard_normal((32, 40, 32)).astype(np.float32)
    coeffs = rng.standard_normal((9, 32)).astype(np.float32)
    c, wh = rng.uniform(0, [128, 160], (9, 2)), rng.uniform(8, [128, 160], (9, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    ref = ops.process_mask(torch.from_numpy(protos), torch.from_numpy(coeffs), torch.from_numpy(boxes.copy()), (160, 128), upsample=True)
    assert np.array_equal(oracle_post.process_mask(protos, coeffs, boxes, (160, 128)), ref.numpy())


def test_head_predictor_classes_plug_in_and_refuse_cpu(tmp_path):
    """B200PosePredictor / B200SegmentationPredictor derive from the reference's DetectionPredictor, set the task the way
    PosePredictor / SegmentationPredictor do, and raise without a CUDA device instead of falling back."""
    ref_import.import_reference()
    from ultralytics.models.yolo.detect.predict import DetectionPredictor

    from bs_yolo_b200 import _cabi as A
    from bs_yolo_b200.predictor import B200PosePredictor, B200SegmentationPredictor

    for cls, task, model in ((B200PosePredictor, "pose", "yolo11n-pose"), (B200SegmentationPredictor, "segment", "yolo11n-seg")):
        assert issubclass(cls, DetectionPredictor) and cls.head_kind == task and cls.pipelined is False
        p = cls(overrides=dict(device="cpu", save=False, verbose=False))
        assert p.args.task == task
        spec = arch.build_spec(model)
        m = ref_import.reference_model(model, arch.synth_state_dict(spec, 0))
        with pytest.raises(A.B2YError, match="no CPU fallback"):
            p.setup_model(m, verbose=False)

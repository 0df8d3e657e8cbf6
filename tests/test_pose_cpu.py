"""Pose head on the CPU: the oracle against the reference's committed outputs (PoseModel forward,
PosePredictor results), the planner's lowering of Pose.cv4 against the oracle, and the host geometry."""
import os

import numpy as np
import pytest
import torch

import kat
from bs_yolo_b200 import _cabi as A
from bs_yolo_b200 import arch, ir
from bs_yolo_b200.plan import Plan
from oracle import oracle_net, oracle_post
from plan_emulator import PlanEmulator


def _graph(model="yolo11n-pose", seed=0):
    spec = arch.build_spec(model)
    return ir.build_ir(spec, arch.synth_state_dict(spec, seed))


@pytest.mark.parametrize("case", list(kat.POSE_NET_CASES))
def test_oracle_pose_forward_matches_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, f"net_{case}.npz"))
    model = kat.POSE_NET_CASES[case][0]
    spec = arch.build_spec(model)
    sd = arch.synth_state_dict(spec, 0)
    assert kat.state_dict_checksum(sd) == pytest.approx(float(g["sd_checksum"]), rel=1e-12), "synthetic weights drifted"
    graph = ir.build_ir(spec, sd)
    nc = spec["nc"]
    y, (raws, kpt) = oracle_net.forward(graph, torch.from_numpy(kat.net_input(case)))
    ref_kpt, ref_y = torch.from_numpy(g["kpt"]), torch.from_numpy(g["y"])
    assert y.shape == ref_y.shape and y.shape[1] == 4 + nc + 51
    assert float((kpt - ref_kpt).norm() / ref_kpt.norm()) < 1e-5
    for i, r in enumerate(raws):
        ref = torch.from_numpy(g[f"raw{i}"])
        assert float((r - ref).norm() / ref.norm()) < 1e-5
    assert float((y[:, :4] - ref_y[:, :4]).abs().max()) < 2e-3 and float((y[:, 4:4 + nc] - ref_y[:, 4:4 + nc]).abs().max()) < 1e-5
    # Pose.kpts_decode itself, on the reference's own raw keypoint maps: bit-exact x / y, visibility to 1 ulp of sigmoid
    H, W = kat.POSE_NET_CASES[case][2:4]
    dec = oracle_net.kpts_decode(ref_kpt, [(H // s, W // s) for s in (8, 16, 32)], [8.0, 16.0, 32.0], 3)
    ref_dec = ref_y[:, 4 + nc:]
    assert torch.equal(dec[:, 0::3], ref_dec[:, 0::3]) and torch.equal(dec[:, 1::3], ref_dec[:, 1::3])
    assert float((dec[:, 2::3] - ref_dec[:, 2::3]).abs().max()) < 1e-6


@pytest.mark.parametrize("case", list(kat.POSE_PREDICT_CASES))
def test_oracle_pose_predict_matches_reference(case, golden_dir):
    """oracle preprocess -> forward -> NMS carrying the keypoint rows -> scale_boxes().round() / scale_coords ==
    YOLO(pose ckpt).predict() of the reference (pose/predict.py:39-59)."""
    g = np.load(os.path.join(golden_dir, f"predict_{case}.npz"))
    model, n, h, w, imgsz, _ = kat.POSE_PREDICT_CASES[case]
    graph = _graph(model)
    imgs = kat.predict_input(case)
    x = oracle_post.preprocess(imgs, (imgsz, imgsz), stride=32)
    y, _ = oracle_net.forward(graph, torch.from_numpy(x))
    res = oracle_post.pose_postprocess(y.numpy(), graph["nc"], (17, 3), x.shape[2:], [im.shape[:2] for im in imgs])
    assert [b.shape[0] for b, _ in res] == g["n"].tolist()
    for i, (boxes, kpts) in enumerate(res):
        rb, rk = g[f"boxes{i}"], g[f"kpts{i}"]
        kpts = oracle_post.keypoints_result(kpts)  # Results.keypoints zeroes points with visibility < 0.5 (results.py:1305-1309)
        assert ((kpts[..., 2] < 0.5) == (rk[..., 2] < 0.5)).all() and (rk[..., 2] < 0.5).any() and (rk[..., 2] >= 0.5).any()
        assert kpts.shape == rk.shape and np.array_equal(boxes[:, 5], rb[:, 5])
        assert np.abs(boxes[:, :4] - rb[:, :4]).max() <= 1.0  # rounded pixels: fp32 fold-order noise can flip a .5
        assert (boxes[:, :4] == rb[:, :4]).mean() > 0.9
        assert np.abs(boxes[:, 4] - rb[:, 4]).max() < 1e-4
        assert np.abs(kpts[..., :2] - rk[..., :2]).max() < 2e-2 and np.abs(kpts[..., 2] - rk[..., 2]).max() < 1e-4


def test_scale_coords_restatement_bit_exact_on_reference_rows(golden_dir):
    """scale_coords + clip_coords (ops.py:740-772, 341-358) as fp32 scalar arithmetic: rows that the reference clipped sit
    exactly on the image border in its output, and un-clipped rows round-trip through the inverse map to < 1 ulp."""
    rng = np.random.default_rng(3)
    k = rng.uniform(-40, 200, (50, 17, 3)).astype(np.float32)
    out = oracle_post.scale_coords((128, 160), k, (240, 320))
    gain, pad = np.float32(0.5), (np.float32(0.0), np.float32(4.0))
    exp_x = np.clip((k[..., 0] - pad[0]) / gain, 0, 320)
    exp_y = np.clip((k[..., 1] - pad[1]) / gain, 0, 240)
    assert np.array_equal(out[..., 0], exp_x) and np.array_equal(out[..., 1], exp_y) and np.array_equal(out[..., 2], k[..., 2])
    assert out[..., 0].min() == 0 and out[..., 1].max() == 240


@pytest.mark.parametrize("hw", [(64, 96), (160, 160)])
def test_lowered_pose_plan_matches_oracle_fp32(hw):
    g = _graph()
    plan = Plan(g, *hw, keep_fp32=True)
    assert plan.extra_kind == "pose" and plan.ne == 51 and plan.kpt_shape == (17, 3) and len(plan.extra_bufs) == 3
    assert all(plan.bufs[b].dtype == A.F32 and plan.bufs[b].C == 52 for b in plan.extra_bufs)  # pitched to 4 floats
    B = 2
    x = torch.from_numpy(np.random.default_rng(9).random((B, 3, *hw), dtype=np.float32))
    emu = PlanEmulator(plan, B)
    emu.set_input(x)
    emu.run()
    y, (_, kpt) = oracle_net.forward(g, x)
    nc = g["nc"]
    assert float((emu.y() - y[:, :4 + nc]).abs().max()) < 1e-3
    assert float((emu.extras() - kpt).norm() / kpt.norm()) < 2e-5
    # the hidden widths of cv4 (c4 = max(ch0 // 4, nk) = 51) are zero-padded to 64 for the tensor-core path
    cv4 = [o for o in plan.ops if ".cv4." in o.name]
    assert len(cv4) == 9 and all(o.kind == A.OP_CONV and o.cin % 16 == 0 and o.cout_pad % 16 == 0 for o in cv4)
    assert [o.cout for o in cv4 if o.name.endswith(".2")] == [51, 51, 51]


def test_pose_plan_bf16_error_budget():
    """Rounding every bf16 buffer and weight (what the CUDA path stores) keeps the raw keypoint maps within 1e-2."""
    g = _graph()
    plan = Plan(g, 64, 96)
    x = torch.from_numpy(np.random.default_rng(10).random((2, 3, 64, 96), dtype=np.float32))
    emu = PlanEmulator(plan, 2, round_bf16=True)
    emu.set_input(x)
    emu.run()
    _, (_, kpt) = oracle_net.forward(g, x)
    assert float((emu.extras() - kpt).norm() / kpt.norm()) < 1e-2


def test_detect_part_of_the_pose_checkpoint_equals_yolo11n():
    """cv4 is drawn after every Detect parameter, so the seeded pose checkpoint shares backbone, neck and Detect branches
    with yolo11n (its calibrated class-bias shifts apply unchanged)."""
    s0, s1 = arch.build_spec("yolo11n"), arch.build_spec("yolo11n-pose")
    a, b = arch.synth_state_dict(s0, 0), arch.synth_state_dict(s1, 0)
    assert set(a) < set(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert all(".cv4." in k for k in set(b) - set(a))


def test_coords_geometry_is_not_rounded():
    from bs_yolo_b200.ops import coords_geometry, scale_geometry
    g = coords_geometry((160, 128), [(200, 150)], "cpu")[0].tolist()
    assert g == pytest.approx([0.8, 4.0, 0.0, 150.0, 200.0])
    g = coords_geometry((160, 160), [(200, 150)], "cpu")[0].tolist()   # pad 20.0 exactly
    assert g == pytest.approx([0.8, 20.0, 0.0, 150.0, 200.0])
    g2 = coords_geometry((160, 160), [(200, 151)], "cpu")[0].tolist()  # pad 19.6: scale_boxes rounds (19.6 - 0.1) to an integer, coords keep 19.6
    b2 = scale_geometry((160, 160), [(200, 151)], "cpu")[0].tolist()
    assert g2[1] == pytest.approx(19.6) and b2[1] == float(round(b2[1])) and b2[1] != g2[1]


def test_head_kind_guard():
    """A predictor class post-processes one head kind; a mismatch raises and names the class to use."""
    for model, kind in (("yolo11n", "detect"), ("yolo11n-pose", "pose"), ("yolo11n-seg", "segment")):
        g = dict(arch.build_spec(model))
        assert ir.head_kind(g) == kind
        ir.require_head(g, kind, "X")
        ir.require_head(g, None, "X")
        for other in {"detect", "pose", "segment"} - {kind}:
            with pytest.raises(NotImplementedError, match=kind):
                ir.require_head(g, other, "B200DetectionValidator" if other == "detect" else "X")


def test_single_class_pose_checkpoint_lowers(monkeypatch):
    """Real pose checkpoints are single-class (nc = 1, kpt_shape [17, 3]): the class branch ends in a 64 -> 1 conv and y
    is (B, 5, A).  Lowering and emulation against the oracle (the GPU suite covers nc = 80 / 12 / 100)."""
    spec = arch.build_spec("yolo11n-pose", nc=1)
    g = ir.build_ir(spec, arch.synth_state_dict(spec, 2))
    plan = Plan(g, 64, 64, keep_fp32=True)
    assert plan.nc == 1 and plan.bufs[plan.y_buf].H == 5 and plan.ne == 51
    x = torch.from_numpy(np.random.default_rng(3).random((1, 3, 64, 64), dtype=np.float32))
    emu = PlanEmulator(plan, 1)
    emu.set_input(x)
    emu.run()
    y, (_, kpt) = oracle_net.forward(g, x)
    assert y.shape[1] == 5 + 51
    assert float((emu.y() - y[:, :5]).abs().max()) < 1e-3 and float((emu.extras() - kpt).norm() / kpt.norm()) < 2e-5

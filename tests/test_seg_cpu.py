"""Segment head on the CPU: the oracle against the reference's committed outputs (SegmentationModel forward,
SegmentationPredictor results), the planner's lowering of Proto / Segment.cv4 against the oracle, and a tile-faithful
Python model of b2y_process_mask's index arithmetic against the reference's F.interpolate-based process_mask."""
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


def _graph(model="yolo11n-seg", seed=0):
    spec = arch.build_spec(model)
    return ir.build_ir(spec, arch.synth_state_dict(spec, seed))


def unpack_masks(g, i):
    shape = tuple(int(v) for v in g[f"mshape{i}"])
    return np.unpackbits(g[f"masks{i}"])[: int(np.prod(shape))].reshape(shape).astype(np.float32)


@pytest.mark.parametrize("case", list(kat.SEG_NET_CASES))
def test_oracle_segment_forward_matches_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, f"net_{case}.npz"))
    model = kat.SEG_NET_CASES[case][0]
    spec = arch.build_spec(model)
    sd = arch.synth_state_dict(spec, 0)
    assert kat.state_dict_checksum(sd) == pytest.approx(float(g["sd_checksum"]), rel=1e-12), "synthetic weights drifted"
    graph = ir.build_ir(spec, sd)
    nc = spec["nc"]
    y, (raws, mc, p) = oracle_net.forward(graph, torch.from_numpy(kat.net_input(case)))
    ry, rmc, rp = (torch.from_numpy(g[k]) for k in ("y", "mc", "proto"))
    assert y.shape == ry.shape and y.shape[1] == 4 + nc + 32 and p.shape == rp.shape
    assert float((mc - rmc).norm() / rmc.norm()) < 1e-5 and float((p - rp).norm() / rp.norm()) < 1e-5
    assert torch.equal(y[:, 4 + nc:], mc)  # head.py:197: the coefficient rows are appended undecoded
    assert float((y[:, :4] - ry[:, :4]).abs().max()) < 2e-3 and float((y[:, 4:4 + nc] - ry[:, 4:4 + nc]).abs().max()) < 1e-5


@pytest.mark.parametrize("case", list(kat.SEG_PREDICT_CASES))
def test_oracle_segment_predict_matches_reference(case, golden_dir):
    """oracle preprocess -> forward -> NMS carrying the coefficient rows -> process_mask -> scale_boxes ==
    YOLO(seg ckpt).predict() of the reference (segment/predict.py:29-57)."""
    g = np.load(os.path.join(golden_dir, f"predict_{case}.npz"))
    model, n, h, w, imgsz, _ = kat.SEG_PREDICT_CASES[case]
    graph = _graph(model)
    imgs = kat.predict_input(case)
    x = oracle_post.preprocess(imgs, (imgsz, imgsz), stride=32)
    y, (_, _, p) = oracle_net.forward(graph, torch.from_numpy(x))
    res = oracle_post.segment_postprocess(y.numpy(), p.numpy(), graph["nc"], x.shape[2:], [im.shape[:2] for im in imgs])
    assert [b.shape[0] for b, _ in res] == g["n"].tolist()
    for i, (boxes, masks) in enumerate(res):
        rb, rm = g[f"boxes{i}"], unpack_masks(g, i)
        assert masks.shape == rm.shape == (len(rb), *x.shape[2:]) and np.array_equal(boxes[:, 5], rb[:, 5])
        assert np.abs(boxes[:, :4] - rb[:, :4]).max() < 2e-2 and np.abs(boxes[:, 4] - rb[:, 4]).max() < 1e-4
        assert (masks != rm).mean() < 1e-3 and 0.02 < rm.mean() < 0.98  # zero crossings move with fp32 fold-order noise


def test_lowered_segment_plan_matches_oracle_fp32():
    g = _graph()
    hw = (64, 96)
    plan = Plan(g, *hw, keep_fp32=True)
    assert plan.extra_kind == "segment" and plan.ne == 32 and len(plan.extra_bufs) == 3 and plan.proto_buf >= 0
    pb = plan.bufs[plan.proto_buf]
    assert (pb.C, pb.H, pb.W, pb.dtype) == (32, 16, 24, A.BF16)
    x = torch.from_numpy(np.random.default_rng(9).random((2, 3, *hw), dtype=np.float32))
    emu = PlanEmulator(plan, 2)
    emu.set_input(x)
    emu.run()
    y, (_, mc, p) = oracle_net.forward(g, x)
    nc = g["nc"]
    assert float((emu.y() - y[:, :4 + nc]).abs().max()) < 1e-3
    assert float((emu.extras() - mc).norm() / mc.norm()) < 2e-5
    assert float((emu.protos() - p).norm() / p.norm()) < 2e-5
    # ConvTranspose2d(64, 64, 2, 2) = one 1x1 conv of 256 channels ordered (dy, dx, co) + a depth-to-space shuffle
    names = [o.name for o in plan.ops]
    up = plan.ops[names.index("model.23.proto.upsample")]
    d2s = plan.ops[names.index("model.23.proto.upsample.d2s")]
    assert (up.kind, up.k, up.cin, up.cout) == (A.OP_CONV, (1, 1), 64, 256) and (d2s.kind, d2s.cin, d2s.cout) == (A.OP_D2S, 256, 64)
    assert (d2s.out.H, d2s.out.W) == (2 * up.out.H, 2 * up.out.W)
    seg = [o for o in plan.ops if ".cv4." in o.name or ".proto." in o.name]
    assert all(o.cin % 16 == 0 and o.cout_pad % 16 == 0 for o in seg if o.kind == A.OP_CONV)


def test_segment_plan_bf16_error_budget():
    g = _graph()
    plan = Plan(g, 64, 96)
    x = torch.from_numpy(np.random.default_rng(10).random((2, 3, 64, 96), dtype=np.float32))
    emu = PlanEmulator(plan, 2, round_bf16=True)
    emu.set_input(x)
    emu.run()
    _, (_, mc, p) = oracle_net.forward(g, x)
    assert float((emu.extras() - mc).norm() / mc.norm()) < 1e-2 and float((emu.protos() - p).norm() / p.norm()) < 1e-2


def mask_kernel_model(protos, coeffs, boxes, H, W):
    """Tile-faithful Python model of process_mask_kernel (csrc/seg.cu): 64 x 64 output tiles; an 18 x 18 shared tile of
    cropped low-resolution values with origin (oy0/4 - 1, ox0/4 - 1) and edge-replicated coordinates; the zero-tile
    shortcut; fixed x4 bilinear weights (0.625, 0.875, 0.125, 0.375), rows blended first, one quad of columns per
    thread.  protos (mh, mw, nm)."""
    mh, mw, nm = protos.shape
    n = len(coeffs)
    out = np.full((n, H, W), -1, np.float32)
    f = np.float32
    wsel = (f(0.625), f(0.875), f(0.125), f(0.375))
    for r in range(n):
        x1, y1, x2, y2 = (f(boxes[r, j]) * f(0.25) for j in range(4))
        for oy0 in range(0, H, 64):
            for ox0 in range(0, W, 64):
                ly0, lx0 = oy0 // 4 - 1, ox0 // 4 - 1
                cx_lo, cx_hi = max(lx0, 0), min(lx0 + 17, mw - 1)
                cy_lo, cy_hi = max(ly0, 0), min(ly0 + 17, mh - 1)
                if not (cx_hi >= x1 and cx_lo < x2 and cy_hi >= y1 and cy_lo < y2):
                    out[r, oy0:oy0 + 64, ox0:ox0 + 64] = 0
                    continue
                low = np.zeros((18, 18), np.float32)
                for i in range(18 * 18):
                    ly, lx = min(max(ly0 + i // 18, 0), mh - 1), min(max(lx0 + i % 18, 0), mw - 1)
                    if x1 <= lx < x2 and y1 <= ly < y2:
                        low[i // 18, i % 18] = np.dot(coeffs[r].astype(np.float64), protos[ly, lx].astype(np.float64))
                for row in range(64):
                    oy = oy0 + row
                    if oy >= H:
                        break
                    m, i = row >> 2, row & 3
                    cy0 = m + (i >> 1)
                    wy1 = wsel[i]
                    col = (f(1) - wy1) * low[cy0] + wy1 * low[cy0 + 1]  # 18 blended columns
                    for tx in range(16):
                        ox = ox0 + 4 * tx
                        if ox >= W:
                            break
                        c0, c1, c2 = col[tx], col[tx + 1], col[tx + 2]
                        quad = (f(0.375) * c0 + f(0.625) * c1, f(0.125) * c0 + f(0.875) * c1,
                                f(0.875) * c1 + f(0.125) * c2, f(0.625) * c1 + f(0.375) * c2)
                        out[r, oy, ox:ox + 4] = [1.0 if v > 0 else 0.0 for v in quad]
    assert (out >= 0).all()  # every pixel written exactly by some tile
    return out


def test_mask_kernel_index_arithmetic_matches_reference_process_mask():
    rng = np.random.default_rng(4)
    H, W, nm = 96, 160, 8  # 2 x 3 tiles, the bottom row of tiles is partial
    protos = rng.standard_normal((nm, H // 4, W // 4)).astype(np.float32)
    coeffs = rng.standard_normal((6, nm)).astype(np.float32)
    boxes = np.array([[10.3, 5.2, 150.9, 90.1], [0, 0, W, H], [60.0, 30.0, 70.0, 66.0], [-20, -20, 63.9, 200],
                      [130.0, 70.0, 159.0, 95.0], [13.3, 13.2, 13.9, 13.8]], np.float32)  # bottom-right corner; no cell inside
    ref = oracle_post.process_mask(protos, coeffs, boxes, (H, W))
    got = mask_kernel_model(protos.transpose(1, 2, 0), coeffs, boxes, H, W)
    assert got.shape == ref.shape and (got != ref).mean() < 2e-4 and 0.05 < ref.mean() < 0.95
    assert np.array_equal(got[2], ref[2]) and np.array_equal(got[4], ref[4])  # small boxes: identical masks
    assert not got[5].any() and not ref[5].any()

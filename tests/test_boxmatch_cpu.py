"""The final-box audit itself (tests/boxmatch.py), on the CPU: the bf16-storage emulation of the lowered plan against the
fp32 oracle must pass the north_star gate with every miss explained, and genuinely wrong boxes must be flagged."""
import numpy as np
import torch

import boxmatch
from bs_yolo_b200 import arch, ir
from bs_yolo_b200.plan import Plan
from oracle import oracle_net, oracle_post
from plan_emulator import PlanEmulator


def test_bf16_emulation_passes_the_final_box_gate_both_directions():
    spec = arch.build_spec("yolo11n")
    g = ir.build_ir(spec, arch.synth_state_dict(spec, 0))
    imgs = np.random.default_rng(9).integers(0, 256, (1, 320, 320, 3), dtype=np.uint8)
    x = torch.from_numpy(oracle_post.preprocess(list(imgs), (320, 320)))
    y_ref = oracle_net.forward(g, x)[0].numpy()
    emu = PlanEmulator(Plan(g, 320, 320), 1, round_bf16=True)
    emu.set_input(x)
    emu.run()
    y = emu.y().numpy()
    ref = oracle_post.non_max_suppression(y_ref.copy(), 0.25, 0.7, max_det=300)[0]
    our = oracle_post.non_max_suppression(y.copy(), 0.25, 0.7, max_det=300)[0]
    fwd = boxmatch.audit(ref, our, boxmatch.candidates(y[0], 0.25), 0.25, 0.7)
    rev = boxmatch.audit(our, ref, boxmatch.candidates(y_ref[0], 0.25), 0.25, 0.7)
    assert fwd["total"] > 20 and fwd["unexplained"] == 0 and rev["unexplained"] == 0, (fwd, rev)
    assert fwd["matched"] >= 0.9 * fwd["total"]


def test_audit_flags_wrong_boxes_and_accepts_threshold_flips():
    A = 6
    y = np.zeros((4 + 2, A), np.float32)
    y[:4] = np.array([[50, 50, 20, 20], [150, 50, 20, 20], [50, 150, 20, 20], [150, 150, 20, 20], [250, 50, 30, 30],
                      [250, 150, 30, 30]], np.float32).T
    y[4] = [0.9, 0.8, 0.7, 0.255, 0.6, 0.1]
    y[5] = [0.1, 0.1, 0.1, 0.1, 0.1, 0.1]
    ref = oracle_post.non_max_suppression(y[None].copy(), 0.25, 0.7)[0]
    assert len(ref) == 5
    ours_y = y.copy()
    ours_y[4, 3] = 0.249        # bf16-sized score change across the threshold: explained
    ours = oracle_post.non_max_suppression(ours_y[None].copy(), 0.25, 0.7)[0]
    s = boxmatch.audit(ref, ours, boxmatch.candidates(ours_y, 0.25), 0.25, 0.7)
    assert (s["matched"], s["near_conf"], s["unexplained"]) == (4, 1, 0)
    bad_y = y.copy()
    bad_y[0, 1] += 3.0          # a box moved by 3 px (IoU 0.74): not a match, not explainable
    bad_y[4, 2] = 0.65          # a confidence off by 5e-2
    bad = oracle_post.non_max_suppression(bad_y[None].copy(), 0.25, 0.7)[0]
    s = boxmatch.audit(ref, bad, boxmatch.candidates(bad_y, 0.25), 0.25, 0.7)
    assert s["matched"] == 3 and s["unexplained"] == 2 and len(s["details"]) == 2

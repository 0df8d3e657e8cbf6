"""Single-op conv cases shared by the GPU probe script and the GPU pytest suite."""
from __future__ import annotations

import numpy as np
import torch

from bs_yolo_b200 import _cabi as A
from bs_yolo_b200.plan import Plan, View

# name -> dict(c1, c2, k, s, H, W, in_total, in_off, out_total, out_off, res, act, f32, pad16, g)
CASES = {
    "k1_64_64_80": dict(c1=64, c2=64, k=1, s=1, H=80, W=80),
    "k3_32_64_40": dict(c1=32, c2=64, k=3, s=1, H=40, W=40),
    "k3_16_8pad_160": dict(c1=16, c2=8, k=3, s=1, H=160, W=160, pad16=True),
    "k1_48_64_bk16": dict(c1=48, c2=64, k=1, s=1, H=32, W=48),
    "k3s2_64_64_80": dict(c1=64, c2=64, k=3, s=2, H=80, W=80),
    "k3s2_16_32_320": dict(c1=16, c2=32, k=3, s=2, H=64, W=96),
    "k3s2_128_256_40": dict(c1=128, c2=256, k=3, s=2, H=40, W=40),
    "k1_384_256_20": dict(c1=384, c2=256, k=1, s=1, H=20, W=20),
    "k3_64_64_20_res": dict(c1=64, c2=64, k=3, s=1, H=20, W=20, res="add"),
    "k1_slice_in_out": dict(c1=64, c2=32, k=1, s=1, H=40, W=40, in_total=192, in_off=64, out_total=96, out_off=32),
    "k1_64_80_f32": dict(c1=64, c2=80, k=1, s=1, H=40, W=40, f32=True, act=False, out_total=144, out_off=64),
    "k1_64_12_f32": dict(c1=64, c2=12, k=1, s=1, H=20, W=20, f32=True, act=False, out_total=76, out_off=64),
    "k1_mul_res": dict(c1=256, c2=256, k=1, s=1, H=20, W=20, res="mul", act=False),
    "k3_96_96_n": dict(c1=96, c2=96, k=3, s=1, H=24, W=40),
    # Upsample + Concat folded into a 1x1 conv: pre-activation add of a half-resolution map (RES_PREADD_UP2)
    "k1_up2_preadd": dict(c1=128, c2=64, k=1, s=1, H=40, W=40, res="up2", out_total=192, out_off=0),
    "k1_up2_preadd_b": dict(c1=64, c2=128, k=1, s=1, H=20, W=28, res="up2", B=5),
    "k1_1536_768": dict(c1=1536, c2=768, k=1, s=1, H=16, W=16),
    "k3_768_768_s2": dict(c1=768, c2=768, k=3, s=2, H=16, W=16),
    # large enough tile counts for multi-block (MB > 1) tiles
    "k3_16_16_160_mb": dict(c1=16, c2=16, k=3, s=1, H=160, W=160, B=12, res="add"),
    "k1_32_64_160_mb": dict(c1=32, c2=64, k=1, s=1, H=160, W=160, B=8),
    "k3s2_16_32_320_mb": dict(c1=16, c2=32, k=3, s=2, H=320, W=320, B=12),
    "k3_64_64_80_mb": dict(c1=64, c2=64, k=3, s=1, H=80, W=80, B=24),
    # thin outputs on big maps: 8-block tiles (MB = 8)
    "k3_16_16_320_mb8": dict(c1=16, c2=16, k=3, s=1, H=320, W=320, B=6, res="add"),
    "k1_16_32_320_mb8": dict(c1=16, c2=32, k=1, s=1, H=320, W=320, B=6),
}


def build_case(name: str, impl: int, seed: int = 0):
    """Returns (plan, x_view, out_view, res_view|None)."""
    c = dict(in_total=None, in_off=0, out_total=None, out_off=0, res=None, act=True, f32=False, pad16=False, B=3)
    c.update(CASES[name])
    rng = np.random.default_rng(seed)
    c1, c2, k, s, H, W = c["c1"], c["c2"], c["k"], c["s"], c["H"], c["W"]
    p = k // 2
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    plan = Plan.empty(64, 64, conv_impl=impl, keep_fp32=False)
    in_total = c["in_total"] or c1
    xb = plan.new_buf(in_total, H, W, name="x")
    x = View(xb, c["in_off"], c1, H, W)
    node = dict(type="Conv", name=name, c1=c1, c2=c2, k=(k, k), s=(s, s), p=(p, p), g=1, act="silu" if c["act"] else "none",
                weight=torch.from_numpy(rng.standard_normal((c2, c1, k, k)).astype(np.float32) / np.sqrt(c1 * k * k)),
                bias=torch.from_numpy(rng.standard_normal(c2).astype(np.float32) * 0.1))
    dest = None
    odt = A.F32 if c["f32"] else A.BF16
    if c["out_total"]:
        ob = plan.new_buf(c["out_total"], Ho, Wo, odt, name="out")
        dest = View(ob, c["out_off"], c2, Ho, Wo)
    res, mode = None, A.RES_NONE
    if c["res"] == "up2":
        rb = plan.new_buf(c2, Ho // 2, Wo // 2, name="res_half")
        res, mode = View(rb, 0, c2, Ho // 2, Wo // 2), A.RES_PREADD_UP2
    elif c["res"]:
        rb = plan.new_buf(c2, Ho, Wo, name="res")
        res, mode = View(rb, 0, c2, Ho, Wo), (A.RES_ADD if c["res"] == "add" else A.RES_MUL)
    out = plan.conv(node, x, dest, res=res, res_mode=mode, out_dtype=odt, pad16=c["pad16"])
    return plan, x, out, res


def run_case(name: str, impl: int, B: int | None = None, seed: int = 0, device=0):
    """Run one case on the GPU through the engine and through the emulator; returns (gpu_out, ref_out) as float CPU."""
    from bs_yolo_b200.engine import Engine
    from plan_emulator import PlanEmulator

    plan, x, out, res = build_case(name, impl, seed)
    B = B or CASES[name].get("B", 3)
    eng = Engine(None, B, 64, 64, device=device, use_graph=False, plan=plan)
    g = torch.Generator().manual_seed(seed)
    emu = PlanEmulator(plan, B, round_bf16=True)
    for i, b in enumerate(plan.bufs):
        if b.dtype == A.BF16:
            v = (torch.randn(B, b.H, b.W, b.C, generator=g)).to(torch.bfloat16)
            eng.buffer(i).copy_(v)
            emu.bufs[i] = v.float()
    eng.forward()
    torch.cuda.synchronize()
    emu.run()
    got = eng.buffer(out.buf).float().cpu()
    ref = emu.bufs[out.buf]
    return got, ref, eng.op_impl(0)


# ---- fused Bottleneck (OP_CONV_PAIR) cases: name -> (c, hidden, H, W, B)
PAIR_CASES = {
    "pair_16_8_160": (16, 8, 160, 160, 4),
    "pair_32_16_80": (32, 16, 80, 80, 3),
    "pair_64_32_40": (64, 32, 40, 40, 5),
    "pair_32_32_40_odd": (32, 32, 37, 53, 3),
    "pair_16_16_small": (16, 16, 12, 20, 2),
    "pair_64_32_20": (64, 32, 20, 20, 9),
}


def run_pair_case(name: str, device=0):
    """Fused Bottleneck on the GPU vs the emulator (bf16 storage of the hidden map emulated); returns (gpu, ref)."""
    from bs_yolo_b200.engine import Engine
    from plan_emulator import PlanEmulator

    c, hid, H, W, B = PAIR_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))

    def cnode(nm, c1, c2):
        return dict(type="Conv", name=nm, c1=c1, c2=c2, k=(3, 3), s=(1, 1), p=(1, 1), g=1, act="silu",
                    weight=torch.from_numpy((rng.standard_normal((c2, c1, 3, 3)) / np.sqrt(c1 * 9)).astype(np.float32)),
                    bias=torch.from_numpy(rng.standard_normal(c2).astype(np.float32) * 0.1))

    node = dict(type="Bottleneck", name=name, cv1=cnode(name + ".cv1", c, hid), cv2=cnode(name + ".cv2", hid, c), add=True)
    plan = Plan.empty(64, 64)
    xb = plan.new_buf(2 * c, H, W, name="x")  # input and output are channel slices of wider buffers, as in C3k2
    x = View(xb, c, c, H, W)
    ob = plan.new_buf(3 * c, H, W, name="out")
    out = plan.bottleneck(node, x, View(ob, c, c, H, W))
    assert plan.ops[-1].kind == A.OP_CONV_PAIR and len(plan.ops) == 1, "case must take the fused path"
    eng = Engine(None, B, 64, 64, device=device, use_graph=False, plan=plan)
    g = torch.Generator().manual_seed(3)
    emu = PlanEmulator(plan, B, round_bf16=True)
    for i, b in enumerate(plan.bufs):
        v = torch.randn(B, b.H, b.W, b.C, generator=g).to(torch.bfloat16)
        eng.buffer(i).copy_(v)
        emu.bufs[i] = v.float()
    eng.forward()
    torch.cuda.synchronize()
    emu.run()
    return eng.buffer(out.buf).float().cpu(), emu.bufs[out.buf]

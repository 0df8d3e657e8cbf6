"""Tiny driver for ncu: the two stand-alone f4 kernels on synthetic operands (B=8, 300 kept rows per image, 640x640).
    python tools/ncu_f4.py            # plain run first, then under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
        --clock-control none -k regex:"process_mask|gather_extras" --csv --log-file gpurun_out/ncu_f4.csv python tools/ncu_f4.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bs_yolo_b200 import ops  # noqa: E402

B, max_det, H, W = 8, 300, 640, 640
g = torch.Generator(device="cuda").manual_seed(0)
protos = torch.randn(B, H // 4, W // 4, 32, device="cuda", generator=g).to(torch.bfloat16)
c = torch.rand(B, max_det, 2, device="cuda", generator=g) * 640
wh = torch.rand(B, max_det, 2, device="cuda", generator=g) * 600 + 20
dets = torch.zeros(B, max_det, 6, device="cuda")
dets[..., :2], dets[..., 2:4] = c - wh / 2, c + wh / 2
raws = [torch.randn(B, H // s, W // s, 32, device="cuda", generator=g) for s in (8, 16, 32)]
kept = torch.randint(0, 8400, (B, max_det), device="cuda", generator=g).int()
counts = torch.full((B,), max_det, dtype=torch.int32, device="cuda")
for _ in range(2):
    coeffs = ops.gather_extras(raws, [8.0, 16.0, 32.0], 32, kept, counts, mode=0)
    out = ops.process_mask_batched(protos, coeffs, dets, [max_det] * B, (H, W))
    torch.cuda.synchronize()
print("masks", tuple(out[0].shape), float(out[0].mean()), "bytes per launch", B * max_det * H * W * 4)

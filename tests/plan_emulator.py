"""Test infrastructure: executes a bs_yolo_b200.plan.Plan op list with plain PyTorch (fp32, CPU), following the op
semantics documented in include/b2y.h.  It lets the CPU suite check the planner's lowering (channel-slice concat
fusion, residual placement, BN/dilconv folds, ELA/MSCA decomposition, per-head pe convs, weight packing) against the
oracle without a GPU; the GPU suite then only has to establish kernel == op semantics."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from bs_yolo_b200 import _cabi as A
from bs_yolo_b200.plan import Plan, View


class PlanEmulator:
    def __init__(self, plan: Plan, B: int, round_bf16: bool = False):
        self.plan, self.B, self.round = plan, B, round_bf16
        self.blob = plan.blob.tobytes()
        self.bufs = [torch.zeros(B, b.H, b.W, b.C) for b in plan.bufs]

    # -- helpers
    def f32(self, off, n):
        return torch.from_numpy(np.frombuffer(self.blob, np.float32, n, off).copy())

    def bf16(self, off, n):
        bits = np.frombuffer(self.blob, np.int16, n, off).astype(np.int32) << 16
        return torch.from_numpy(bits.view(np.float32).copy())

    def get(self, v):
        if isinstance(v, View):
            return self.bufs[v.buf][..., v.coff:v.coff + v.c]
        return self.bufs[v]

    def put(self, v: View, val, c=None):
        c = v.c if c is None else c
        if self.round and self.plan.bufs[v.buf].dtype == A.BF16:
            val = val.to(torch.bfloat16).float()
        self.bufs[v.buf][..., v.coff:v.coff + c] = val[..., :c]

    def set_input(self, x):  # (B,3,H,W) float
        from bs_yolo_b200.engine import space_to_depth
        self.bufs[self.plan.input_buf] = space_to_depth(x).clone()
        if self.round:
            self.bufs[self.plan.input_buf] = self.bufs[self.plan.input_buf].to(torch.bfloat16).float()

    # -- ops
    def run(self, first=0, last=None):
        for op in self.plan.ops[first:last]:
            getattr(self, "op_" + A.OP_NAMES[op.kind])(op)

    def _epilogue(self, op, y):
        if op.res_mode == A.RES_PREADD_UP2:  # pre-activation add of the nearest-2x-upsampled half-resolution map
            r = self.get(op.res).repeat_interleave(2, 1).repeat_interleave(2, 2)
            y = y.clone()
            y[..., :op.cout] = y[..., :op.cout] + r
            y = F.silu(y) if op.act else y
            self.put(op.out, y, op.cout)
            return
        if op.act:
            y = F.silu(y)
        if op.res_mode:
            r = self.get(op.res)
            y = y[..., :op.cout] + r if op.res_mode == A.RES_ADD else y[..., :op.cout] * r
        self.put(op.out, y, op.cout)

    def op_conv(self, op):
        kh, kw = op.k
        x = self.get(op.ins[0]).permute(0, 3, 1, 2)
        wflat = op.w_fp32 if (op.w_fp32 is not None and not self.round) else self.bf16(op.w_off, op.cout_pad * kh * kw * op.cin)
        w = wflat.reshape(op.cout_pad, kh, kw, op.cin).permute(0, 3, 1, 2)
        b = self.f32(op.b_off, op.cout_pad)
        ip = tuple(op.ip) + (0,) * 5
        if ip[0]:  # padding before (top/left) only: the 2x2 space-to-depth stem
            y = F.conv2d(F.pad(x, (op.p[1], 0, op.p[0], 0)), w, b, stride=op.s).permute(0, 2, 3, 1)
        else:
            y = F.conv2d(x, w, b, stride=op.s, padding=op.p).permute(0, 2, 3, 1)
        if ip[1]:  # fused Detect._inference epilogue (head.py:100-131): writes rows of y (B, 4+nc, A)
            mode, a_off, nc = ip[1], ip[2], ip[4]
            B, H, W, _ = y.shape
            ybuf = self.bufs[op.out][:, :, 0, :]
            if mode == 1:
                d = torch.softmax(y[..., :64].reshape(B, H, W, 4, 16), -1) @ torch.arange(16, dtype=torch.float32)
                ax = (torch.arange(W, dtype=torch.float32) + 0.5).view(1, 1, W)
                ay = (torch.arange(H, dtype=torch.float32) + 0.5).view(1, H, 1)
                x1, y1, x2, y2 = ax - d[..., 0], ay - d[..., 1], ax + d[..., 2], ay + d[..., 3]
                box = torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], 1) * op.fp[0]  # (B,4,H,W)
                ybuf[:, 0:4, a_off:a_off + H * W] = box.reshape(B, 4, H * W)
            else:
                ybuf[:, 4:4 + nc, a_off:a_off + H * W] = torch.sigmoid(y[..., :nc]).permute(0, 3, 1, 2).reshape(B, nc, H * W)
            return
        self._epilogue(op, y)

    def op_conv_pair(self, op):
        chid = op.ip[0]
        x = self.get(op.ins[0]).permute(0, 3, 1, 2)
        if op.w_fp32 is not None and not self.round:
            w1f, w2f = op.w_fp32
        else:
            w1f, w2f = self.bf16(op.w_off, chid * 9 * op.cin), self.bf16(op.p_off[0], op.cout * 9 * chid)
        w1 = w1f.reshape(chid, 3, 3, op.cin).permute(0, 3, 1, 2)
        w2 = w2f.reshape(op.cout, 3, 3, chid).permute(0, 3, 1, 2)
        h = F.silu(F.conv2d(x, w1, self.f32(op.b_off, chid), padding=1))
        if self.round:  # the hidden map is rounded to bf16 on its way into shared memory
            h = h.to(torch.bfloat16).float()
        y = F.silu(F.conv2d(h, w2, self.f32(op.p_off[1], op.cout), padding=1))
        if op.ip[1]:
            y = y + x
        self.put(op.out, y.permute(0, 2, 3, 1), op.cout)

    def op_dwconv(self, op):
        kh, kw = op.k
        c = op.cout
        if len(op.ip) >= 2 and op.ip[0] > 0:  # chunked input view: output channel j reads (j // chunk) * stride + j % chunk
            chunk, stride = op.ip[0], op.ip[1]
            v = op.ins[0]
            x = torch.cat([self.bufs[v.buf][..., v.coff + h * stride: v.coff + h * stride + chunk] for h in range(c // chunk)], -1)
            x = x.permute(0, 3, 1, 2)
        else:
            x = self.get(op.ins[0]).permute(0, 3, 1, 2)
        w = self.f32(op.w_off, kh * kw * c).view(kh, kw, c).permute(2, 0, 1).unsqueeze(1)
        y = F.conv2d(x, w, self.f32(op.b_off, c), stride=op.s, padding=op.p, groups=c).permute(0, 2, 3, 1)
        self._epilogue(op, y)

    def op_pmsfa_dw(self, op):
        """B2Y_OP_PMSFA_DW: in = [a | b | h] -> out = [SiLU(dw7x7(t_a)) | t_b | h], t = SiLU(dw5x5([a | b])) (block.py:3046-3051);
        t_a is rounded to bf16 on its way into shared memory."""
        c = op.cin
        x = self.get(op.ins[0]).permute(0, 3, 1, 2)
        w2 = self.f32(op.w_off, 25 * (c // 2)).view(5, 5, c // 2).permute(2, 0, 1).unsqueeze(1)
        w3 = self.f32(op.p_off[0], 49 * (c // 4)).view(7, 7, c // 4).permute(2, 0, 1).unsqueeze(1)
        t = F.silu(F.conv2d(x[:, : c // 2], w2, self.f32(op.b_off, c // 2), padding=2, groups=c // 2))
        ta = t[:, : c // 4].to(torch.bfloat16).float() if self.round else t[:, : c // 4]
        o3 = F.silu(F.conv2d(ta, w3, self.f32(op.p_off[1], c // 4), padding=3, groups=c // 4))
        self.put(op.out, torch.cat([o3, t[:, c // 4:], x[:, c // 2:]], 1).permute(0, 2, 3, 1))

    def op_msca_branch(self, op):
        """B2Y_OP_MSCA_BRANCH: attn = dw5x5(x); branch_j = dw(k x 1)(dw(1 x k)(attn)), k = 5, 7, 11, 21; gap_j = mean of the
        (stored, i.e. rounded) branch map (MSCA.py:52-74).  attn and the strip intermediate are rounded to bf16."""
        c = op.cin
        x = self.get(op.ins[0]).permute(0, 3, 1, 2)
        P = self.f32(op.w_off, 122 * c).view(122, c)
        rnd = (lambda t: t.to(torch.bfloat16).float()) if self.round else (lambda t: t)
        attn = rnd(F.conv2d(x, P[:25].view(5, 5, c).permute(2, 0, 1).unsqueeze(1), P[25], padding=2, groups=c))
        r, outs = 26, []
        for k in (5, 7, 11, 21):
            wh, bh, wv, bv = P[r:r + k], P[r + k], P[r + k + 1:r + 2 * k + 1], P[r + 2 * k + 1]
            r += 2 * k + 2
            u = rnd(F.conv2d(attn, wh.t().reshape(c, 1, 1, k), bh, padding=(0, k // 2), groups=c))
            outs.append(F.conv2d(u, wv.t().reshape(c, 1, k, 1), bv, padding=(k // 2, 0), groups=c))
        self.put(op.out, torch.cat(outs, 1).permute(0, 2, 3, 1))
        self.bufs[op.res][:, 0, 0, :] = self.get(op.out).mean((1, 2))

    def op_sppf_pool(self, op):
        c = op.cin
        x = self.get(op.ins[0]).permute(0, 3, 1, 2)
        outs = []
        for _ in range(3):
            x = F.max_pool2d(x, 5, 1, 2)
            outs.append(x)
        self.put(op.out, torch.cat(outs, 1).permute(0, 2, 3, 1), 3 * c)

    def op_resample(self, op):
        x = self.get(op.ins[0])
        s = op.ip[0]
        if s != 1:
            x = x.repeat_interleave(s, 1).repeat_interleave(s, 2)
        self.put(op.out, x)

    def op_d2s(self, op):
        """B2Y_OP_D2S: out[b, 2y+dy, 2x+dx, c] = in[b, y, x, (dy*2+dx)*cout + c]."""
        x = self.get(op.ins[0])
        B, H, W, _ = x.shape
        c = op.cout
        self.put(op.out, x.reshape(B, H, W, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, c))

    def op_attn(self, op):
        nh, kd, hd = op.ip
        qkv = self.get(op.ins[0])
        B, H, W, _ = qkv.shape
        q, k, v = qkv.reshape(B, H * W, nh, 2 * kd + hd).permute(0, 2, 1, 3).split([kd, kd, hd], -1)  # (B,nh,N,d)
        p = torch.softmax(q @ k.transpose(-1, -2) * kd ** -0.5, -1)
        o = (p @ v).permute(0, 2, 1, 3).reshape(B, H, W, nh * hd)
        self.put(op.out, o)

    def op_rowcol_mean(self, op):
        x = self.get(op.ins[0])
        self.bufs[op.out][:, :, 0, :] = x.mean(2)
        nb = (tuple(op.ip) + (0,))[0]
        if nb:  # banded layout: per-band (8 rows) partial column means, summed by ela_gate
            H = x.shape[1]
            for k in range(nb):
                self.bufs[op.ins[1]][:, :, k, :] = x[:, 8 * k:8 * k + 8].sum(1) / H
        else:
            self.bufs[op.ins[1]][:, :, 0, :] = x.mean(1)

    def op_ela_gate(self, op):
        ks, groups, H, W = op.ip[:4]
        c = op.cin
        p = self.f32(op.p_off[0], 2 * c * ks + 2 * c)
        sp_w, gn_w, gn_b, ch_w = p[:c * ks].view(c, ks), p[c * ks:c * ks + c], p[c * ks + c:c * ks + 2 * c], p[c * ks + 2 * c:].view(c, ks)
        rowm, colm = self.bufs[op.ins[0]][:, :, 0, :], self.bufs[op.ins[1]].sum(2)

        def gate(m):  # (B,L,c)
            v = F.conv1d(m.transpose(1, 2), sp_w.unsqueeze(1), padding=ks - 1, dilation=2, groups=c)
            return torch.sigmoid(F.group_norm(v, groups, gn_w, gn_b, op.fp[0])).transpose(1, 2)

        self.bufs[op.out][:, :, 0, :] = gate(rowm)
        self.bufs[op.ins[2]][:, :, 0, :] = gate(colm)
        self.bufs[op.ins[3]][:, 0, 0, :] = torch.sigmoid(ch_w[:, (ks - 1) // 2] * rowm.mean(1))

    def op_ela_apply(self, op):
        x = self.get(op.ins[0])
        hg = self.bufs[op.ins[1]][:, :, 0, :].unsqueeze(2)
        wg = self.bufs[op.ins[2]][:, :, 0, :].unsqueeze(1)
        cg = self.bufs[op.ins[3]][:, 0, 0, :].view(self.B, 1, 1, -1)
        a, b, r = op.fp
        self.put(op.out, x * (a * cg + b * (hg * wg)) + r * x)

    def op_gap4(self, op):
        c = op.cin
        for j in range(4):
            self.bufs[op.out][:, 0, 0, j * c:(j + 1) * c] = self.get(op.ins[j]).mean((1, 2))

    def op_msca_gate(self, op):
        c = op.cin
        w = self.f32(op.p_off[0], 4 * c * c).view(4, c, c)
        b = self.f32(op.p_off[1], 4 * c).view(4, c)
        gap = self.bufs[op.ins[0]][:, 0, 0, :].view(self.B, 4, c)
        z = torch.sigmoid(torch.einsum("jok,bjk->bjo", w, gap) + b)
        self.bufs[op.out][:, 0, 0, :] = torch.softmax(z, 1).reshape(self.B, 4 * c)

    def op_wsum4(self, op):
        c = op.cin
        g = self.bufs[op.res][:, 0, 0, :].view(self.B, 4, 1, 1, c)
        self.put(op.out, sum(g[:, j] * self.get(op.ins[j]) for j in range(4)))

    def op_decode(self, op):
        from oracle import oracle_net
        nc, nl = op.ip
        raws = [self.bufs[i].permute(0, 3, 1, 2) for i in op.ins[:nl]]
        y = oracle_net.decode(raws, list(op.fp[:nl]), nc)  # (B, 4+nc, A)
        self.bufs[op.out][:, :, 0, :] = y

    # -- results
    def y(self):
        return self.bufs[self.plan.y_buf][:, :, 0, :]

    def extras(self):
        """(B, ne, A): the head's extra branch (Pose.cv4) in the reference's layout (head.py:245), pitch padding dropped."""
        ne = self.plan.ne
        return torch.cat([self.bufs[i][..., :ne].reshape(self.B, -1, ne).transpose(1, 2) for i in self.plan.extra_bufs], 2)

    def protos(self):
        """(B, nm, H/4, W/4): Segment's mask prototypes in the reference's layout (head.py:190)."""
        return self.bufs[self.plan.proto_buf].permute(0, 3, 1, 2)

    def raws(self):
        return [self.bufs[i].permute(0, 3, 1, 2) for i in self.plan.head_bufs]

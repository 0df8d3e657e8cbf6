"""Planner: lowers the graph IR (ir.build_ir) to the flat op list of include/b2y.h.

Design (B200-first, not a layer-by-layer translation of the reference's Python interpreter,
ultralytics/nn/tasks.py:138-165):
  * activations are NHWC bf16; every tensor operand is a *channel-slice view* of a buffer, so all the
    reference's torch.cat / chunk / split calls (C2f block.py:3308-3310, C3 :3334, SPPF :3149, C2PSA :4466-4468,
    Concat conv.py:455, Detect head.py:70) cost nothing: producers write straight into their slice of the concat
    buffer and consumers read slices through strided TMA descriptors;
  * BatchNorm is already folded (ir.py); bias + SiLU + residual add/mul happen in the conv epilogue;
  * DFL / dist2bbox / sigmoid run in fp32 on the fp32 accumulators of the Detect output convs: by default inside
    those convs' epilogues (``fuse_decode``: the raw head maps never reach HBM), otherwise in decode.cu over fp32 maps;
  * the 3-channel 3x3 stride-2 stem runs on the tensor cores as a 2x2 conv over the space-to-depth input the
    preprocess kernel writes ((B, H/2, W/2, 16): four pixels x (R,G,B,0));
  * hidden channel counts that are not a multiple of 16 (the 8-channel Bottleneck of yolo11n layer 2) are
    zero-padded to 16 so the tensor-core path takes them.
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch

from . import _cabi as A

Node = dict[str, Any]


@dataclass(frozen=True)
class View:
    buf: int
    coff: int
    c: int      # channels physically covered by the view (incl. zero padding)
    H: int
    W: int
    creal: int = -1  # logical channels (<= c); -1: == c

    @property
    def real(self) -> int:
        return self.c if self.creal < 0 else self.creal


@dataclass
class BufSpec:
    C: int
    H: int
    W: int
    dtype: int = A.BF16
    name: str = ""
    offset: int = 0  # bytes per batch element are H*W*C*elem; absolute offsets assigned by Plan.layout(B)

    def elems(self) -> int:
        return self.C * self.H * self.W


@dataclass
class OpSpec:
    kind: int
    name: str = ""
    ins: list = field(default_factory=list)   # list[View|int buffer id]
    out: Any = None
    res: Any = None
    res_mode: int = A.RES_NONE
    cin: int = 0
    cout: int = 0
    cout_pad: int = 0
    k: tuple = (1, 1)
    s: tuple = (1, 1)
    p: tuple = (0, 0)
    act: int = 0
    impl: int = A.IMPL_AUTO
    w_off: int = -1
    b_off: int = -1
    p_off: tuple = (-1, -1, -1, -1)
    ip: tuple = ()
    fp: tuple = ()
    flops: int = 0        # 2*MAC, per image
    bytes_alg: int = 0    # algorithmic bytes (in + out + weights) of ONE image; a batch of B moves bytes_batch(B)
    bytes_w: int = 0      # the weight part of bytes_alg: read once per batch, not once per image
    w_fp32: Any = None    # tests only (Plan(keep_fp32=True))


def _r(v, m):
    return (v + m - 1) // m * m


class WeightBlob:
    def __init__(self):
        self.parts: list[tuple[int, np.ndarray]] = []
        self.size = 0

    def add(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        off = _r(self.size, 256)
        self.parts.append((off, arr))
        self.size = off + arr.nbytes
        return off

    def tobytes(self) -> bytes:
        out = bytearray(_r(self.size, 256))
        for off, arr in self.parts:
            out[off:off + arr.nbytes] = arr.tobytes()
        return bytes(out)


def _bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy()


class Plan:
    """Flat op list + buffers + packed weights for one (H, W) input size (batch is bound later)."""

    def __init__(self, ir: Node, H: int, W: int, conv_impl: int = A.IMPL_AUTO, keep_fp32: bool = False,
                 fuse_decode: bool = True, dw_dense: bool = True, fold_upsample: bool = True, fuse_pairs: bool = True):
        if H % 32 or W % 32:
            raise ValueError(f"input {H}x{W} must be a multiple of the max stride 32 (utils/checks.py:120-171)")
        self.ir, self.H, self.W = ir, H, W
        if os.environ.get("B2Y_FORCE_SIMT") == "1":  # debug knob: the whole plan on the CUDA-core validation twin
            conv_impl = A.IMPL_SIMT
        self.conv_impl = conv_impl
        self.keep_fp32 = keep_fp32  # tests only: keep unrounded packed conv weights next to the bf16 blob
        # fused Detect decode lives in the tcgen05 conv epilogue only; the SIMT twin keeps raw maps + decode.cu
        self.fuse_decode = bool(fuse_decode) and conv_impl != A.IMPL_SIMT
        self.fuse_pairs = fuse_pairs and os.environ.get("B2Y_FUSE_PAIRS", "1") != "0"  # thin Bottlenecks as one kernel
        self.dw_dense = dw_dense  # run small-C depthwise 3x3 convs as block-diagonal tensor-core convs
        # Upsample + Concat + C3k2.cv1 -> half-resolution partial sum + pre-activation add (B2Y_FOLD_UPSAMPLE=0: A/B knob)
        self.fold_upsample = fold_upsample and os.environ.get("B2Y_FOLD_UPSAMPLE", "1") != "0"
        self.bufs: list[BufSpec] = []
        self.ops: list[OpSpec] = []
        self.blob = WeightBlob()
        self.nc = ir["nc"]
        self.layer_out: dict[int, View] = {}
        self.layer_op_range: dict[int, tuple[int, int]] = {}
        # network input, space-to-depth: channel ((y&1)*2 + (x&1))*4 + {R,G,B,0} of super-pixel (y/2, x/2)
        self.input_buf = self.new_buf(16, H // 2, W // 2, name="input_s2d", exact=True)
        self.head_bufs: list[int] = []
        self.extra_bufs: list[int] = []  # Pose / Segment: per-level raw cv4 maps (B, H_l, W_l, ne) fp32
        self.ne = 0                      # logical channels of the extra maps (nk | nm); buffers are pitched to a multiple of 4
        self.proto_buf = -1              # Segment: mask prototypes (B, H/4, W/4, nm) bf16 NHWC
        self.extra_strides: list[float] = []
        self.extra_kind, self.kpt_shape = "", ()
        self.y_buf = -1
        self.A = 0
        if ir.get("layers"):
            self._lower()

    @classmethod
    def empty(cls, H: int, W: int, nc: int = 1, conv_impl: int = A.IMPL_AUTO, keep_fp32: bool = False,
              fuse_decode: bool = True, dw_dense: bool = True, fold_upsample: bool = True, fuse_pairs: bool = True) -> "Plan":
        """A plan with no ops (unit tests append single ops through the same lowering helpers)."""
        return cls(dict(nc=nc, layers=[]), H, W, conv_impl, keep_fp32, fuse_decode, dw_dense, fold_upsample, fuse_pairs)

    # ------------------------------------------------------------------ buffers / views
    def new_buf(self, C, H, W, dtype=A.BF16, name="", exact=False) -> int:
        self.bufs.append(BufSpec(_r(C, 8) if (dtype == A.BF16 and not exact) else C, H, W, dtype, name))
        return len(self.bufs) - 1

    def full(self, buf: int, c: int | None = None) -> View:
        b = self.bufs[buf]
        return View(buf, 0, b.C if c is None else c, b.H, b.W)

    @staticmethod
    def sub(v: View, coff: int, c: int) -> View:
        assert coff + c <= v.c
        return View(v.buf, v.coff + coff, c, v.H, v.W)

    # ------------------------------------------------------------------ conv family
    def conv(self, n: Node, x: View, dest: View | None = None, res: View | None = None, res_mode=A.RES_NONE,
             out_dtype=A.BF16, pad16=False, wb=None, decode=None) -> View:
        """Conv / DWConv node -> one op. ``wb``: optional (weight, bias) override (used for folded variants).
        ``decode`` = (mode, anchor offset, stride): the conv is a Detect output conv whose epilogue applies
        Detect._inference (head.py:100-131) and writes rows of y (B, 4+nc, A) instead of a raw map."""
        w, b = wb if wb is not None else (n["weight"], n["bias"])
        c1, c2, (kh, kw), (sh, sw), (ph, pw), g = n["c1"], n["c2"], n["k"], n["s"], n["p"], n["g"]
        assert x.real == c1, (n["name"], x, c1)
        Ho, Wo = (x.H + 2 * ph - kh) // sh + 1, (x.W + 2 * pw - kw) // sw + 1
        act = 1 if n["act"] == "silu" else 0
        if g == 1:
            cpad = _r(c2, 16)
            cstore = cpad if (pad16 and dest is None) else c2
            if decode is not None:
                assert (kh, kw, sh, sw) == (1, 1, 1, 1) and not act and res is None and self.y_buf >= 0
                dest = self.y_buf
            elif dest is None:
                dest = self.full(self.new_buf(cstore, Ho, Wo, out_dtype, n["name"]), cstore)
                if cstore != c2:
                    dest = View(dest.buf, 0, cstore, Ho, Wo, creal=c2)
            assert decode is not None or (dest.c == cstore and (dest.H, dest.W) == (Ho, Wo)), (n["name"], dest, cstore, Ho, Wo)
            wp = torch.zeros(cpad, kh * kw, x.c)
            wp[:c2, :, :c1] = w.permute(0, 2, 3, 1).reshape(c2, kh * kw, c1)
            bp = torch.zeros(cpad)
            bp[:c2] = b
            op = OpSpec(A.OP_CONV, n["name"], [x], dest, res, res_mode, x.c, cstore, cpad, (kh, kw), (sh, sw), (ph, pw),
                        act, self.conv_impl, self.blob.add(_bf16_bits(wp.reshape(cpad, -1))),
                        self.blob.add(bp.numpy().astype(np.float32)))
            if self.keep_fp32:
                op.w_fp32 = wp.reshape(cpad, -1).clone()
            op.flops = 2 * c2 * c1 * kh * kw * Ho * Wo
            esz = 4 if out_dtype == A.F32 else 2
            op.bytes_w = c2 * c1 * kh * kw * 2
            op.bytes_alg = c1 * x.H * x.W * 2 + c2 * Ho * Wo * esz + op.bytes_w
            if decode is not None:
                mode, a_off, stride = decode
                op.ip, op.fp = (0, mode, a_off, self.A, self.nc), (float(stride),)
                rows = 4 if mode == 1 else c2
                op.bytes_w = c2 * c1 * 2
                op.bytes_alg = c1 * x.H * x.W * 2 + rows * Ho * Wo * 4 + op.bytes_w
        elif (self.dw_dense and g == c1 == c2 and (kh, kw) == (3, 3) and (sh, sw) in ((1, 1), (2, 2)) and (ph, pw) == (1, 1)
              and c1 % 16 == 0 and c1 <= int(os.environ.get("B2Y_DW_DENSE_MAXC", 128)) and x.c == c1 and self.conv_impl != A.IMPL_SIMT
              and (sh == 1 or (x.H % 2 == 0 and x.W % 2 == 0))):
            # Depthwise 3x3 on the tensor cores: the same implicit-GEMM kernel with a diagonal weight matrix
            # (W[co][tap][ci] = w[co][tap] iff ci == co) run in groups of 16 channels (ip[5] = 1): K-chunk kc only feeds
            # accumulator columns [16kc, 16kc+16) with its 16x16 diagonal block, so the weights in shared memory stay
            # tiny.  It wastes 16x the MACs of a pipe that idles anyway on these memory-bound layers and inherits the
            # TMA halo loads / TMA stores (measured: 2-3x the CUDA-core kernel).  The SIMT twin and the emulator execute
            # the same op with the full diagonal matrix.
            dense = torch.zeros(c2, c1, kh, kw)
            dense[torch.arange(c2), torch.arange(c1)] = w[:, 0]
            node = dict(n, g=1, weight=dense, bias=b)
            out = self.conv(node, x, dest, res, res_mode, out_dtype)
            op = self.ops[-1]
            op.ip = (0, 0, 0, 0, 0, 1)
            op.flops = 2 * c2 * kh * kw * Ho * Wo
            op.bytes_w = c2 * kh * kw * 2
            op.bytes_alg = c1 * x.H * x.W * 2 + c2 * Ho * Wo * 2 + op.bytes_w
            return out
        else:
            assert g == c1 == c2, f"{n['name']}: only depthwise grouped convs are on the hot path"
            if dest is None:
                dest = self.full(self.new_buf(c2, Ho, Wo, name=n["name"]), c2)
            assert dest.c == c2 and (dest.H, dest.W) == (Ho, Wo)
            wp = w.reshape(c2, kh * kw).t().contiguous()  # [kh*kw][C]
            op = OpSpec(A.OP_DWCONV, n["name"], [x], dest, res, res_mode, c1, c2, c2, (kh, kw), (sh, sw), (ph, pw), act,
                        A.IMPL_SIMT, self.blob.add(wp.numpy().astype(np.float32)),
                        self.blob.add(b.numpy().astype(np.float32)))
            op.flops = 2 * c2 * kh * kw * Ho * Wo
            op.bytes_w = c2 * kh * kw * 2
            op.bytes_alg = c1 * x.H * x.W * 2 + c2 * Ho * Wo * 2 + op.bytes_w
        self.ops.append(op)
        return dest

    def stem(self, n: Node, dest: View | None) -> View:
        """3-channel 3x3 s2 p1 stem as a 2x2 s1 conv over the space-to-depth input (K = 4 taps x 16), padded before
        only: input row 2*ho-1+r lies in super-row ho-1+R with parity py, (r -> R,py) = (0 -> 0,1), (1 -> 1,0), (2 -> 1,1)."""
        c2 = n["c2"]
        assert n["c1"] == 3 and n["k"] == (3, 3) and n["s"] == (2, 2) and n["p"] == (1, 1) and n["g"] == 1
        Ho, Wo = self.H // 2, self.W // 2
        cpad = _r(c2, 16)
        if dest is None:
            dest = self.full(self.new_buf(c2, Ho, Wo, name=n["name"]), c2)
        rp = ((0, 1), (1, 0), (1, 1))
        wp = torch.zeros(cpad, 2, 2, 4, 4)  # [co][R][S][py*2+px][rgb0]
        for r in range(3):
            for s_ in range(3):
                (R, py), (S, px) = rp[r], rp[s_]
                wp[:c2, R, S, py * 2 + px, :3] = n["weight"][:, :, r, s_]
        bp = torch.zeros(cpad)
        bp[:c2] = n["bias"]
        op = OpSpec(A.OP_CONV, n["name"], [View(self.input_buf, 0, 16, Ho, Wo)], dest, None, A.RES_NONE, 16, c2, cpad,
                    (2, 2), (1, 1), (1, 1), 1 if n["act"] == "silu" else 0, self.conv_impl,
                    self.blob.add(_bf16_bits(wp.reshape(cpad, -1))), self.blob.add(bp.numpy().astype(np.float32)), ip=(1,))
        if self.keep_fp32:
            op.w_fp32 = wp.reshape(cpad, -1).clone()
        op.flops = 2 * c2 * 27 * Ho * Wo
        op.bytes_w = c2 * 27 * 2
        op.bytes_alg = 3 * self.H * self.W * 2 + c2 * Ho * Wo * 2 + op.bytes_w
        self.ops.append(op)
        return dest

    def copy(self, src: View, dest: View, scale=1, name="copy"):
        assert src.c == dest.c and dest.H == src.H * scale and dest.W == src.W * scale
        op = OpSpec(A.OP_RESAMPLE, name, [src], dest, cin=src.c, cout=src.c, ip=(scale,))
        op.bytes_alg = src.c * src.H * src.W * 2 + dest.c * dest.H * dest.W * 2
        self.ops.append(op)
        return dest

    # ------------------------------------------------------------------ blocks
    def bottleneck(self, n, x, dest):
        """Bottleneck (block.py:3405-3419).  The thin ones -- c1 == c2 in {16,32,64}, two 3x3 s1 convs, residual -- run as
        ONE kernel (OP_CONV_PAIR): the hidden map stays in shared memory, the residual is read back from the input tile."""
        cv1, cv2 = n["cv1"], n["cv2"]
        chid = _r(cv1["c2"], 16)
        if (self.fuse_pairs and self.conv_impl != A.IMPL_SIMT and n["add"] and x.real == x.c == cv2["c2"] and x.c in (16, 32, 64)
                and chid in (16, 32, 64) and (x.c // 16, chid // 16) in ((1, 1), (2, 1), (2, 2), (4, 2), (4, 4))
                and all(cv[k] == v for cv in (cv1, cv2) for k, v in (("k", (3, 3)), ("s", (1, 1)), ("p", (1, 1)), ("g", 1), ("act", "silu")))
                and A.lib().b2y_conv_pair_fits(x.c, chid, 64, x.H, x.W)):  # (the fit does not depend on the batch beyond tile counts)
            c1, c2 = x.c, cv2["c2"]
            if dest is None:
                dest = self.full(self.new_buf(c2, x.H, x.W, name=n["name"]), c2)
            assert dest.c == c2 and (dest.H, dest.W) == (x.H, x.W)
            w1 = torch.zeros(chid, 9, c1)
            w1[:cv1["c2"]] = cv1["weight"].permute(0, 2, 3, 1).reshape(cv1["c2"], 9, c1)
            b1 = torch.zeros(chid)
            b1[:cv1["c2"]] = cv1["bias"]
            w2 = torch.zeros(c2, 9, chid)
            w2[:, :, :cv1["c2"]] = cv2["weight"].permute(0, 2, 3, 1).reshape(c2, 9, cv1["c2"])
            op = OpSpec(A.OP_CONV_PAIR, n["name"], [x], dest, None, A.RES_NONE, c1, c2, c2, (3, 3), (1, 1), (1, 1), 1,
                        A.IMPL_TCGEN05, self.blob.add(_bf16_bits(w1.reshape(chid, -1))), self.blob.add(b1.numpy().astype(np.float32)),
                        p_off=(self.blob.add(_bf16_bits(w2.reshape(c2, -1))), self.blob.add(cv2["bias"].numpy().astype(np.float32)), -1, -1),
                        ip=(chid, 1))
            if self.keep_fp32:
                op.w_fp32 = (w1.reshape(chid, -1).clone(), w2.reshape(c2, -1).clone())
            op.flops = 2 * 9 * x.H * x.W * (cv1["c2"] * c1 + c2 * cv1["c2"])
            op.bytes_w = 9 * (cv1["c2"] * c1 + c2 * cv1["c2"]) * 2
            op.bytes_alg = (c1 + c2) * x.H * x.W * 2 + op.bytes_w
            self.ops.append(op)
            return dest
        t = self.conv(cv1, x, pad16=True)
        return self.conv(cv2, t, dest, res=x if n["add"] else None, res_mode=A.RES_ADD if n["add"] else A.RES_NONE)

    def pmsfa(self, n, x, dest):
        """PMSFA (block.py:3035-3053).  Default: conv1 (tensor cores) -> ONE fused depthwise kernel (OP_PMSFA_DW: conv2 5x5 on
        the first half, conv3 7x7 on its first quarter, the cat) -> conv4 (+ x in the epilogue).  B2Y_FUSE_PMSFA=0 or the
        CUDA-core twin: one launch per conv plus the cat copy."""
        c = n["c"]
        cat = self.full(self.new_buf(c, x.H, x.W, name=n["name"] + ".cat"), c)
        fused = (self.conv_impl != A.IMPL_SIMT and os.environ.get("B2Y_FUSE_PMSFA", "1") != "0" and c in (16, 32, 64, 128)
                 and all(n[k]["act"] == "silu" and n[k]["s"] == (1, 1) for k in ("conv2", "conv3"))
                 and n["conv2"]["k"] == (5, 5) and n["conv3"]["k"] == (7, 7))
        if fused:
            t = self.conv(n["conv1"], x)  # [a | b | h]
            c2, c3 = n["conv2"], n["conv3"]
            w2 = c2["weight"].reshape(c // 2, 25).t().contiguous().numpy().astype(np.float32)
            w3 = c3["weight"].reshape(c // 4, 49).t().contiguous().numpy().astype(np.float32)
            op = OpSpec(A.OP_PMSFA_DW, n["name"] + ".dw", [t], cat, cin=c, cout=c, act=1,
                        w_off=self.blob.add(w2), b_off=self.blob.add(c2["bias"].numpy().astype(np.float32)),
                        p_off=(self.blob.add(w3), self.blob.add(c3["bias"].numpy().astype(np.float32)), -1, -1))
            hw = x.H * x.W
            op.flops = 2 * hw * (25 * (c // 2) + 49 * (c // 4))
            op.bytes_w = (25 * (c // 2) + 49 * (c // 4)) * 4
            op.bytes_alg = 2 * c * hw * 2 + op.bytes_w
            self.ops.append(op)
            return self.conv(n["conv4"], cat, dest, res=x, res_mode=A.RES_ADD)
        self.conv(n["conv1"], x, cat)
        t2 = self.conv(n["conv2"], self.sub(cat, 0, c // 2))
        self.conv(n["conv3"], self.sub(t2, 0, c // 4), self.sub(cat, 0, c // 4))
        self.copy(self.sub(t2, c // 4, c // 4), self.sub(cat, c // 4, c // 4), name=n["name"] + ".cat2")
        return self.conv(n["conv4"], cat, dest, res=x, res_mode=A.RES_ADD)

    def inner(self, n, x, dest):
        return {"Bottleneck": self.bottleneck, "PMSFA": self.pmsfa, "C3k": self.c3k}[n["type"]](n, x, dest)

    def c3k(self, n, x, dest):
        """C3k / C3 (block.py:3320-3334): cv3(cat(m(cv1(x)), cv2(x))).  cv1 and cv2 are two 1x1 convs over the same input: with
        at least two inner blocks they run as ONE conv of 2c_ output channels written straight into the concat buffer (the
        last inner block then overwrites the cv1 half, which nothing else reads any more) -- one launch less per block on
        the latency-bound P4/P5 chains."""
        c_ = n["cv1"]["c2"]
        z = self.full(self.new_buf(2 * c_, x.H, x.W, name=n["name"] + ".cat"), 2 * c_)
        cv1, cv2 = n["cv1"], n["cv2"]
        if (len(n["m"]) >= 2 and os.environ.get("B2Y_MERGE_C3K", "0") == "1" and c_ % 8 == 0 and cv2["c2"] == c_
                and all(cv[k] == v for cv in (cv1, cv2) for k, v in (("k", (1, 1)), ("s", (1, 1)), ("g", 1), ("c1", cv1["c1"]),
                                                                      ("act", cv1["act"])))):
            both = dict(cv1, c2=2 * c_, name=cv1["name"] + "+cv2")
            self.conv(both, x, z, wb=(torch.cat([cv1["weight"], cv2["weight"]], 0), torch.cat([cv1["bias"], cv2["bias"]], 0)))
            t = self.sub(z, 0, c_)
            for i, m in enumerate(n["m"]):
                t = self.inner(m, t, self.sub(z, 0, c_) if i == len(n["m"]) - 1 else None)
            return self.conv(n["cv3"], z, dest)
        t = self.conv(n["cv1"], x)
        for i, m in enumerate(n["m"]):
            t = self.inner(m, t, self.sub(z, 0, c_) if i == len(n["m"]) - 1 else None)
        if not n["m"]:
            self.copy(t, self.sub(z, 0, c_))
        self.conv(n["cv2"], x, self.sub(z, c_, c_))
        return self.conv(n["cv3"], z, dest)

    def c3k2(self, n, x, dest, up=None):
        """``up`` = (low, skip, low_first): the block's input is Concat(Upsample2x(low), skip) (yaml head).  cv1 is a 1x1
        conv, so cv1(cat(up(a), b)) = act(up(W_a a) + W_b b + bias): W_a a is computed at half resolution and added,
        upsampled on the fly, before the activation of the conv over ``skip`` (RES_PREADD_UP2) -- the upsampled and the
        concatenated tensors are never written."""
        c, nm = n["c"], len(n["m"])
        if up is not None:
            low, skip, low_first = up
            Hy, Wy = skip.H, skip.W
        else:
            Hy, Wy = x.H, x.W
        y = self.full(self.new_buf((2 + nm) * c, Hy, Wy, name=n["name"] + ".cat"), (2 + nm) * c)
        if up is not None:
            cv1 = n["cv1"]
            w, b = cv1["weight"], cv1["bias"]
            ca, cb = low.real, skip.real
            assert cv1["k"] == (1, 1) and cv1["c1"] == ca + cb and cv1["c2"] == 2 * c
            w_low, w_skip = (w[:, :ca], w[:, ca:]) if low_first else (w[:, cb:], w[:, :cb])
            t = self.conv(dict(cv1, c1=ca, act="none", name=cv1["name"] + ".low"), low, wb=(w_low.contiguous(), torch.zeros_like(b)))
            self.conv(dict(cv1, c1=cb), skip, self.sub(y, 0, 2 * c), res=t, res_mode=A.RES_PREADD_UP2, wb=(w_skip.contiguous(), b))
        else:
            self.conv(n["cv1"], x, self.sub(y, 0, 2 * c))
        for i, m in enumerate(n["m"]):
            self.inner(m, self.sub(y, (1 + i) * c, c), self.sub(y, (2 + i) * c, c))
        return self.conv(n["cv2"], y, dest)

    def sppf(self, n, x, dest):
        c_ = n["cv1"]["c2"]
        if n["k"] != 5:
            raise NotImplementedError("SPPF kernel size other than 5 is not on the B200 path")
        s = self.full(self.new_buf(4 * c_, x.H, x.W, name=n["name"] + ".cat"), 4 * c_)
        self.conv(n["cv1"], x, self.sub(s, 0, c_))
        op = OpSpec(A.OP_SPPF_POOL, n["name"] + ".pool", [self.sub(s, 0, c_)], self.sub(s, c_, 3 * c_), cin=c_, cout=3 * c_)
        op.bytes_alg = 4 * c_ * x.H * x.W * 2
        self.ops.append(op)
        return self.conv(n["cv2"], s, dest)

    def psablock(self, n, x, dest):
        a = n["attn"]
        nh, kd, hd = a["nh"], a["kd"], a["hd"]
        c = nh * hd
        N = x.H * x.W
        qkv = self.conv(a["qkv"], x)
        ao = self.full(self.new_buf(c, x.H, x.W, name=a["name"] + ".av"), c)
        op = OpSpec(A.OP_ATTN, a["name"], [qkv], ao, cin=qkv.c, cout=c, ip=(nh, kd, hd))
        op.flops = 2 * nh * N * N * (kd + hd) + 5 * nh * N * N
        op.bytes_alg = (qkv.c + c) * N * 2
        self.ops.append(op)
        ao2 = self.full(self.new_buf(c, x.H, x.W, name=a["name"] + ".av_pe"), c)
        pe = a["pe"]
        vstride = 2 * kd + hd
        if (self.conv_impl != A.IMPL_SIMT and os.environ.get("B2Y_PE_ONE", "0") == "1" and hd % 8 == 0 and hd <= 64 and vstride % 8 == 0
                and (qkv.coff + 2 * kd) % 8 == 0 and self.bufs[qkv.buf].C % 8 == 0 and pe["k"] == (3, 3) and pe["s"] == (1, 1)):
            # pe(v) + attention output (block.py:4286) as ONE depthwise launch: output channel j reads the v slice of head
            # j // hd of qkv (a chunked input view, dw3_tile.cu) instead of one launch per head
            wp = pe["weight"].reshape(c, 9).t().contiguous()
            op = OpSpec(A.OP_DWCONV, pe["name"], [View(qkv.buf, qkv.coff + 2 * kd, hd, x.H, x.W)], ao2, self.full(ao.buf, c), A.RES_ADD,
                        c, c, c, (3, 3), (1, 1), (1, 1), 1 if pe["act"] == "silu" else 0, A.IMPL_SIMT,
                        self.blob.add(wp.numpy().astype(np.float32)), self.blob.add(pe["bias"].numpy().astype(np.float32)),
                        ip=(hd, vstride))
            op.flops = 2 * c * 9 * x.H * x.W
            op.bytes_w = c * 9 * 2
            op.bytes_alg = 3 * c * x.H * x.W * 2 + op.bytes_w
            self.ops.append(op)
            add = A.RES_ADD if n["add"] else A.RES_NONE
            b1 = self.conv(a["proj"], ao2, res=x if n["add"] else None, res_mode=add)
            f = self.conv(n["ffn"][0], b1)
            return self.conv(n["ffn"][1], f, dest, res=b1 if n["add"] else None, res_mode=add)
        for h in range(nh):  # pe(v): depthwise 3x3 over the v channels of head h, + attention output (block.py:4286)
            sl = slice(h * hd, (h + 1) * hd)
            sub_node = dict(pe, c1=hd, c2=hd, g=hd, name=f"{pe['name']}[h{h}]")
            self.conv(sub_node, self.sub(qkv, h * (2 * kd + hd) + 2 * kd, hd), self.sub(ao2, h * hd, hd),
                      res=self.sub(ao, h * hd, hd), res_mode=A.RES_ADD, wb=(pe["weight"][sl], pe["bias"][sl]))
        add = A.RES_ADD if n["add"] else A.RES_NONE
        b1 = self.conv(a["proj"], ao2, res=x if n["add"] else None, res_mode=add)
        f = self.conv(n["ffn"][0], b1)
        return self.conv(n["ffn"][1], f, dest, res=b1 if n["add"] else None, res_mode=add)

    def c2psa(self, n, x, dest):
        c = n["c"]
        p = self.full(self.new_buf(2 * c, x.H, x.W, name=n["name"] + ".cat"), 2 * c)
        self.conv(n["cv1"], x, p)
        b = self.sub(p, c, c)
        for i, m in enumerate(n["m"]):
            b = self.psablock(m, b, self.sub(p, c, c) if i == len(n["m"]) - 1 else None)
        return self.conv(n["cv2"], p, dest)

    def scdown(self, n, x, dest):
        return self.conv(n["cv2"], self.conv(n["cv1"], x), dest)

    MSCA_STRIPS = (5, 7, 11, 21)

    def msca(self, n, x, dest):
        """MSCAAttention (MSCA.py:52-88).  Default: ONE kernel up to the four branch maps and their global average pools
        (OP_MSCA_BRANCH), then the SE gate, the gate-weighted sum and conv4 (x u in the epilogue).  B2Y_FUSE_MSCA=0, the
        CUDA-core twin, or a map too large for shared memory: one launch per depthwise conv + the pooling kernel."""
        c = n["c"]
        s, t = n["dilconv"]["weight"].reshape(c), n["dilconv"]["bias"]
        strips = [(n[f"conv{b}_1"]["k"], n[f"conv{b}_2"]["k"]) for b in range(4)]
        fused = (self.conv_impl != A.IMPL_SIMT and os.environ.get("B2Y_FUSE_MSCA", "1") != "0" and c % 8 == 0
                 and x.H * x.W <= 1600 and x.real == x.c == c and n["conv0"]["k"] == (5, 5)
                 and strips == [((1, k), (k, 1)) for k in self.MSCA_STRIPS])
        if fused:
            rows = [n["conv0"]["weight"].reshape(c, 25).t(), n["conv0"]["bias"].reshape(1, c)]
            for b, k in enumerate(self.MSCA_STRIPS):
                n1, n2 = n[f"conv{b}_1"], n[f"conv{b}_2"]
                w2, b2 = n2["weight"].reshape(c, k), n2["bias"]
                if b < 3:  # fold the shared per-channel 1x1 "dilconv" (MSCA.py:28-29,57,61,65) into the second strip
                    w2, b2 = w2 * s.view(c, 1), b2 * s + t
                rows += [n1["weight"].reshape(c, k).t(), n1["bias"].reshape(1, c), w2.t(), b2.reshape(1, c)]
            params = torch.cat(rows, 0).contiguous()
            assert params.shape == (122, c)
            br4 = self.full(self.new_buf(4 * c, x.H, x.W, name=n["name"] + ".branches"), 4 * c)
            gap = self.new_buf(4 * c, 1, 1, A.F32, n["name"] + ".gap")
            hw = x.H * x.W
            self.ops.append(OpSpec(A.OP_MSCA_BRANCH, n["name"] + ".branches", [x], br4, res=gap, cin=c, cout=4 * c,
                                   w_off=self.blob.add(params.numpy().astype(np.float32)),
                                   flops=2 * 113 * c * hw, bytes_alg=5 * c * hw * 2 + 122 * c * 4, bytes_w=122 * c * 4))
            br = [self.sub(br4, b * c, c) for b in range(4)]
        else:
            attn = self.conv(n["conv0"], x)
            br = []
            for b in range(4):
                u = self.conv(n[f"conv{b}_1"], attn)
                n2 = n[f"conv{b}_2"]
                wb = None
                if b < 3:  # fold the shared per-channel 1x1 "dilconv" (MSCA.py:28-29,57,61,65) into the strip conv
                    wb = (n2["weight"] * s.view(c, 1, 1, 1), n2["bias"] * s + t)
                br.append(self.conv(n2, u, wb=wb))
            gap = self.new_buf(4 * c, 1, 1, A.F32, n["name"] + ".gap")
            self.ops.append(OpSpec(A.OP_GAP4, n["name"] + ".gap", list(br), gap, cin=c, cout=c,
                                   bytes_alg=4 * c * x.H * x.W * 2))
        gate = self.new_buf(4 * c, 1, 1, A.F32, n["name"] + ".gate")
        se_w = torch.stack([n[f"SE{j}"]["weight"].reshape(c, c) for j in range(1, 5)])
        se_b = torch.stack([n[f"SE{j}"]["bias"] for j in range(1, 5)])
        self.ops.append(OpSpec(A.OP_MSCA_GATE, n["name"] + ".se", [gap], gate, cin=c, cout=c,
                               p_off=(self.blob.add(se_w.numpy().astype(np.float32)),
                                      self.blob.add(se_b.numpy().astype(np.float32)), -1, -1),
                               flops=2 * 4 * c * c, bytes_alg=4 * c * c * 4, bytes_w=4 * c * c * 4))
        xa = self.full(self.new_buf(c, x.H, x.W, name=n["name"] + ".xatt"), c)
        self.ops.append(OpSpec(A.OP_WSUM4, n["name"] + ".wsum", list(br), xa, res=gate, cin=c, cout=c,
                               bytes_alg=5 * c * x.H * x.W * 2))
        return self.conv(n["conv4"], xa, dest, res=x, res_mode=A.RES_MUL)

    def ela(self, n, x, dest):
        c, ks = n["c"], n["ks"]
        if dest is None:
            dest = self.full(self.new_buf(c, x.H, x.W, name=n["name"]), c)
        # banded single-pass statistics (misc_ops.cu:ela_stats_kernel): colm holds per-band partial column means
        banded = c in (64, 128, 256) and x.W * (c // 8) <= 1024 and x.coff % 8 == 0 and self.bufs[x.buf].C % 8 == 0
        nb = -(-x.H // 8) if banded else 0
        rowm = self.new_buf(c, x.H, 1, A.F32, n["name"] + ".rowm")
        colm = self.new_buf(c, x.W, max(nb, 1), A.F32, n["name"] + ".colm")
        self.ops.append(OpSpec(A.OP_ROWCOL_MEAN, n["name"] + ".means", [x, colm], rowm, cin=c, cout=c, ip=(nb,),
                               bytes_alg=c * x.H * x.W * 2))
        hg = self.new_buf(c, x.H, 1, A.F32, n["name"] + ".hgate")
        wg = self.new_buf(c, x.W, 1, A.F32, n["name"] + ".wgate")
        cg = self.new_buf(c, 1, 1, A.F32, n["name"] + ".chgate")
        params = torch.cat([n["sp_w"].reshape(-1), n["gn_w"], n["gn_b"], n["ch_w"].reshape(-1)])
        self.ops.append(OpSpec(A.OP_ELA_GATE, n["name"] + ".gate", [rowm, colm, wg, cg], hg, cin=c, cout=c,
                               p_off=(self.blob.add(params.numpy().astype(np.float32)), -1, -1, -1),
                               ip=(ks, n["groups"], x.H, x.W, nb), fp=(n["gn_eps"],)))
        co = [float(v) for v in n["coef"]]
        self.ops.append(OpSpec(A.OP_ELA_APPLY, n["name"] + ".apply", [x, hg, wg, cg], dest, cin=c, cout=c, fp=tuple(co),
                               bytes_alg=2 * c * x.H * x.W * 2))
        return dest

    def detect(self, n, xs):
        nc = n["nc"]
        no = 64 + nc
        self.A = sum(x.H * x.W for x in xs)
        self.y_buf = self.new_buf(self.A, 4 + nc, 1, A.F32, n["name"] + ".y")
        a_off = 0
        # The fused decode lives in the tcgen05 conv epilogue, which needs cin % 16 == 0.  The box branch has c2 =
        # max(16, ch0/4, 64) and the class branch c3 = max(ch0, min(nc, 100)) hidden channels (head.py:41-42): a width such as
        # 100 (nc >= 100 on an n-scale model) is zero-padded to the next multiple of 16 by the producing conv (pad16); if a
        # level still cannot take the tensor-core path the whole plan falls back to raw head maps + OP_DECODE (decode.cu).
        fuse = self.fuse_decode
        for i, x in enumerate(xs):
            t = self.conv(n["cv2"][i][1], self.conv(n["cv2"][i][0], x), pad16=fuse)
            c3 = n["cv3"][i]
            s = self.conv(c3[0][1], self.conv(c3[0][0], x))
            s = self.conv(c3[1][1], self.conv(c3[1][0], s), pad16=fuse)
            if fuse and (t.c % 16 or s.c % 16):
                raise NotImplementedError(f"{n['name']}: head conv input widths {t.c}/{s.c} cannot be padded to 16")
            if fuse:
                stride = float(n["stride"][i])
                self.conv(n["cv2"][i][2], t, out_dtype=A.F32, decode=(1, a_off, stride))
                self.conv(c3[2], s, out_dtype=A.F32, decode=(2, a_off, stride))
            else:
                hb = self.new_buf(no, x.H, x.W, A.F32, f"{n['name']}.raw{i}")
                self.head_bufs.append(hb)
                head = View(hb, 0, no, x.H, x.W)
                self.conv(n["cv2"][i][2], t, self.sub(head, 0, 64), out_dtype=A.F32)
                self.conv(c3[2], s, self.sub(head, 64, nc), out_dtype=A.F32)
            a_off += x.H * x.W
        if not fuse:
            op = OpSpec(A.OP_DECODE, n["name"] + ".decode", list(self.head_bufs), self.y_buf, cin=no, cout=4 + nc,
                        ip=(nc, len(xs)), fp=tuple(float(s) for s in n["stride"]))
            op.bytes_alg = self.A * (no * 4 + (4 + nc) * 4)
            self.ops.append(op)
        if n.get("cv4"):
            self.extras(n, xs)

    def conv_t2(self, n, x: View) -> View:
        """nn.ConvTranspose2d(c1, c2, 2, 2, 0, bias=True) (Proto.upsample, block.py:91).  Kernel = stride = 2: every output
        pixel (2y+dy, 2x+dx) is ONE tap of input pixel (y, x) -- four 1x1 convs.  They run as one tensor-core 1x1 conv
        with 4*c2 output channels ordered (dy, dx, co), followed by a depth-to-space shuffle (OP_D2S)."""
        c1, c2 = n["c1"], n["c2"]
        assert x.real == x.c == c1 and c2 % 8 == 0
        w = n["weight"].permute(2, 3, 1, 0).reshape(4 * c2, c1, 1, 1).contiguous()  # [(dy, dx, co)][ci]
        node = dict(type="Conv", name=n["name"], c1=c1, c2=4 * c2, k=(1, 1), s=(1, 1), p=(0, 0), g=1, act="none")
        t = self.conv(node, x, wb=(w, n["bias"].repeat(4)))
        out = self.full(self.new_buf(c2, 2 * x.H, 2 * x.W, name=n["name"] + ".d2s"), c2)
        op = OpSpec(A.OP_D2S, n["name"] + ".d2s", [t], out, cin=4 * c2, cout=c2)
        op.bytes_alg = 2 * 4 * c2 * x.H * x.W * 2
        self.ops.append(op)
        return out

    def proto(self, n, x: View):
        """Proto (block.py:80-97): cv3(cv2(upsample(cv1(x)))) on the P3 map -> (B, H/4, W/4, nm) mask prototypes, kept bf16
        NHWC: one pixel = nm contiguous values, the layout b2y_process_mask's dot products want."""
        t = self.conv_t2(n["upsample"], self.conv(n["cv1"], x))
        nm = n["cv3"]["c2"]
        self.proto_buf = self.new_buf(nm, t.H, t.W, name=n["name"])
        self.conv(n["cv3"], self.conv(n["cv2"], t), self.full(self.proto_buf, nm))

    def extras(self, n, xs):
        """The extra branch of a Detect-derived head -- Pose.cv4 (head.py:239-246) / Segment.cv4 (head.py:186,193): Conv(x, c4, 3)
        -> Conv(c4, c4, 3) -> Conv2d(c4, ne, 1) per level; a Segment head also runs Proto on the P3 map.  The reference concatenates the decoded (B, ne, A) rows under y and drags them through
        NMS (ops.py:245-246 `box, cls, mask = x.split((4, nc, nm), 1)`); here the branch stops at its raw fp32 NHWC maps:
        NMS runs on (B, 4+nc, A) as for Detect and b2y_gather_extras decodes only the <= max_det rows NMS kept.
        c4 = max(ch0 // 4, nk) is 51 on the n scale: both hidden maps are zero-padded to 64 channels (pad16)."""
        self.ne = int(n["ne"])
        self.extra_strides = [float(v) for v in n["stride"]]
        self.extra_kind, self.kpt_shape = n["head"], tuple(n.get("kpt_shape", ()))
        for i, x in enumerate(xs):
            c4 = n["cv4"][i]
            t = self.conv(c4[1], self.conv(c4[0], x, pad16=True), pad16=True)
            eb = self.new_buf(_r(self.ne, 4), x.H, x.W, A.F32, f"{n['name']}.extra{i}")
            self.extra_bufs.append(eb)
            self.conv(c4[2], t, View(eb, 0, self.ne, x.H, x.W), out_dtype=A.F32)
        if n["head"] == "segment":
            self.proto(n["proto"], xs[0])

    # ------------------------------------------------------------------ whole graph
    def _lower(self):
        layers = self.ir["layers"]
        # shapes
        shape: dict[int, tuple[int, int, int]] = {}
        prev = (3, self.H, self.W)

        def src_of(i, f):
            return i - 1 if f == -1 else f

        for L in layers:
            i, f, m = L["i"], L["f"], L["module"]
            ty = m["type"]
            ins = [prev if (i == 0 and ff == -1) else shape[src_of(i, ff)] for ff in ([f] if isinstance(f, int) else f)]
            c, h, w = ins[0]
            if ty == "Conv":
                (kh, kw), (sh, sw), (ph, pw) = m["k"], m["s"], m["p"]
                o = (m["c2"], (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1)
            elif ty in ("C3k2", "SPPF", "C2PSA"):
                o = (m["cv2"]["c2"], h, w)
            elif ty == "SCDown":
                o = (m["cv2"]["c2"], h // 2, w // 2)
            elif ty in ("MSCA", "ELA"):
                o = (c, h, w)
            elif ty == "Upsample":
                o = (c, 2 * h, 2 * w)
            elif ty == "Concat":
                o = (sum(s[0] for s in ins), h, w)
            elif ty == "Detect":
                o = None
            else:
                raise NotImplementedError(ty)
            shape[i] = o
        # Upsample -> Concat([up, skip]) -> C3k2 whose cv1 is 1x1, with no other consumers: folded into C3k2.cv1
        consumers: dict[int, list[int]] = {}
        for L in layers:
            for ff in ([L["f"]] if isinstance(L["f"], int) else L["f"]):
                if not (L["i"] == 0 and ff == -1):
                    consumers.setdefault(src_of(L["i"], ff), []).append(L["i"])
        upcat: dict[int, tuple[int, int, bool]] = {}  # C3k2 layer -> (low source layer, skip layer, low first in the cat)
        folded: set[int] = set()
        if self.fold_upsample and self.conv_impl != A.IMPL_SIMT:
            for L in layers:
                i, m = L["i"], L["module"]
                if m["type"] != "C3k2" or not isinstance(L["f"], int) or m["cv1"]["k"] != (1, 1):
                    continue
                jc = src_of(i, L["f"])
                Lc = layers[jc]
                if Lc["module"]["type"] != "Concat" or consumers.get(jc) != [i] or len(Lc["f"]) != 2:
                    continue
                srcs = [src_of(jc, ff) for ff in Lc["f"]]
                ups = [k for k in srcs if layers[k]["module"]["type"] == "Upsample" and consumers.get(k) == [jc]]
                if len(ups) != 1:
                    continue
                ju = ups[0]
                skip = srcs[1] if srcs[0] == ju else srcs[0]
                cl, cs = shape[src_of(ju, layers[ju]["f"])][0], shape[skip][0]
                if layers[skip]["module"]["type"] in ("Concat", "Upsample") or cl % 16 or cs % 16 or m["cv1"]["c2"] % 16:
                    continue
                upcat[i] = (src_of(ju, layers[ju]["f"]), skip, srcs[0] == ju)
                folded.update((ju, jc))
        # concat destinations: the first Concat consumer of a layer gets a direct write.  A Concat that is itself an input of
        # a later Concat (BS-YOLO: layer 13's 384-channel cat is re-used by layer 21, yolo11.yaml:43) lives INSIDE the later
        # one's buffer, so its producers write there directly and the nested cat is never copied.
        dest: dict[int, View] = {}
        cat_view: dict[int, View] = {}
        cats = [L for L in layers if L["module"]["type"] == "Concat" and L["i"] not in folded]
        for L in reversed(cats):  # outer (later) concats first
            i = L["i"]
            C, Hc, Wc = shape[i]
            if i not in cat_view:
                cat_view[i] = self.full(self.new_buf(C, Hc, Wc, name=f"model.{i}.cat"), C)
            off = 0
            for ff in L["f"]:
                j = src_of(i, ff)
                cj = shape[j][0]
                if (layers[j]["module"]["type"] == "Concat" and j not in folded and j not in cat_view and cj % 8 == 0
                        and off % 8 == 0 and os.environ.get("B2Y_NEST_CONCAT", "1") != "0"):
                    cat_view[j] = self.sub(cat_view[i], off, cj)
                off += cj
        for L in cats:
            i = L["i"]
            cv = cat_view[i]
            off = 0
            for ff in L["f"]:
                j = src_of(i, ff)
                cj = shape[j][0]
                if j not in dest and layers[j]["module"]["type"] != "Concat" and cj % 8 == 0 and off % 8 == 0:
                    dest[j] = self.sub(cv, off, cj)
                off += cj
        # emit
        fn = {"C3k2": self.c3k2, "SPPF": self.sppf, "C2PSA": self.c2psa, "SCDown": self.scdown, "MSCA": self.msca,
              "ELA": self.ela}
        for L in layers:
            i, f, m = L["i"], L["f"], L["module"]
            ty = m["type"]
            first = len(self.ops)
            d = dest.get(i)
            if i in folded:  # Upsample / Concat folded into the consuming C3k2.cv1: nothing to launch, nothing to store
                self.layer_op_range[i] = (first, first)
                continue
            if i in upcat:
                lo, sk, low_first = upcat[i]
                self.layer_out[i] = self.c3k2(m, None, d, up=(self.layer_out[lo], self.layer_out[sk], low_first))
                self.layer_op_range[i] = (first, len(self.ops))
                continue
            if ty == "Detect":
                self.detect(m, [self.layer_out[src_of(i, ff)] for ff in f])
            elif ty == "Concat":
                cv, off = cat_view[i], 0
                for ff in f:
                    j = src_of(i, ff)
                    v = self.layer_out[j]
                    tgt = self.sub(cv, off, v.c)
                    if (v.buf, v.coff) != (tgt.buf, tgt.coff):
                        self.copy(v, tgt, name=f"model.{i}.copy{j}")
                    off += v.c
                self.layer_out[i] = cv
            else:
                if i == 0:
                    if ty != "Conv" or m["c1"] != 3:
                        raise NotImplementedError("first layer must be the 3-channel stem Conv")
                    out = self.stem(m, d)
                else:
                    x = self.layer_out[src_of(i, f)]
                    if ty == "Conv":
                        out = self.conv(m, x, d)
                    elif ty == "Upsample":
                        if d is None:
                            d = self.full(self.new_buf(x.c, 2 * x.H, 2 * x.W, name=m["name"]), x.c)
                        out = self.copy(x, d, scale=2, name=m["name"])
                    else:
                        out = fn[ty](m, x, d)
                self.layer_out[i] = out
            self.layer_op_range[i] = (first, len(self.ops))

    # ------------------------------------------------------------------ binding to a batch size
    def layout(self, B: int) -> tuple[list[int], int]:
        """Byte offsets of every buffer for batch B (256-B aligned, no aliasing) and the workspace size."""
        offs, cur = [], 0
        for b in self.bufs:
            offs.append(cur)
            cur = _r(cur + B * b.elems() * (4 if b.dtype == A.F32 else 2), 256)
        return offs, cur

    def c_arrays(self, B: int):
        offs, total = self.layout(B)
        bufs = (A.Buf * len(self.bufs))()
        for i, b in enumerate(self.bufs):
            bufs[i] = A.Buf(offs[i], b.C, b.H, b.W, b.dtype)
        ops = (A.Op * len(self.ops))()
        for i, o in enumerate(self.ops):
            c = A.Op()
            c.kind = o.kind
            for j in range(4):
                c.in_[j], c.in_coff[j] = -1, 0
            for j, v in enumerate(o.ins):
                if isinstance(v, View):
                    c.in_[j], c.in_coff[j] = v.buf, v.coff
                else:
                    c.in_[j] = int(v)
            if isinstance(o.out, View):
                c.out, c.out_coff = o.out.buf, o.out.coff
            else:
                c.out, c.out_coff = int(o.out), 0
            if isinstance(o.res, View):
                c.res, c.res_coff = o.res.buf, o.res.coff
            elif o.res is not None:
                c.res, c.res_coff = int(o.res), 0
            else:
                c.res, c.res_coff = -1, 0
            c.res_mode, c.cin, c.cout, c.cout_pad = o.res_mode, o.cin, o.cout, o.cout_pad
            (c.kh, c.kw), (c.sh, c.sw), (c.ph, c.pw) = o.k, o.s, o.p
            c.act, c.impl, c.w_off, c.b_off = o.act, o.impl, o.w_off, o.b_off
            for j in range(4):
                c.p_off[j] = o.p_off[j] if j < len(o.p_off) else -1
            for j, v in enumerate(o.ip):
                c.ip[j] = int(v)
            for j, v in enumerate(o.fp):
                c.fp[j] = float(v)
            ops[i] = c
        return ops, bufs, total

    # ------------------------------------------------------------------ roofline accounting
    def work_per_image(self) -> tuple[int, int]:
        """(FLOPs, algorithmic bytes) of one image at batch 1 (weights included once)."""
        return sum(o.flops for o in self.ops), sum(o.bytes_alg for o in self.ops)

    @staticmethod
    def op_bytes(o: OpSpec, B: int) -> int:
        """Algorithmic bytes of op `o` for a batch of B images: activations scale with B, weights are read once."""
        return (o.bytes_alg - o.bytes_w) * B + o.bytes_w

    def work_per_batch(self, B: int) -> tuple[int, int]:
        return B * sum(o.flops for o in self.ops), sum(self.op_bytes(o, B) for o in self.ops)

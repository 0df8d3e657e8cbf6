"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement, in plain PyTorch fp32 (NCHW), of the reference's detect forward for the module
classes on the hot path.  It interprets the same IR the CUDA planner consumes
(``bs_yolo_b200.ir.build_ir``), so "CUDA path == oracle" and "oracle == reference" together give
"CUDA path == reference".  Parity is PINNED: ``tests/golden/make_golden.py`` runs the unmodified
reference (imported from /root/reference in the build container) on the seeded checkpoints of
``bs_yolo_b200.arch`` and commits its outputs; ``tests/test_oracle_golden.py`` checks this file against
them, and ``tests/test_oracle_vs_reference.py`` re-checks in-process whenever /root/reference exists.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Each function cites the reference lines it follows (paths relative to /root/reference/ultralytics).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def conv(n, x):
    """nn/modules/conv.py:145-151 with BN pre-folded (utils/torch_utils.py:242-269)."""
    y = F.conv2d(x, n["weight"], n["bias"], stride=n["s"], padding=n["p"], groups=n["g"])
    return F.silu(y) if n["act"] == "silu" else y


def bottleneck(n, x):
    """nn/modules/block.py:3416-3419."""
    y = conv(n["cv2"], conv(n["cv1"], x))
    return x + y if n["add"] else y


def pmsfa(n, x):
    """nn/modules/block.py:3044-3053."""
    c1 = conv(n["conv1"], x)
    c1a, c1b = c1.chunk(2, 1)
    c2 = conv(n["conv2"], c1a)
    c2a, c2b = c2.chunk(2, 1)
    c3 = conv(n["conv3"], c2a)
    return conv(n["conv4"], torch.cat([c3, c2b, c1b], 1)) + x


def _inner(n, x):
    return {"Bottleneck": bottleneck, "PMSFA": pmsfa, "C3k": c3k}[n["type"]](n, x)


def c3k(n, x):
    """nn/modules/block.py:3332-3334 (C3.forward; C3k/C3k_gai only swap self.m)."""
    y = conv(n["cv1"], x)
    for m in n["m"]:
        y = _inner(m, y)
    return conv(n["cv3"], torch.cat((y, conv(n["cv2"], x)), 1))


def c3k2(n, x):
    """nn/modules/block.py:3306-3310 (C2f.forward)."""
    y = list(conv(n["cv1"], x).chunk(2, 1))
    for m in n["m"]:
        y.append(_inner(m, y[-1]))
    return conv(n["cv2"], torch.cat(y, 1))


def sppf(n, x):
    """nn/modules/block.py:3145-3149."""
    k = n["k"]
    y = [conv(n["cv1"], x)]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], k, 1, k // 2))
    return conv(n["cv2"], torch.cat(y, 1))


def attention(n, x):
    """nn/modules/block.py:4277-4288."""
    B, C, H, W = x.shape
    N = H * W
    nh, kd, hd = n["nh"], n["kd"], n["hd"]
    qkv = conv(n["qkv"], x)
    q, k, v = qkv.view(B, nh, kd * 2 + hd, N).split([kd, kd, hd], dim=2)
    attn = (q.transpose(-2, -1) @ k) * (kd ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + conv(n["pe"], v.reshape(B, C, H, W))
    return conv(n["proj"], y)


def psablock(n, x):
    """nn/modules/block.py:4379-4383."""
    x = x + attention(n["attn"], x) if n["add"] else attention(n["attn"], x)
    f = conv(n["ffn"][1], conv(n["ffn"][0], x))
    return x + f if n["add"] else f


def c2psa(n, x):
    """nn/modules/block.py:4464-4468."""
    a, b = conv(n["cv1"], x).split((n["c"], n["c"]), dim=1)
    for m in n["m"]:
        b = psablock(m, b)
    return conv(n["cv2"], torch.cat((a, b), 1))


def scdown(n, x):
    """nn/modules/block.py:4533-4535."""
    return conv(n["cv2"], conv(n["cv1"], x))


def msca(n, x):
    """nn/Addmodules/MSCA.py:52-88."""
    attn = conv(n["conv0"], x)
    br = []
    for b in range(4):
        t = conv(n[f"conv{b}_2"], conv(n[f"conv{b}_1"], attn))
        br.append(conv(n["dilconv"], t) if b < 3 else t)
    w = torch.cat([conv(n[f"SE{b + 1}"], br[b].mean((2, 3), keepdim=True)) for b in range(4)], 2)  # (B,C,4,1)
    w = torch.softmax(torch.sigmoid(w), dim=2)
    x_att = sum(w[:, :, b].unsqueeze(2) * br[b] for b in range(4))
    return conv(n["conv4"], x_att) * x


def ela(n, x):
    """nn/Addmodules/ELA.py:70-100."""
    B, C, H, W = x.shape
    ks = n["ks"]
    ch = torch.sigmoid(F.conv1d(x.mean((2, 3)).unsqueeze(2), n["ch_w"].unsqueeze(1), padding=(ks - 1) // 2, groups=C))

    def gate(v):  # (B,C,L)
        v = F.conv1d(v, n["sp_w"].unsqueeze(1), padding=ks - 1, dilation=2, groups=C)
        return torch.sigmoid(F.group_norm(v, n["groups"], n["gn_w"], n["gn_b"], n["gn_eps"]))

    h_att = gate(x.mean(3)).view(B, C, H, 1)
    w_att = gate(x.mean(2)).view(B, C, 1, W)
    ch_c, sp_c, res_c = n["coef"]
    mask = ch_c * ch.view(B, C, 1, 1) + sp_c * (h_att * w_att)
    return x * mask + res_c * x


def detect_raw(n, xs):
    """nn/modules/head.py:69-70: per level cat(cv2[i](x), cv3[i](x)) -> (B, 64+nc, H, W)."""
    out = []
    for i, x in enumerate(xs):
        b = x
        for c in n["cv2"][i]:
            b = conv(c, b)
        c3 = n["cv3"][i]
        s = conv(c3[0][1], conv(c3[0][0], x))
        s = conv(c3[1][1], conv(c3[1][0], s))
        s = conv(c3[2], s)
        out.append(torch.cat((b, s), 1))
    return out


def make_anchors(shapes, strides, offset=0.5):
    """utils/tal.py:371-383. ``shapes``: list of (h, w). Returns (A,2) anchor points, (A,1) strides."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts), torch.cat(st)


def decode(raws, strides, nc):
    """nn/modules/head.py:100-131 (_inference), nn/modules/block.py:73-76 (DFL), utils/tal.py:386-395 (dist2bbox).

    raws: list of (B, 64+nc, H, W) fp32 -> (B, 4+nc, A) fp32, rows cx,cy,w,h,p0..p{nc-1}."""
    B = raws[0].shape[0]
    no = 64 + nc
    x_cat = torch.cat([r.reshape(B, no, -1) for r in raws], 2)
    anchors, st = make_anchors([r.shape[2:] for r in raws], strides)
    anchors, st = anchors.t().unsqueeze(0), st.t()
    box, cls = x_cat.split((64, nc), 1)
    a = box.shape[-1]
    proj = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    dist = F.conv2d(box.view(B, 4, 16, a).transpose(2, 1).softmax(1), proj).view(B, 4, a)
    lt, rb = dist.chunk(2, 1)
    x1y1, x2y2 = anchors - lt, anchors + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


def pose_kpt_raw(n, xs):
    """nn/modules/head.py:245 (Pose) / :193 (Segment): cat over levels of cv4[i](x[i]).view(bs, ne, -1) -> (B, ne, A) raw
    keypoint maps / mask coefficients."""
    out = []
    for i, x in enumerate(xs):
        t = x
        for c in n["cv4"][i]:
            t = conv(c, t)
        out.append(t.reshape(t.shape[0], n["ne"], -1))
    return torch.cat(out, -1)


def proto(n, x):
    """nn/modules/block.py:95-97 (Proto.forward): cv3(cv2(upsample(cv1(x)))), upsample = ConvTranspose2d(c_, c_, 2, 2, 0)."""
    t = F.conv_transpose2d(conv(n["cv1"], x), n["upsample"]["weight"], n["upsample"]["bias"], stride=2)
    return conv(n["cv3"], conv(n["cv2"], t))


def kpts_decode(kpts, shapes, strides, ndim):
    """nn/modules/head.py:274-279 (Pose.kpts_decode, non-export branch).  kpts (B, nk, A) raw -> decoded, model-input pixels:
    x, y = (v * 2 + (anchor - 0.5)) * stride, visibility = sigmoid(v) when ndim == 3."""
    anchors, st = make_anchors(shapes, strides)
    anchors, st = anchors.t(), st.t()  # (2, A), (1, A): self.anchors / self.strides (head.py:108)
    y = kpts.clone()
    if ndim == 3:
        y[:, 2::3] = y[:, 2::3].sigmoid()
    y[:, 0::ndim] = (y[:, 0::ndim] * 2.0 + (anchors[0] - 0.5)) * st
    y[:, 1::ndim] = (y[:, 1::ndim] * 2.0 + (anchors[1] - 0.5)) * st
    return y


_TOP = {"Conv": conv, "C3k2": c3k2, "SPPF": sppf, "C2PSA": c2psa, "SCDown": scdown, "MSCA": msca, "ELA": ela}


@torch.no_grad()
def forward(ir, x, capture=None):
    """nn/tasks.py:138-165 (_predict_once) + head.py:64-74. x: (B,3,H,W) fp32 in [0,1] RGB.

    Returns (y, raws): y (B,4+nc,A) decoded, raws the three (B,64+nc,H,W) head maps.  A Pose head (head.py:242-250)
    returns (cat(y, decoded keypoints) (B,4+nc+nk,A), (raws, raw keypoints (B,nk,A))); a Segment head (head.py:188-197)
    (cat(y, mask coefficients) (B,4+nc+nm,A), (raws, mask coefficients (B,nm,A), prototypes (B,nm,H/4,W/4))).
    ``capture``: optional dict filled with {layer index: output} for every top-level layer."""
    ys = []
    save = set(ir["save"])
    for L in ir["layers"]:
        m, f = L["module"], L["f"]
        if isinstance(f, int):
            xin = x if f == -1 else ys[f]
        else:
            xin = [x if j == -1 else ys[j] for j in f]
        ty = m["type"]
        if ty == "Upsample":
            x = F.interpolate(xin, scale_factor=2, mode="nearest")
        elif ty == "Concat":
            x = torch.cat(xin, 1)
        elif ty == "Detect":
            raws = detect_raw(m, xin)
            if capture is not None:
                capture[L["i"]] = raws
            y = decode(raws, m["stride"], m["nc"])
            if m.get("head") == "pose":
                kpt = pose_kpt_raw(m, xin)
                pk = kpts_decode(kpt, [r.shape[2:] for r in raws], m["stride"], m["kpt_shape"][1])
                return torch.cat((y, pk), 1), (raws, kpt)
            if m.get("head") == "segment":  # head.py:188-197: (cat(y, mc), (x, mc, p))
                mc = pose_kpt_raw(m, xin)
                return torch.cat((y, mc), 1), (raws, mc, proto(m["proto"], xin[0]))
            return y, raws
        else:
            x = _TOP[ty](m, xin)
        if capture is not None:
            capture[L["i"]] = x
        ys.append(x if L["i"] in save else None)
    raise RuntimeError("graph has no Detect layer")

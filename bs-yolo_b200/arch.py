"""Architecture specs for the detect graphs on the hot path, and seeded synthetic checkpoints.

The real ``yolo11n.pt`` / ``best.pt`` blobs are stripped from the reference tree
(``/root/reference/.MISSING_LARGE_BLOBS``) and ``/root/reference`` does not travel to the GPU
box, so the graphs are re-stated here as *spec trees*: nested dicts whose node types, attribute
names and parameter keys mirror the reference's module classes one-to-one, so that

* a ``state_dict`` generated here loads with ``strict=True`` into the reference's own
  ``DetectionModel`` (used in this container to pin the oracle, ``tests/golden/make_golden.py``);
* ``ir.spec_from_model`` (walking a live, unpickled reference model) yields the same tree.

Reference anchors:
  channel/depth scaling ........ ultralytics/nn/tasks.py:940-1105 (parse_model), utils/ops.py:130-143
  stock yolo11 graph ........... ultralytics/cfg/models/11/yolo11-seg.yaml:15-46 (+ Detect on 16,19,22)
  BS-YOLO graph ................ ultralytics/cfg/models/11/yolo11.yaml:15-52
  Conv / DWConv ................ ultralytics/nn/modules/conv.py:133-151, 224-229
  C2f/C3/C3k/C3k2/Bottleneck ... ultralytics/nn/modules/block.py:3295-3334, 3405-3419, 3796-3815
  PMSFA / C3k_gai / C3k2_gai ... ultralytics/nn/modules/block.py:3035-3095
  SPPF / C2PSA / PSABlock / Attention / SCDown ... block.py:3114-3149, 4235-4288, 4348-4383, 4429-4468, 4503-4535
  MSCAAttention / ELA .......... ultralytics/nn/Addmodules/MSCA.py:19-88, ELA.py:33-100
  Detect / DFL ................. ultralytics/nn/modules/head.py:21-74, block.py:58-76
  Pose .......... ultralytics/nn/modules/head.py:230-250, cfg/models/11/yolo11-pose.yaml:15-47
  Segment / Proto (§8 f4) ...... ultralytics/nn/modules/head.py:175-197, block.py:80-97, cfg/models/11/yolo11-seg.yaml:15-47
"""
from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch

Node = dict[str, Any]

SCALES = {  # depth, width, max_channels (yolo11.yaml:8-12)
    "n": (0.50, 0.25, 1024),
    "s": (0.50, 0.50, 1024),
    "m": (0.50, 1.00, 512),
    "l": (1.00, 1.00, 512),
    "x": (1.00, 1.50, 512),
}


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _div8(x: float) -> int:
    return int(math.ceil(x / 8) * 8)


# ----------------------------------------------------------------------------- leaf nodes
def conv_bn(name, c1, c2, k=1, s=1, g=1, act=True) -> Node:
    """`Conv` module: bias-free conv + BatchNorm(eps 1e-3) + SiLU|Identity, 'same' padding."""
    kh, kw = _pair(k)
    return dict(type="Conv", name=name, c1=c1, c2=c2, k=(kh, kw), s=_pair(s), p=(kh // 2, kw // 2), g=g,
                act="silu" if act else "none", w=f"{name}.conv.weight", b=None, bn=f"{name}.bn", bn_eps=1e-3)


def conv_plain(name, c1, c2, k=1, p=0, g=1) -> Node:
    """Bare `nn.Conv2d` with bias, no norm, no activation."""
    return dict(type="Conv", name=name, c1=c1, c2=c2, k=_pair(k), s=(1, 1), p=_pair(p), g=g,
                act="none", w=f"{name}.weight", b=f"{name}.bias", bn=None, bn_eps=0.0)


def conv_t2(name, c1, c2) -> Node:
    """`nn.ConvTranspose2d(c1, c2, 2, 2, 0, bias=True)` (Proto.upsample, block.py:91): weight (c1, c2, 2, 2)."""
    return dict(type="ConvT", name=name, c1=c1, c2=c2, k=(2, 2), s=(2, 2), w=f"{name}.weight", b=f"{name}.bias")


# ----------------------------------------------------------------------------- blocks
def bottleneck(name, c1, c2, shortcut=True, k=(3, 3), e=0.5) -> Node:
    c_ = int(c2 * e)
    return dict(type="Bottleneck", name=name, cv1=conv_bn(f"{name}.cv1", c1, c_, k[0]),
                cv2=conv_bn(f"{name}.cv2", c_, c2, k[1]), add=bool(shortcut and c1 == c2))


def pmsfa(name, inc) -> Node:
    return dict(type="PMSFA", name=name, c=inc,
                conv1=conv_bn(f"{name}.conv1", inc, inc, 3),
                conv2=conv_bn(f"{name}.conv2", inc // 2, inc // 2, 5, g=inc // 2),
                conv3=conv_bn(f"{name}.conv3", inc // 4, inc // 4, 7, g=inc // 4),
                conv4=conv_bn(f"{name}.conv4", inc, inc, 1))


def c3k(name, c1, c2, n=2, shortcut=True, e=0.5, k=3, gai=False) -> Node:
    c_ = int(c2 * e)
    inner = [pmsfa(f"{name}.m.{i}", c_) if gai else bottleneck(f"{name}.m.{i}", c_, c_, shortcut, (k, k), 1.0)
             for i in range(n)]
    return dict(type="C3k", name=name, cv1=conv_bn(f"{name}.cv1", c1, c_, 1), cv2=conv_bn(f"{name}.cv2", c1, c_, 1),
                cv3=conv_bn(f"{name}.cv3", 2 * c_, c2, 1), m=inner)


def c3k2(name, c1, c2, n=1, use_c3k=False, e=0.5, shortcut=True, gai=False) -> Node:
    c = int(c2 * e)
    inner = []
    for i in range(n):
        if use_c3k:
            inner.append(c3k(f"{name}.m.{i}", c, c, 2, shortcut, gai=gai))
        elif gai:
            inner.append(pmsfa(f"{name}.m.{i}", c))
        else:
            inner.append(bottleneck(f"{name}.m.{i}", c, c, shortcut))
    return dict(type="C3k2", name=name, c=c, cv1=conv_bn(f"{name}.cv1", c1, 2 * c, 1),
                cv2=conv_bn(f"{name}.cv2", (2 + n) * c, c2, 1), m=inner)


def sppf(name, c1, c2, k=5) -> Node:
    c_ = c1 // 2
    return dict(type="SPPF", name=name, k=k, cv1=conv_bn(f"{name}.cv1", c1, c_, 1),
                cv2=conv_bn(f"{name}.cv2", 4 * c_, c2, 1))


def psablock(name, c, num_heads) -> Node:
    hd = c // num_heads
    kd = int(hd * 0.5)
    h = c + num_heads * kd * 2
    attn = dict(type="Attention", name=f"{name}.attn", nh=num_heads, kd=kd, hd=hd,
                qkv=conv_bn(f"{name}.attn.qkv", c, h, 1, act=False),
                proj=conv_bn(f"{name}.attn.proj", c, c, 1, act=False),
                pe=conv_bn(f"{name}.attn.pe", c, c, 3, g=c, act=False))
    return dict(type="PSABlock", name=name, attn=attn, add=True,
                ffn=[conv_bn(f"{name}.ffn.0", c, 2 * c, 1), conv_bn(f"{name}.ffn.1", 2 * c, c, 1, act=False)])


def c2psa(name, c1, c2, n=1, e=0.5) -> Node:
    assert c1 == c2
    c = int(c1 * e)
    return dict(type="C2PSA", name=name, c=c, cv1=conv_bn(f"{name}.cv1", c1, 2 * c, 1),
                cv2=conv_bn(f"{name}.cv2", 2 * c, c1, 1), m=[psablock(f"{name}.m.{i}", c, c // 64) for i in range(n)])


def scdown(name, c1, c2, k, s) -> Node:
    return dict(type="SCDown", name=name, cv1=conv_bn(f"{name}.cv1", c1, c2, 1),
                cv2=conv_bn(f"{name}.cv2", c2, c2, k, s, g=c2, act=False))


def msca(name, dim) -> Node:
    n = dict(type="MSCA", name=name, c=dim, conv0=conv_plain(f"{name}.conv0", dim, dim, 5, 2, dim),
             dilconv=conv_plain(f"{name}.dilconv", dim, dim, 1, 0, dim), conv4=conv_plain(f"{name}.conv4", dim, dim, 1))
    for b, k in enumerate((5, 7, 11, 21)):
        n[f"conv{b}_1"] = conv_plain(f"{name}.conv{b}_1", dim, dim, (1, k), (0, k // 2), dim)
        n[f"conv{b}_2"] = conv_plain(f"{name}.conv{b}_2", dim, dim, (k, 1), (k // 2, 0), dim)
        n[f"SE{b + 1}"] = conv_plain(f"{name}.SE{b + 1}.conv.0", dim, dim, 1)
    return n


def ela(name, channel, b=1, gamma=2) -> Node:
    ks = int(abs((math.log(channel, 2) + b) / gamma))
    ks = ks if ks % 2 else ks + 1
    return dict(type="ELA", name=name, c=channel, ks=ks, groups=max(1, channel // 16), gn_eps=1e-5)


def detect(name, nc, ch) -> Node:
    c2, c3 = max(16, ch[0] // 4, 64), max(ch[0], min(nc, 100))
    cv2, cv3 = [], []
    for i, x in enumerate(ch):
        cv2.append([conv_bn(f"{name}.cv2.{i}.0", x, c2, 3), conv_bn(f"{name}.cv2.{i}.1", c2, c2, 3),
                    conv_plain(f"{name}.cv2.{i}.2", c2, 64, 1)])
        cv3.append([[conv_bn(f"{name}.cv3.{i}.0.0", x, x, 3, g=x), conv_bn(f"{name}.cv3.{i}.0.1", x, c3, 1)],
                    [conv_bn(f"{name}.cv3.{i}.1.0", c3, c3, 3, g=c3), conv_bn(f"{name}.cv3.{i}.1.1", c3, c3, 1)],
                    conv_plain(f"{name}.cv3.{i}.2", c3, nc, 1)])
    return dict(type="Detect", name=name, nc=nc, reg_max=16, no=nc + 64, nl=len(ch), ch=list(ch), cv2=cv2, cv3=cv3,
                stride=[8.0, 16.0, 32.0][: len(ch)])


def pose(name, nc, ch, kpt_shape=(17, 3)) -> Node:
    """`Pose(Detect)` (head.py:230-250): the Detect branches plus one keypoint branch per level, cv4[i] =
    Conv(x, c4, 3) -> Conv(c4, c4, 3) -> Conv2d(c4, nk, 1) with c4 = max(ch[0] // 4, nk).  Kept as a Detect node with
    ``head="pose"`` so that everything that only needs the Detect part treats it as one."""
    n = detect(name, nc, ch)
    nk = int(kpt_shape[0]) * int(kpt_shape[1])
    c4 = max(ch[0] // 4, nk)
    n.update(head="pose", kpt_shape=(int(kpt_shape[0]), int(kpt_shape[1])), ne=nk,
             cv4=[[conv_bn(f"{name}.cv4.{i}.0", x, c4, 3), conv_bn(f"{name}.cv4.{i}.1", c4, c4, 3),
                   conv_plain(f"{name}.cv4.{i}.2", c4, nk, 1)] for i, x in enumerate(ch)])
    return n


def segment(name, nc, ch, nm=32, npr=256) -> Node:
    """`Segment(Detect)` (head.py:175-197): the Detect branches, Proto(ch[0], npr, nm) on the P3 map (block.py:80-97:
    Conv 3x3 -> ConvTranspose2d 2x2 s2 -> Conv 3x3 -> Conv 1x1) and one mask-coefficient branch per level, cv4[i] =
    Conv(x, c4, 3) -> Conv(c4, c4, 3) -> Conv2d(c4, nm, 1) with c4 = max(ch[0] // 4, nm)."""
    n = detect(name, nc, ch)
    c4 = max(ch[0] // 4, nm)
    n.update(head="segment", nm=nm, npr=npr, ne=nm,
             proto=dict(type="Proto", name=f"{name}.proto", cv1=conv_bn(f"{name}.proto.cv1", ch[0], npr, 3),
                        upsample=conv_t2(f"{name}.proto.upsample", npr, npr), cv2=conv_bn(f"{name}.proto.cv2", npr, npr, 3),
                        cv3=conv_bn(f"{name}.proto.cv3", npr, nm, 1)),
             cv4=[[conv_bn(f"{name}.cv4.{i}.0", x, c4, 3), conv_bn(f"{name}.cv4.{i}.1", c4, c4, 3),
                   conv_plain(f"{name}.cv4.{i}.2", c4, nm, 1)] for i, x in enumerate(ch)])
    return n


# ----------------------------------------------------------------------------- whole graphs
_STOCK = [  # (from, repeats, module, args)  -- yolo11-seg.yaml:15-46 with Detect as the last layer
    (-1, 1, "Conv", [64, 3, 2]), (-1, 1, "Conv", [128, 3, 2]), (-1, 2, "C3k2", [256, False, 0.25]),
    (-1, 1, "Conv", [256, 3, 2]), (-1, 2, "C3k2", [512, False, 0.25]), (-1, 1, "Conv", [512, 3, 2]),
    (-1, 2, "C3k2", [512, True]), (-1, 1, "Conv", [1024, 3, 2]), (-1, 2, "C3k2", [1024, True]),
    (-1, 1, "SPPF", [1024, 5]), (-1, 2, "C2PSA", [1024]),
    (-1, 1, "Upsample", []), ([-1, 6], 1, "Concat", []), (-1, 2, "C3k2", [512, False]),
    (-1, 1, "Upsample", []), ([-1, 4], 1, "Concat", []), (-1, 2, "C3k2", [256, False]),
    (-1, 1, "Conv", [256, 3, 2]), ([-1, 13], 1, "Concat", []), (-1, 2, "C3k2", [512, False]),
    (-1, 1, "Conv", [512, 3, 2]), ([-1, 10], 1, "Concat", []), (-1, 2, "C3k2", [1024, True]),
    ([16, 19, 22], 1, "Detect", []),
]

_BSYOLO = [  # yolo11.yaml:15-52 of the fork; absolute `from` indices kept verbatim
    (-1, 1, "Conv", [64, 3, 2]), (-1, 1, "Conv", [128, 3, 2]), (-1, 2, "C3k2_gai", [256, False, 0.25]),
    (-1, 1, "Conv", [256, 3, 2]), (-1, 2, "C3k2_gai", [512, False, 0.25]), (-1, 1, "SCDown", [512, 3, 2]),
    (-1, 2, "C3k2_gai", [512, True]), (-1, 1, "SCDown", [1024, 3, 2]), (-1, 2, "C3k2_gai", [1024, True]),
    (-1, 1, "SPPF", [1024, 5]), (-1, 2, "C2PSA", [1024]), (-1, 1, "MSCAAttention", []),
    (-1, 1, "Upsample", []), ([-1, 6], 1, "Concat", []), (-1, 2, "C3k2", [512, False]), (-1, 1, "ELA", [512]),
    (-1, 1, "Upsample", []), ([-1, 4], 1, "Concat", []), (-1, 2, "C3k2", [256, False]), (-1, 1, "ELA", [256]),
    (-1, 1, "Conv", [256, 3, 2]), ([-1, 13], 1, "Concat", []), (-1, 2, "C3k2", [512, False]), (-1, 1, "ELA", [512]),
    (-1, 1, "SCDown", [512, 3, 2]), ([-1, 10], 1, "Concat", []), (-1, 2, "C3k2", [1024, True]), (-1, 1, "ELA", [1024]),
    ([19, 23, 27], 1, "Detect", []),
]

_STOCK_POSE = _STOCK[:-1] + [([16, 19, 22], 1, "Pose", [(17, 3)])]  # yolo11-pose.yaml:15-47 (same backbone / neck)

_STOCK_SEG = _STOCK[:-1] + [([16, 19, 22], 1, "Segment", [32, 256])]  # yolo11-seg.yaml:47 as shipped

HEAD_GAIN = 8.0
SYNTH_CLS_BIAS = {  # model -> per-level bias shift of the final class conv (seed 0 calibration, see synth_state_dict)
    "yolo11n": (-3.678, -4.087, -3.846),
    "bsyolo_n": (-2.583, -3.309, -3.207),
    "yolo11x": (-3.446, -3.569, -3.91),
    "yolo11n-pose": (-3.678, -4.087, -3.846),  # cv4 is drawn after every Detect parameter: the Detect part equals yolo11n's
    "yolo11n-seg": (-3.678, -4.087, -3.846),   # likewise (proto and cv4 come last)
}

MODELS = {  # name -> (table, scale, nc)
    "yolo11n": (_STOCK, "n", 80),
    "yolo11s": (_STOCK, "s", 80),
    "yolo11m": (_STOCK, "m", 80),
    "yolo11l": (_STOCK, "l", 80),
    "yolo11x": (_STOCK, "x", 80),
    "bsyolo_n": (_BSYOLO, "n", 12),
    "yolo11n-pose": (_STOCK_POSE, "n", 80),  # nc: 80 and kpt_shape: [17, 3] are the yaml's own defaults
    "yolo11n-seg": (_STOCK_SEG, "n", 80),
}


def build_spec(model: str, nc: int | None = None) -> Node:
    """Instantiate the layer table the way parse_model does (tasks.py:962-1103)."""
    table, scale, nc0 = MODELS[model]
    nc = nc0 if nc is None else nc
    depth, width, max_ch = SCALES[scale]
    ch: list[int] = []
    layers, save = [], []
    c_prev = 3
    for i, (f, n, kind, args) in enumerate(table):
        n = max(round(n * depth), 1) if n > 1 else n
        name = f"model.{i}"
        c1 = c_prev if isinstance(f, int) and f == -1 else (ch[f] if isinstance(f, int) else None)
        if kind in ("Conv", "C3k2", "C3k2_gai", "SPPF", "C2PSA", "SCDown"):
            c2 = _div8(min(args[0], max_ch) * width)
            if kind == "Conv":
                node = conv_bn(name, c1, c2, args[1], args[2])
            elif kind in ("C3k2", "C3k2_gai"):
                use_c3k = bool(args[1])
                if kind == "C3k2" and scale in "mlx":
                    use_c3k = True  # tasks.py:1046-1049
                e = args[2] if len(args) > 2 else 0.5
                node = c3k2(name, c1, c2, n, use_c3k, e, gai=(kind == "C3k2_gai"))
            elif kind == "SPPF":
                node = sppf(name, c1, c2, args[1])
            elif kind == "C2PSA":
                node = c2psa(name, c1, c2, n)
            else:
                node = scdown(name, c1, c2, args[1], args[2])
        elif kind == "MSCAAttention":
            c2, node = c1, msca(name, c1)
        elif kind == "ELA":
            c2, node = c1, ela(name, c1)
        elif kind == "Upsample":
            c2, node = c1, dict(type="Upsample", name=name, scale=2)
        elif kind == "Concat":
            c2, node = sum(ch[x] if x != -1 else c_prev for x in f), dict(type="Concat", name=name)
        elif kind == "Detect":
            c2, node = None, detect(name, nc, [ch[x] for x in f])
        elif kind == "Pose":
            c2, node = None, pose(name, nc, [ch[x] for x in f], args[0])
        elif kind == "Segment":  # tasks.py:1079-1080: npr is width-scaled like a channel count
            c2, node = None, segment(name, nc, [ch[x] for x in f], args[0], _div8(min(args[1], max_ch) * width))
        else:
            raise ValueError(kind)
        fl = [f] if isinstance(f, int) else list(f)
        save.extend(x % i for x in fl if x != -1)
        layers.append(dict(i=i, f=f, module=node))
        ch.append(c2)
        c_prev = c2
    return dict(model=model, nc=nc, layers=layers, save=sorted(set(save)), names={k: f"class{k}" for k in range(nc)})


# ----------------------------------------------------------------------------- synthetic weights
def iter_nodes(node):
    """Depth-first walk over every dict node of a spec tree (deterministic order)."""
    if isinstance(node, dict):
        if "type" in node:
            yield node
        for k in node:
            if k != "names":
                yield from iter_nodes(node[k])
    elif isinstance(node, (list, tuple)):
        for v in node:
            yield from iter_nodes(v)


def synth_state_dict(spec: Node, seed: int = 0) -> dict[str, torch.Tensor]:
    """Seeded random parameters for every node, keyed exactly like the reference's ``state_dict()``.

    Weights are variance-preserving through SiLU so ~25 stacked layers neither blow up nor vanish;
    BN statistics are randomised (the default (0,1,1,0) would make the fold a no-op and hide bugs)."""
    rng = np.random.default_rng(seed)
    sd: dict[str, torch.Tensor] = {}

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    # convs whose output is added to a skip connection get a small gain so the residual stream's variance
    # stays O(1) through the deep (l/x) graphs
    damped = set()
    for n in iter_nodes(spec):
        if n["type"] == "Bottleneck" and n["add"]:
            damped.add(n["cv2"]["name"])
        elif n["type"] == "PMSFA":
            damped.add(n["conv4"]["name"])
        elif n["type"] == "PSABlock":
            damped.update((n["attn"]["proj"]["name"], n["ffn"][1]["name"]))

    for n in iter_nodes(spec):
        ty = n["type"]
        if ty == "Conv":
            c1, c2, (kh, kw), g = n["c1"], n["c2"], n["k"], n["g"]
            fan_in = (c1 // g) * kh * kw
            gain = 1.55 if n["act"] == "silu" else 1.0
            if n["name"] in damped:
                gain *= 0.35
            sd[n["w"]] = t(rng.standard_normal((c2, c1 // g, kh, kw)) * (gain / math.sqrt(fan_in)))
            if n["b"]:
                sd[n["b"]] = t(rng.standard_normal(c2) * 0.1)
            if n["bn"]:
                sd[n["bn"] + ".weight"] = t(rng.uniform(0.8, 1.2, c2))
                sd[n["bn"] + ".bias"] = t(rng.standard_normal(c2) * 0.1)
                sd[n["bn"] + ".running_mean"] = t(rng.standard_normal(c2) * 0.1)
                sd[n["bn"] + ".running_var"] = t(rng.uniform(0.8, 1.2, c2))
                sd[n["bn"] + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
        elif ty == "ConvT":
            c1, c2 = n["c1"], n["c2"]
            sd[n["w"]] = t(rng.standard_normal((c1, c2, 2, 2)) * (1.0 / math.sqrt(c1)))
            sd[n["b"]] = t(rng.standard_normal(c2) * 0.1)
        elif ty == "ELA":
            c, ks, name = n["c"], n["ks"], n["name"]
            sd[f"{name}.ch_att.2.weight"] = t(rng.standard_normal((c, 1, ks)) / math.sqrt(ks))
            sd[f"{name}.spatial_conv.weight"] = t(rng.standard_normal((c, 1, ks)) / math.sqrt(ks))
            for k in ("ch_weight", "sp_weight", "res_weight"):
                sd[f"{name}.{k}"] = t(rng.standard_normal(1) * 0.5)
            sd[f"{name}.gn.weight"] = t(rng.uniform(0.8, 1.2, c))
            sd[f"{name}.gn.bias"] = t(rng.standard_normal(c) * 0.1)
        elif ty == "Detect":
            sd[f"{n['name']}.dfl.conv.weight"] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    # Detect output convs: widen the logit range (DFL bins / class scores would otherwise be near-uniform) and
    # re-bias the class branch so scores straddle conf=0.25 (head.py:133-144 bias_init leaves them ~1e-4, i.e. zero
    # candidates and an NMS that does no work -- SURVEY §8c caveat (i)). Shifts were calibrated once with the oracle
    # on uniform-random uint8 images so that ~6 % of anchors per level clear the threshold.
    det = spec["layers"][-1]["module"]
    shifts = SYNTH_CLS_BIAS.get(spec["model"], (-3.0, -3.0, -3.0))
    for i in range(det["nl"]):
        for br in ("cv2", "cv3"):
            sd[f"{det['name']}.{br}.{i}.2.weight"] *= HEAD_GAIN
        sd[f"{det['name']}.cv3.{i}.2.bias"] += float(shifts[i])
    return sd

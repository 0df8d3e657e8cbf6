"""Graph IR: spec tree + BatchNorm-folded fp32 parameters.

``build_ir`` is the boundary between "weight loading unchanged" and this runtime: it consumes a
spec tree (``arch.build_spec`` or ``spec_from_model`` over a live, unpickled reference
``DetectionModel``) plus that model's ``state_dict()`` and folds every BatchNorm into its
convolution itself -- the reference only folds modules that pass its isinstance test
(ultralytics/nn/tasks.py:200-228), which the fork's class shadowing defeats for most of
``block.py``, so both fused and unfused modules must be handled.

Fold (ultralytics/utils/torch_utils.py:242-269):  W' = W * g/sqrt(var+eps),  b' = beta - g*mean/sqrt(var+eps) (+ b*scale)
"""
from __future__ import annotations

import copy
from typing import Any

import torch

from .arch import iter_nodes

Node = dict[str, Any]


def build_ir(spec: Node, sd: dict[str, torch.Tensor]) -> Node:
    """Return a deep copy of ``spec`` with ``weight``/``bias`` (folded, fp32, CPU) on every Conv node
    and the small parameter tensors on ELA nodes."""
    ir = copy.deepcopy(spec)
    for n in iter_nodes(ir):
        ty = n["type"]
        if ty == "Conv":
            w = sd[n["w"]].detach().float().cpu()
            b = sd[n["b"]].detach().float().cpu() if n["b"] else torch.zeros(w.shape[0])
            if n["bn"]:
                g = sd[n["bn"] + ".weight"].detach().float().cpu()
                beta = sd[n["bn"] + ".bias"].detach().float().cpu()
                mu = sd[n["bn"] + ".running_mean"].detach().float().cpu()
                var = sd[n["bn"] + ".running_var"].detach().float().cpu()
                scale = g / torch.sqrt(var + n["bn_eps"])
                w = w * scale.view(-1, 1, 1, 1)
                b = beta - mu * scale + b * scale
            assert tuple(w.shape) == (n["c2"], n["c1"] // n["g"], *n["k"]), (n["name"], w.shape)
            n["weight"], n["bias"] = w.contiguous(), b.contiguous()
        elif ty == "ConvT":
            n["weight"] = sd[n["w"]].detach().float().cpu().contiguous()  # (c1, c2, 2, 2)
            n["bias"] = sd[n["b"]].detach().float().cpu().contiguous()
            assert tuple(n["weight"].shape) == (n["c1"], n["c2"], 2, 2), (n["name"], n["weight"].shape)
        elif ty == "ELA":
            p = n["name"]
            n["ch_w"] = sd[f"{p}.ch_att.2.weight"].detach().float().cpu().reshape(n["c"], n["ks"]).contiguous()
            n["sp_w"] = sd[f"{p}.spatial_conv.weight"].detach().float().cpu().reshape(n["c"], n["ks"]).contiguous()
            n["gn_w"] = sd[f"{p}.gn.weight"].detach().float().cpu().contiguous()
            n["gn_b"] = sd[f"{p}.gn.bias"].detach().float().cpu().contiguous()
            n["coef"] = torch.stack([sd[f"{p}.{k}"].detach().float().cpu().reshape(())
                                     for k in ("ch_weight", "sp_weight", "res_weight")]).sigmoid()
    return ir


def detect_node(ir: Node) -> Node:
    return ir["layers"][-1]["module"]


def head_kind(ir: Node) -> str:
    """"detect", "pose" or "segment": which Detect-derived head the graph ends in (head.py:21, 175, 230)."""
    return detect_node(ir).get("head", "detect")


def require_head(ir: Node, want: str | None, who: str) -> None:
    """A predictor / validator class post-processes ONE head kind (the reference picks the class by task,
    models/yolo/model.py task_map); pairing it with another head would silently drop the keypoints / masks (or look for
    ones that do not exist), so the mismatch raises and names the class to use."""
    kind = head_kind(ir)
    if want is not None and kind != want:
        use = {"detect": "B200DetectionPredictor", "pose": "B200PosePredictor", "segment": "B200SegmentationPredictor"}[kind]
        raise NotImplementedError(f"{who} post-processes a {want} head, but the model ends in a {kind} head: use "
                                  f"bs_yolo_b200.predictor.{use}" + (" (no B200 validator exists for this head)" if "alidator" in who and kind != "detect" else ""))


def calibrate_cls_bias(ir: Node, head_logits: list[torch.Tensor], frac: float = 0.06, conf: float = 0.25) -> None:
    """Shift the class-branch bias of each level so that ``frac`` of its anchors clear ``conf``.

    A default-initialised Detect head scores ~1e-4 everywhere (head.py:133-144 bias_init), i.e. zero
    candidates at conf 0.25 and an NMS that does no work; synthetic checkpoints are re-biased so that
    scores straddle the threshold (SURVEY §8c weights caveat (i)).  ``head_logits[i]``: (B, no, H, W) raw
    maps from a forward with the *current* biases."""
    det = detect_node(ir)
    target = float(torch.logit(torch.tensor(conf)))
    for i, raw in enumerate(head_logits):
        m = raw[:, 64:].amax(1).flatten()
        q = torch.quantile(m, 1.0 - frac)
        det["cv3"][i][2]["bias"] = det["cv3"][i][2]["bias"] + (target - float(q))


# ----------------------------------------------------------------------------- live reference model -> spec
def _act_name(m) -> str:
    name = type(m.act).__name__
    if name == "SiLU":
        return "silu"
    if name == "Identity":
        return "none"
    raise NotImplementedError(f"activation {name} is not on the B200 path")


def _conv_node(m, name: str) -> Node:
    """`Conv`/`DWConv` (conv.py:133-151, Addmodules/conv.py) or a bare nn.Conv2d."""
    tn = type(m).__name__
    if tn == "Conv2d":
        conv, bn, act = m, None, "none"
        wkey, bkey = f"{name}.weight", (f"{name}.bias" if m.bias is not None else None)
    elif tn in ("Conv", "DWConv"):
        conv, bn, act = m.conv, getattr(m, "bn", None), _act_name(m)
        wkey, bkey = f"{name}.conv.weight", (f"{name}.conv.bias" if conv.bias is not None else None)
    else:
        raise NotImplementedError(f"{tn} at {name}")
    if tuple(conv.dilation) != (1, 1) and tuple(conv.kernel_size) != (1, 1):
        raise NotImplementedError(f"dilated conv at {name}")
    if isinstance(conv.padding, str):
        raise NotImplementedError(f"string padding at {name}")
    return dict(type="Conv", name=name, c1=conv.in_channels, c2=conv.out_channels, k=tuple(conv.kernel_size),
                s=tuple(conv.stride), p=tuple(conv.padding), g=conv.groups, act=act, w=wkey, b=bkey,
                bn=(f"{name}.bn" if bn is not None else None), bn_eps=(float(bn.eps) if bn is not None else 0.0))


def _module_node(m, name: str) -> Node:
    tn = type(m).__name__
    C = _conv_node
    if tn in ("Conv", "DWConv", "Conv2d"):
        return C(m, name)
    if tn == "Bottleneck":
        return dict(type="Bottleneck", name=name, cv1=C(m.cv1, f"{name}.cv1"), cv2=C(m.cv2, f"{name}.cv2"), add=bool(m.add))
    if tn == "PMSFA":
        return dict(type="PMSFA", name=name, c=m.conv1.conv.in_channels,
                    **{k: C(getattr(m, k), f"{name}.{k}") for k in ("conv1", "conv2", "conv3", "conv4")})
    if tn in ("C3k", "C3", "C3k_gai"):
        return dict(type="C3k", name=name, cv1=C(m.cv1, f"{name}.cv1"), cv2=C(m.cv2, f"{name}.cv2"),
                    cv3=C(m.cv3, f"{name}.cv3"), m=[_module_node(x, f"{name}.m.{i}") for i, x in enumerate(m.m)])
    if tn in ("C3k2", "C2f", "C3k2_gai"):
        return dict(type="C3k2", name=name, c=int(m.c), cv1=C(m.cv1, f"{name}.cv1"), cv2=C(m.cv2, f"{name}.cv2"),
                    m=[_module_node(x, f"{name}.m.{i}") for i, x in enumerate(m.m)])
    if tn == "SPPF":
        k = m.m.kernel_size if isinstance(m.m.kernel_size, int) else m.m.kernel_size[0]
        return dict(type="SPPF", name=name, k=int(k), cv1=C(m.cv1, f"{name}.cv1"), cv2=C(m.cv2, f"{name}.cv2"))
    if tn == "PSABlock":
        a = m.attn
        attn = dict(type="Attention", name=f"{name}.attn", nh=int(a.num_heads), kd=int(a.key_dim), hd=int(a.head_dim),
                    qkv=C(a.qkv, f"{name}.attn.qkv"), proj=C(a.proj, f"{name}.attn.proj"), pe=C(a.pe, f"{name}.attn.pe"))
        return dict(type="PSABlock", name=name, attn=attn, add=bool(m.add),
                    ffn=[C(m.ffn[0], f"{name}.ffn.0"), C(m.ffn[1], f"{name}.ffn.1")])
    if tn == "C2PSA":
        return dict(type="C2PSA", name=name, c=int(m.c), cv1=C(m.cv1, f"{name}.cv1"), cv2=C(m.cv2, f"{name}.cv2"),
                    m=[_module_node(x, f"{name}.m.{i}") for i, x in enumerate(m.m)])
    if tn == "SCDown":
        return dict(type="SCDown", name=name, cv1=C(m.cv1, f"{name}.cv1"), cv2=C(m.cv2, f"{name}.cv2"))
    if tn == "MSCAAttention":
        n = dict(type="MSCA", name=name, c=m.conv0.in_channels)
        for k in ("conv0", "dilconv", "conv4", "conv0_1", "conv0_2", "conv1_1", "conv1_2", "conv2_1", "conv2_2",
                  "conv3_1", "conv3_2"):
            n[k] = C(getattr(m, k), f"{name}.{k}")
        for j in range(1, 5):
            n[f"SE{j}"] = C(getattr(m, f"SE{j}").conv[0], f"{name}.SE{j}.conv.0")
        return n
    if tn == "ELA":
        sc = m.spatial_conv
        assert sc.dilation[0] == 2 and sc.padding[0] == sc.kernel_size[0] - 1
        return dict(type="ELA", name=name, c=sc.in_channels, ks=int(sc.kernel_size[0]), groups=int(m.gn.num_groups),
                    gn_eps=float(m.gn.eps))
    if tn == "Upsample":
        if m.mode != "nearest" or float(m.scale_factor) != 2.0:
            raise NotImplementedError("only nearest 2x upsample is on the B200 path")
        return dict(type="Upsample", name=name, scale=2)
    if tn == "Concat":
        assert m.d == 1
        return dict(type="Concat", name=name)
    if tn in ("Detect", "Pose", "Segment"):
        if getattr(m, "end2end", False) or m.reg_max != 16:
            raise NotImplementedError("end2end / non-DFL16 Detect is not on the B200 path")
        cv2 = [[C(s[j], f"{name}.cv2.{i}.{j}") for j in range(3)] for i, s in enumerate(m.cv2)]
        cv3 = []
        for i, s in enumerate(m.cv3):
            if type(s[0]).__name__ != "Sequential":
                raise NotImplementedError("legacy Detect cls branch (head.py:46-49) is not on the B200 path")
            cv3.append([[C(s[0][0], f"{name}.cv3.{i}.0.0"), C(s[0][1], f"{name}.cv3.{i}.0.1")],
                        [C(s[1][0], f"{name}.cv3.{i}.1.0"), C(s[1][1], f"{name}.cv3.{i}.1.1")],
                        C(s[2], f"{name}.cv3.{i}.2")])
        node = dict(type="Detect", name=name, nc=int(m.nc), reg_max=16, no=int(m.no), nl=int(m.nl),
                    ch=[s[0].conv.in_channels for s in m.cv2], cv2=cv2, cv3=cv3, stride=[float(v) for v in m.stride])
        if tn == "Pose":  # head.py:230-250: Detect + one keypoint branch per level
            node.update(head="pose", kpt_shape=(int(m.kpt_shape[0]), int(m.kpt_shape[1])), ne=int(m.nk),
                        cv4=[[C(s[j], f"{name}.cv4.{i}.{j}") for j in range(3)] for i, s in enumerate(m.cv4)])
        if tn == "Segment":  # head.py:175-197: Detect + Proto on P3 + one mask-coefficient branch per level
            p, up = m.proto, m.proto.upsample
            if (type(up).__name__ != "ConvTranspose2d" or tuple(up.kernel_size) != (2, 2) or tuple(up.stride) != (2, 2)
                    or tuple(up.padding) != (0, 0) or up.groups != 1 or up.bias is None):
                raise NotImplementedError(f"{name}.proto.upsample: only ConvTranspose2d(c, c, 2, 2, 0, bias=True) is on the B200 path")
            node.update(head="segment", nm=int(m.nm), npr=int(m.npr), ne=int(m.nm),
                        proto=dict(type="Proto", name=f"{name}.proto", cv1=C(p.cv1, f"{name}.proto.cv1"),
                                   upsample=dict(type="ConvT", name=f"{name}.proto.upsample", c1=up.in_channels,
                                                 c2=up.out_channels, k=(2, 2), s=(2, 2), w=f"{name}.proto.upsample.weight",
                                                 b=f"{name}.proto.upsample.bias"),
                                   cv2=C(p.cv2, f"{name}.proto.cv2"), cv3=C(p.cv3, f"{name}.proto.cv3")),
                        cv4=[[C(s[j], f"{name}.cv4.{i}.{j}") for j in range(3)] for i, s in enumerate(m.cv4)])
        return node
    raise NotImplementedError(f"module {tn} ({name}) is not on the B200 hot path (no CPU fallback by design)")


def spec_from_model(det_model) -> Node:
    """Walk a live reference ``DetectionModel`` exactly as ``_predict_once`` does (tasks.py:138-165):
    layer list ``model.model``, routing by ``m.f``, keep-list ``model.save``.  Follows ``m.f`` literally
    (the BS-YOLO yaml's absolute indices were never renumbered)."""
    layers = []
    for i, m in enumerate(det_model.model):
        f = m.f if isinstance(m.f, int) else list(m.f)
        layers.append(dict(i=i, f=f, module=_module_node(m, f"model.{i}")))
    nc = layers[-1]["module"]["nc"]
    names = getattr(det_model, "names", None) or {k: f"class{k}" for k in range(nc)}
    return dict(model=type(det_model).__name__, nc=nc, layers=layers, save=sorted(det_model.save), names=dict(names))


def ir_from_model(det_model) -> Node:
    return build_ir(spec_from_model(det_model), det_model.state_dict())


def strip_tensors(node):
    """Spec view of an IR/spec tree (tensors and tuple/list distinctions removed) for structural comparison."""
    if isinstance(node, dict):
        return {k: strip_tensors(v) for k, v in node.items() if not isinstance(v, torch.Tensor) and k != "names" and k != "model"}
    if isinstance(node, (list, tuple)):
        return [strip_tensors(v) for v in node]
    return node

"""Engine: one lowered network bound to (device, batch, H, W).  PyTorch is plumbing only (device memory,
streams); every kernel on the path is ours, launched through the C ABI (include/b2y.h).

Replaces, for one predictor, AutoBackend's pt branch (ultralytics/nn/autobackend.py:144-155, 523-524) plus
Detect._inference; the reference-facing wrapper lives in predictor.py."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _cabi as A
from .plan import Plan


def letterbox_geometry(shape, new_shape, auto: bool, stride: int = 32):
    """LetterBox geometry (ultralytics/data/augment.py:1560-1588; scaleup=True, center=True, scaleFill=False).
    Returns (H', W', top, left, uh, uw): canvas size, placement and size of the resized image."""
    h, w = shape
    r = min(new_shape[0] / h, new_shape[1] / w)
    uw, uh = int(round(w * r)), int(round(h * r))
    dw, dh = new_shape[1] - uw, new_shape[0] - uh
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return uh + top + bottom, uw + left + right, top, left, uh, uw


def space_to_depth(x: torch.Tensor) -> torch.Tensor:
    """(B,3,H,W) -> (B,H/2,W/2,16): channel ((y&1)*2 + (x&1))*4 + {R,G,B,0}, the network-input layout of b2y_preprocess."""
    B, _, H, W = x.shape
    x4 = torch.cat([x, torch.zeros_like(x[:, :1])], 1)
    return x4.view(B, 4, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, H // 2, W // 2, 16)


class Engine:
    def __init__(self, ir, batch: int, H: int, W: int, device: int | torch.device = 0, conv_impl: int = A.IMPL_AUTO,
                 use_graph: bool = True, plan: Plan | None = None, autotune: bool = True, fuse_decode: bool = True,
                 mid_stride: int = 0):
        if not torch.cuda.is_available():
            raise A.B2YError("bs_yolo_b200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = A.lib()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.B, self.H, self.W = batch, H, W
        self.plan = plan if plan is not None else Plan(ir, H, W, conv_impl, fuse_decode=fuse_decode)
        self.nc, self.A = self.plan.nc, self.plan.A
        ops, bufs, ws_bytes = self.plan.c_arrays(batch)
        self._ops, self._bufs = ops, bufs  # keep alive
        self.offsets = [b.offset for b in bufs]
        with torch.cuda.device(self.device):
            self.ws = torch.zeros(ws_bytes + 256, dtype=torch.uint8, device=self.device)
            blob = self.plan.blob.tobytes() or bytes(256)  # (plans of weight-less ops only: unit tests)
            self.weights = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(self.device)
            handle = C.c_void_p()
            A.check(self.lib.b2y_plan_create(ops, len(ops), bufs, len(bufs), self.ws.data_ptr(), ws_bytes,
                                             self.weights.data_ptr(), self.weights.numel(), batch,
                                             self.device.index or 0, C.byref(handle)))
        self.handle = handle
        self.use_graph = use_graph
        self._graph_ready = False
        self._tuned = not autotune
        self.max_det_alloc = 0
        self._nms_ws = None
        # mid-point marker (b2y_plan_set_mid_op): in front of the first op that works on a map of stride >= mid_stride --
        # from there on the layers are thin and latency-bound, and DetectPipeline parks its side-stream kernels behind it
        self.mid_op = -1
        if mid_stride > 0:
            for i, o in enumerate(self.plan.ops):
                if getattr(o.out, "H", None) is not None and o.out.H * mid_stride <= H:
                    self.mid_op = i
                    break
            A.check(self.lib.b2y_plan_set_mid_op(self.handle, self.mid_op))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b2y_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ memory views (torch plumbing)
    def buffer(self, idx: int) -> torch.Tensor:
        b = self.plan.bufs[idx]
        dt, esz = (torch.float32, 4) if b.dtype == A.F32 else (torch.bfloat16, 2)
        n = self.B * b.elems() * esz
        return self.ws[self.offsets[idx]: self.offsets[idx] + n].view(dt).view(self.B, b.H, b.W, b.C)

    def view_tensor(self, v) -> torch.Tensor:
        return self.buffer(v.buf)[..., v.coff: v.coff + v.c]

    @property
    def input_nhwc(self) -> torch.Tensor:
        return self.buffer(self.plan.input_buf)

    @property
    def y(self) -> torch.Tensor:
        """(B, 4+nc, A) fp32 -- the tensor Detect._inference returns (head.py:131)."""
        return self.buffer(self.plan.y_buf).view(self.B, 4 + self.nc, self.A)

    def raw(self, level: int) -> torch.Tensor:
        """(B, H_l, W_l, 64+nc) fp32 NHWC head map (permute(0,3,1,2) gives the reference's x[i], head.py:70).
        Only materialised by plans built with fuse_decode=False."""
        if not self.plan.head_bufs:
            raise A.B2YError("raw head maps are not materialised when Detect decode is fused (use fuse_decode=False)")
        return self.buffer(self.plan.head_bufs[level])

    def extra(self, level: int) -> torch.Tensor:
        """(B, H_l, W_l, pitch) fp32 NHWC raw map of the head's extra branch (Pose.cv4, head.py:245); channels [0, plan.ne)
        are the reference's ``cv4[level](x[level])`` (permute(0,3,1,2)), the rest is pitch padding."""
        if not self.plan.extra_bufs:
            raise A.B2YError("this plan has no extra head branch (not a Pose / Segment model)")
        return self.buffer(self.plan.extra_bufs[level])

    @property
    def protos(self) -> torch.Tensor:
        """(B, H/4, W/4, nm) bf16 NHWC mask prototypes of a Segment head (Proto, block.py:95-97; permute(0,3,1,2) gives the
        reference's ``p``, head.py:190)."""
        if self.plan.proto_buf < 0:
            raise A.B2YError("this plan has no mask prototypes (not a Segment model)")
        return self.buffer(self.plan.proto_buf)

    def mask_coeffs(self, kept: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
        """(B, max_det, nm) fp32 mask coefficients of the rows NMS kept -- the `mask` columns non_max_suppression carries
        along in the reference (ops.py:245-246, 268-273)."""
        from .ops import gather_extras
        if self.plan.extra_kind != "segment":
            raise A.B2YError("mask_coeffs: not a Segment model")
        return gather_extras([self.extra(l) for l in range(len(self.plan.extra_bufs))], self.plan.extra_strides, self.plan.ne,
                             kept, counts, mode=0)

    def keypoints(self, kept: torch.Tensor, counts: torch.Tensor, geom: torch.Tensor | None = None,
                  round_dets: torch.Tensor | None = None) -> torch.Tensor:
        """Decoded keypoints (B, max_det, nkpt, ndim) of the rows NMS kept (``nms(return_anchors=True)``): Pose.kpts_decode +
        the gather non_max_suppression does + (with ``geom`` = ops.coords_geometry) ops.scale_coords, one launch."""
        from .ops import gather_extras
        if self.plan.extra_kind != "pose":
            raise A.B2YError("keypoints: not a Pose model")
        out = gather_extras([self.extra(l) for l in range(len(self.plan.extra_bufs))], self.plan.extra_strides, self.plan.ne,
                            kept, counts, mode=1, ndim=self.plan.kpt_shape[1], geom=geom, round_dets=round_dets)
        return out.view(out.shape[0], out.shape[1], *self.plan.kpt_shape)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ stages
    def set_input(self, x: torch.Tensor):
        """Tensor source (predictor.py:123-134 skips LetterBox for tensors): (B,3,H,W) float RGB in [0,1]."""
        assert tuple(x.shape) == (self.B, 3, self.H, self.W), (x.shape, (self.B, 3, self.H, self.W))
        self.input_nhwc.copy_(space_to_depth(x.to(self.device)).to(torch.bfloat16))

    def preprocess(self, src: torch.Tensor, auto: bool = False):
        """uint8 (B,h,w,3) BGR device tensor -> letterboxed, normalised NHWC bf16 network input (one kernel)."""
        assert src.dtype == torch.uint8 and src.is_cuda and src.dim() == 4 and src.shape[0] == self.B and src.shape[3] == 3
        h, w = int(src.shape[1]), int(src.shape[2])
        Hc, Wc, top, left, uh, uw = letterbox_geometry((h, w), (self.H, self.W), auto)
        if (Hc, Wc) != (self.H, self.W):
            raise ValueError(f"letterboxed size {Hc}x{Wc} != engine size {self.H}x{self.W}")
        src = src.contiguous()
        A.check(self.lib.b2y_preprocess(src.data_ptr(), self.B, h, w, self.input_nhwc.data_ptr(), self.H, self.W, top, left,
                                        uh, uw, 0, self._stream()))

    def forward(self):
        """Backbone + neck + Detect + decode for the batch already in the input buffer."""
        st = self._stream()
        if not self._tuned:  # once: pick the fastest tile configuration per tcgen05 conv, measured on this device
            A.check(self.lib.b2y_plan_autotune(self.handle, st))
            self._tuned = True
        if self.use_graph:
            if not self._graph_ready:
                A.check(self.lib.b2y_plan_run(self.handle, 0, -1, st))  # warm-up: function attributes, lazy loading
                torch.cuda.current_stream(self.device).synchronize()
                A.check(self.lib.b2y_plan_graph_build(self.handle, st))
                self._graph_ready = True
            A.check(self.lib.b2y_plan_graph_launch(self.handle, st))
        else:
            A.check(self.lib.b2y_plan_run(self.handle, 0, -1, st))
        return self.y if self.plan.y_buf >= 0 else None

    def wait_mid(self, stream: torch.cuda.Stream):
        """Make `stream` wait until the most recent forward() has passed its mid-point marker (no-op without marker)."""
        A.check(self.lib.b2y_plan_wait_mid(self.handle, stream.cuda_stream))

    def run_ops(self, first: int, last: int):
        A.check(self.lib.b2y_plan_run(self.handle, first, last, self._stream()))

    def op_impl(self, i: int) -> int:
        return self.lib.b2y_plan_op_impl(self.handle, i)

    @property
    def n_launches(self) -> int:
        return self.lib.b2y_plan_num_launches(self.handle)

    def nms(self, conf=0.25, iou=0.7, max_det=300, classes=None, agnostic=False, multi_label=False, max_nms=30000,
            max_wh=7680.0, pred: torch.Tensor | None = None, return_anchors=False):
        from .ops import nms_raw
        return nms_raw(self.y if pred is None else pred, conf, iou, max_det, classes, agnostic, multi_label, max_nms,
                       max_wh, return_anchors)

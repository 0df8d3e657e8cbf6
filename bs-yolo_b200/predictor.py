"""Drop-in predictor: the reference's own plug point for the detect hot path.

    from bs_yolo_b200.predictor import B200DetectionPredictor
    results = YOLO("best.pt").predict(frames, predictor=B200DetectionPredictor)     # engine/model.py:549-551

Weight loading is untouched: ``YOLO(path)`` unpickles the reference's ``DetectionModel`` exactly as before
(ultralytics/nn/tasks.py attempt_load_one_weight); ``setup_model`` then lowers that live module tree (ir.ir_from_model,
BatchNorm folded here because the fork's class shadowing leaves most blocks unfused) and every
stage of ``BasePredictor.stream_inference`` (engine/predictor.py:217-296) runs on the B200 path:

    preprocess   (predictor.py:116-134 + LetterBox augment.py:1535-1601)  -> b2y_preprocess, one kernel per shape group
    inference    (AutoBackend.forward -> DetectionModel._predict_once)    -> CUDA-graphed plan (b2y_plan_graph_launch)
    postprocess  (models/yolo/detect/predict.py:23-41)                    -> b2y_nms + b2y_scale_boxes

``stream_inference`` itself is inherited, so callbacks, ``Results``, ``Results.speed`` and the per-predictor lock keep
working.  There is no CPU fallback: a non-CUDA device, ``half=True``, ``augment``, ``visualize`` and ``embed`` raise.

The reference package is imported lazily (it is not a dependency of the kernels or of bench.py).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi as A
from . import ir as _ir
from .engine import Engine, letterbox_geometry
from .ops import nms_raw, scale_boxes_batched

_CLS = None


class PreprocessedBatch:
    """What ``preprocess`` hands to ``inference``/``postprocess``: the batch already sits in the engine's input buffer;
    only its (B, 3, H, W) shape is needed downstream (predict.py:39 ``img.shape[2:]``, predictor.py:329)."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.shape = torch.Size((engine.B, 3, engine.H, engine.W))
        self.device = engine.device

    def __len__(self):
        return self.shape[0]


class B200Backend:
    """Stands where ``AutoBackend`` stands (engine/predictor.py:308-321): the attributes the predictor reads
    (pt, triton, fp16, stride, names, device, imx, dynamic) and ``warmup`` / ``eval`` / ``__call__``."""

    pt, triton, fp16, imx, dynamic, jit, nn_module = True, False, False, False, False, False, True

    def __init__(self, det_model, device: torch.device, use_graph: bool = True):
        if device.type != "cuda":
            raise A.B2YError(f"B200DetectionPredictor needs a CUDA device, got '{device}' (no CPU fallback on this path)")
        self.model = det_model  # kept for attribute access by callers (e.g. tracker callbacks read predictor.model.model)
        self.device = device
        self.ir = _ir.ir_from_model(det_model)
        self.names = dict(getattr(det_model, "names", None) or self.ir["names"])
        self.stride = max(int(max(float(s) for s in det_model.stride)), 32)
        self.nc = self.ir["nc"]
        self.use_graph = use_graph
        self._engines: dict[tuple[int, int, int], Engine] = {}

    def engine(self, B: int, H: int, W: int) -> Engine:
        key = (B, H, W)
        e = self._engines.get(key)
        if e is None:
            if len(self._engines) >= 4:  # a stream of changing shapes must not accumulate workspaces
                self._engines.pop(next(iter(self._engines)))
            e = Engine(self.ir, B, H, W, device=self.device, use_graph=self.use_graph)
            self._engines[key] = e
        return e

    def eval(self):
        return self

    def warmup(self, imgsz=(1, 3, 640, 640)):
        """AutoBackend.warmup (autobackend.py:719-732) runs a dummy forward; here the (B,H,W)-specific plan is built,
        autotuned and graph-captured on first use instead, so there is nothing shape-independent to warm."""

    def __call__(self, im, augment=False, visualize=False, embed=None):
        if augment or visualize or embed:
            raise NotImplementedError("augment / visualize / embed are not on the B200 detect path")
        if isinstance(im, PreprocessedBatch):
            eng = im.engine
        else:  # a (B,3,H,W) float tensor handed to the backend directly
            eng = self.engine(int(im.shape[0]), int(im.shape[2]), int(im.shape[3]))
            eng.set_input(im.float())
        return eng.forward()  # (B, 4+nc, A) fp32 == Detect._inference's y (head.py:131)

    forward = __call__


def _build_class():
    from ultralytics.models.yolo.detect.predict import DetectionPredictor
    from ultralytics.engine.results import Results
    from ultralytics.utils.torch_utils import select_device

    class B200DetectionPredictor(DetectionPredictor):
        """DetectionPredictor whose preprocess / inference / postprocess run as hand-written sm_100a kernels."""

        def setup_model(self, model, verbose=True):
            if self.args.half:
                raise NotImplementedError("half=True (fp16) is a reference-only dtype; the B200 path computes in bf16/fp32")
            model = model if model is not None else self.args.model
            if isinstance(model, (str, bytes)) or hasattr(model, "__fspath__"):
                from ultralytics.nn.tasks import attempt_load_weights  # the reference's own loader, unchanged
                model = attempt_load_weights(str(model), device="cpu", inplace=True, fuse=False)
            device = select_device(self.args.device, verbose=verbose)
            self.model = B200Backend(model, torch.device(device))
            self.device = self.model.device
            self.args.half = False

        def preprocess(self, im):
            dev = self.device
            if isinstance(im, torch.Tensor):  # tensor sources skip LetterBox (predictor.py:123-134)
                eng = self.model.engine(int(im.shape[0]), int(im.shape[2]), int(im.shape[3]))
                eng.set_input(im.to(dev).float())
                return PreprocessedBatch(eng)
            shapes = [x.shape for x in im]
            same = len(set(shapes)) == 1
            new_shape = tuple(self.imgsz) if not isinstance(self.imgsz, int) else (self.imgsz, self.imgsz)
            geo = [letterbox_geometry(s[:2], new_shape, auto=same and self.model.pt, stride=self.model.stride) for s in shapes]
            H, W = geo[0][0], geo[0][1]
            if any((g[0], g[1]) != (H, W) for g in geo):
                raise ValueError("letterboxed images of one batch must share one canvas size")
            eng = self.model.engine(len(im), H, W)
            lib, st = A.lib(), torch.cuda.current_stream(dev).cuda_stream
            dst = eng.input_nhwc
            if same:  # one H2D copy + one kernel for the whole batch
                src = torch.from_numpy(np.ascontiguousarray(np.stack(im))).to(dev, non_blocking=True)
                _, _, top, left, uh, uw = geo[0]
                A.check(lib.b2y_preprocess(src.data_ptr(), len(im), shapes[0][0], shapes[0][1], dst.data_ptr(), H, W, top, left,
                                           uh, uw, 0, st))
                self._keep = src
            else:  # ragged batch: LetterBox(auto=False) per image into that image's slot of the input buffer
                keep = []
                for i, x in enumerate(im):
                    src = torch.from_numpy(np.ascontiguousarray(x)).to(dev, non_blocking=True)
                    keep.append(src)
                    _, _, top, left, uh, uw = geo[i]
                    A.check(lib.b2y_preprocess(src.data_ptr(), 1, x.shape[0], x.shape[1], dst[i].data_ptr(), H, W, top, left,
                                               uh, uw, 0, st))
                self._keep = keep
            return PreprocessedBatch(eng)

        def inference(self, im, *args, **kwargs):
            if self.args.augment or self.args.visualize or self.args.embed:
                raise NotImplementedError("augment / visualize / embed are not on the B200 detect path")
            return self.model(im)

        def postprocess(self, preds, img, orig_imgs):
            if isinstance(preds, (list, tuple)):
                preds = preds[0]
            a = self.args
            dets, counts = nms_raw(preds, a.conf, a.iou, a.max_det, a.classes, bool(a.agnostic_nms))
            if not isinstance(orig_imgs, list):
                from ultralytics.utils import ops as ref_ops
                orig_imgs = ref_ops.convert_torch2numpy_batch(orig_imgs)
            scale_boxes_batched(dets, counts, tuple(img.shape[2:]), [o.shape[:2] for o in orig_imgs])
            n = counts.tolist()  # the one device->host sync of the batch
            return [Results(o, path=p, names=self.model.names, boxes=dets[i, : n[i]])
                    for i, (o, p) in enumerate(zip(orig_imgs, self.batch[0]))]

    return B200DetectionPredictor


def __getattr__(name):  # PEP 562: resolve the class on first use so importing this module never imports ultralytics
    global _CLS
    if name == "B200DetectionPredictor":
        if _CLS is None:
            _CLS = _build_class()
        return _CLS
    raise AttributeError(name)

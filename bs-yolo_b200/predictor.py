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

A list of same-shape uint8 frames (``predict(list_of_ndarrays)``, batched file/video sources) takes the pipelined route:
``preprocess`` only records the frames, ``inference`` streams them in chunks of ``B2Y_PREDICT_BATCH`` (64) through
``DetectPipeline.predict_stream`` (pinned staging ring filled by a thread pool -> H2D -> preprocess -> network -> NMS ->
scale_boxes, neighbouring chunks overlapped; boxes stay on the GPU, only the counts cross PCIe) and builds the
``Results`` of chunk i under the GPU work of chunk i+1.  Ragged lists and tensor sources take the single-engine route.

``stream_inference`` itself is inherited, so callbacks, ``Results``, ``Results.speed`` and the per-predictor lock keep
working.  There is no CPU fallback: a non-CUDA device, ``augment``, ``visualize`` and ``embed`` raise; ``half=True`` is accepted
(the path's only 16-bit mode is bf16 operands with fp32 accumulation, and it is always on).

The reference package is imported lazily (it is not a dependency of the kernels or of bench.py).
"""
from __future__ import annotations

import ctypes
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

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


class HostFrames:
    """A list of same-shape uint8 BGR frames still in host memory (the common ``predict(list_of_ndarrays)`` /
    batched-file-source case): ``inference`` streams them through a DetectPipeline in chunks -- pinned staging filled
    by a thread pool, host->device copies, preprocess, network and NMS of neighbouring chunks overlapped."""

    def __init__(self, frames, src_hw, canvas_hw, device):
        self.frames = frames
        self.pinned = getattr(frames, "pinned", None)  # ingest.PinnedFrames: the pixels already sit in a pinned block
        self.src_hw, self.canvas_hw = src_hw, canvas_hw
        self.shape = torch.Size((len(frames), 3, *canvas_hw))
        self.device = device

    def __len__(self):
        return self.shape[0]


class _Staging:
    """Ring of pinned (B,h,w,3) uint8 buffers filled by a worker thread (+ a copy pool: numpy releases the GIL while it
    copies), so that staging chunk i+1 overlaps the GPU work of chunk i and the ``Results`` construction of chunk i-1.
    A buffer is refilled only after the host->device copy that read it has completed (CUDA event from `uploaded`)."""

    DEPTH = 4

    def __init__(self, B, h, w, depth: int | None = None):
        self.B, self.h, self.w = B, h, w
        if depth is not None:
            self.DEPTH = depth  # the single-engine route stages one batch at a time
        self.bufs = [torch.empty(B, h, w, 3, dtype=torch.uint8).pin_memory() for _ in range(self.DEPTH)]
        self.views = [b.numpy() for b in self.bufs]
        self.free = [threading.Event() for _ in range(self.DEPTH)]
        self.copied = [None] * self.DEPTH
        for f in self.free:
            f.set()
        self.pool = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1)))

    def uploaded(self, frames, event):  # called by DetectPipeline.predict_stream once the H2D copy has been enqueued
        for r, b in enumerate(self.bufs):
            if frames is b:
                self.copied[r] = event
                self.free[r].set()
                return

    def chunks(self, frames):
        """Generator of pinned chunk tensors over `frames` (the last chunk is padded with stale rows)."""
        n, B = len(frames), self.B
        q: queue.Queue = queue.Queue(maxsize=2)

        nbytes = self.h * self.w * 3
        workers = self.pool._max_workers

        def fill_range(args):  # one task per worker: a run of frames, copied without the GIL (ctypes.memmove); frames that
            r, lo, a, b_ = args  # are adjacent in memory (views of one array, a decoder's ring) go in ONE call, so the worker
            base = self.bufs[r].data_ptr()  # does not have to win the GIL back from the main thread between frames
            view = self.views[r]
            j = a
            while j < b_:
                f = frames[j]
                if not f.flags.c_contiguous:
                    np.copyto(view[j - lo], f)
                    j += 1
                    continue
                src, k = f.ctypes.data, j + 1
                while k < b_ and frames[k].flags.c_contiguous and frames[k].ctypes.data == src + (k - j) * nbytes:
                    k += 1
                ctypes.memmove(base + (j - lo) * nbytes, src, (k - j) * nbytes)
                j = k

        def produce():
            try:
                for c, lo in enumerate(range(0, n, B)):
                    r = c % self.DEPTH
                    self.free[r].wait()
                    if self.copied[r] is not None:
                        self.copied[r].synchronize()
                        self.copied[r] = None
                    self.free[r].clear()
                    hi = min(lo + B, n)
                    per = -(-(hi - lo) // workers)
                    list(self.pool.map(fill_range, [(r, lo, a, min(a + per, hi)) for a in range(lo, hi, per)]))
                    q.put(self.bufs[r])
                q.put(None)
            except BaseException as e:  # noqa: BLE001 -- surfaced in the consumer thread
                q.put(e)

        threading.Thread(target=produce, daemon=True).start()
        while True:
            item = q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item


class B200Backend:
    """Stands where ``AutoBackend`` stands (engine/predictor.py:308-321): the attributes the predictor reads
    (pt, triton, fp16, stride, names, device, imx, dynamic) and ``warmup`` / ``eval`` / ``__call__``."""

    pt, triton, fp16, imx, dynamic, jit, nn_module = True, False, False, False, False, False, True
    max_engines = 4  # plans kept per backend (one per (batch, H, W)); the validator raises it for rect batches

    def __init__(self, det_model, device: torch.device, use_graph: bool = True, head: str | None = None, who: str = "B200Backend"):
        if device.type != "cuda":
            raise A.B2YError(f"B200DetectionPredictor needs a CUDA device, got '{device}' (no CPU fallback on this path)")
        self.model = det_model  # kept for attribute access by callers (e.g. tracker callbacks read predictor.model.model)
        self.device = device
        self.ir = _ir.ir_from_model(det_model)
        _ir.require_head(self.ir, head, who)  # e.g. a pose checkpoint behind B200DetectionPredictor: raise, do not drop keypoints
        self.names = dict(getattr(det_model, "names", None) or self.ir["names"])
        self.stride = max(int(max(float(s) for s in det_model.stride)), 32)
        self.nc = self.ir["nc"]
        head = _ir.detect_node(self.ir)
        if head.get("head") == "pose":  # AutoBackend forwards the checkpoint's kpt_shape (autobackend.py:149-150)
            self.kpt_shape = list(head["kpt_shape"])
        self.use_graph = use_graph
        self._engines: dict[tuple[int, int, int], Engine] = {}
        self._pipes: dict[tuple, tuple] = {}
        self._stagings: dict[tuple, _Staging] = {}

    def engine(self, B: int, H: int, W: int) -> Engine:
        key = (B, H, W)
        e = self._engines.get(key)
        if e is None:
            if len(self._engines) >= self.max_engines:  # a stream of changing shapes must not accumulate workspaces
                self._engines.pop(next(iter(self._engines)))
            with torch.cuda.device(self.device):
                e = Engine(self.ir, B, H, W, device=self.device, use_graph=self.use_graph)
            self._engines[key] = e
        return e

    def pipeline(self, B: int, src_hw, imgsz):
        """(DetectPipeline, staging ring) for chunks of B same-shape frames; cached per (B, source shape, imgsz)."""
        from .pipeline import DetectPipeline

        key = (B, tuple(src_hw), tuple(imgsz))
        hit = self._pipes.get(key)
        if hit is None:
            if len(self._pipes) >= 3:
                self._pipes.pop(next(iter(self._pipes)))
            with torch.cuda.device(self.device):
                pipe = DetectPipeline(self.ir, B, tuple(imgsz), tuple(src_hw), device=self.device.index or 0, auto=True,
                                      use_graph=self.use_graph)
            hit = (pipe, _Staging(B, *src_hw))
            self._pipes[key] = hit
        return hit

    def staging(self, B: int, src_hw) -> "_Staging":
        """Pinned staging ring for same-shape batches on the single-engine route (Pose / Segment predictors), cached per
        (B, source shape): allocating pinned memory per batch (``.pin_memory()``) costs more than the whole GPU step."""
        key = (B, tuple(src_hw))
        hit = self._stagings.get(key)
        if hit is None:
            if len(self._stagings) >= 2:  # a stream of changing shapes must not accumulate pinned memory
                self._stagings.pop(next(iter(self._stagings)))
            hit = _Staging(B, *src_hw, depth=1)
            self._stagings[key] = hit
        return hit

    def eval(self):
        return self

    def warmup(self, imgsz=(1, 3, 640, 640)):
        """AutoBackend.warmup (autobackend.py:719-732) runs a dummy forward; here the (B,H,W)-specific plan is built,
        autotuned and graph-captured on first use instead, so there is nothing shape-independent to warm."""

    def __call__(self, im, augment=False, visualize=False, embed=None):
        """AutoBackend.forward for a pt model returns the model's output: Detect's eval-mode ``(y, x)`` (head.py:74).
        Here: ``(y, None)`` -- y (B, 4+nc, A) fp32; the raw per-level maps never exist (decode is fused into the head
        convs' epilogues).  A (B,3,H,W) float tensor handed to the backend directly is space-to-depth'd into the input."""
        if augment or visualize or embed:
            raise NotImplementedError("augment / visualize / embed are not on the B200 detect path")
        if isinstance(im, PreprocessedBatch):
            eng = im.engine
        else:
            eng = self.engine(int(im.shape[0]), int(im.shape[2]), int(im.shape[3]))
            eng.set_input(im.float())
        with torch.cuda.device(self.device):
            return eng.forward(), None

    forward = __call__


def _build_class():
    from ultralytics.models.yolo.detect.predict import DetectionPredictor
    from ultralytics.engine.results import Results
    from ultralytics.utils.torch_utils import select_device

    class B200DetectionPredictor(DetectionPredictor):
        """DetectionPredictor whose preprocess / inference / postprocess run as hand-written sm_100a kernels."""

        max_chunk = int(os.environ.get("B2Y_PREDICT_BATCH", "64"))  # frames per pipelined chunk of a same-shape list
        pipelined = True  # same-shape uint8 lists stream through DetectPipeline (boxes only); False: single-engine route
        head_kind = "detect"  # the head this class post-processes (ir.require_head)

        def setup_model(self, model, verbose=True):
            if self.args.half:
                # engine/predictor.py:131 + autobackend.py:153: the reference casts model and input to fp16.  The B200 path has one
                # 16-bit mode and always runs in it: bf16 operands, fp32 accumulation and fp32 decode / NMS (fp16's range would
                # need loss scaling in the DFL exponentials; bf16 does not).  half=True therefore selects nothing new -- it is
                # accepted so that callers' scripts keep working, and the flag is cleared (model.fp16 stays False: no .half()
                # of the input anywhere downstream).  Parity gate: the same 1e-2 / IoU 0.99 gate as half=False.
                from ultralytics.utils import LOGGER
                LOGGER.info("bs_yolo_b200: half=True -> the B200 path already computes in 16 bit (bf16 operands, fp32 accumulate)")
            model = model if model is not None else self.args.model
            if isinstance(model, (str, bytes)) or hasattr(model, "__fspath__"):
                from ultralytics.nn.tasks import attempt_load_weights  # the reference's own loader, unchanged
                model = attempt_load_weights(str(model), device="cpu", inplace=True, fuse=False)
            device = select_device(self.args.device, verbose=verbose)
            self.model = B200Backend(model, torch.device(device), head=self.head_kind, who=type(self).__name__)
            self.device = self.model.device
            self.args.half = False

        def setup_source(self, source):
            """engine/predictor.py:298-323, then: a source of image FILES gets the prefetching loader (ingest.py) in place of
            LoadImagesAndVideos -- parallel decode ahead of the GPU, batches handed over in pinned memory."""
            super().setup_source(source)
            from ultralytics.data.loaders import LoadImagesAndVideos

            from .ingest import PrefetchImageLoader

            ds = self.dataset
            if (type(ds) is LoadImagesAndVideos and ds.bs > 1 and not any(ds.video_flag) and os.environ.get("B2Y_INGEST", "1") != "0"
                    and not any(f.rsplit(".", 1)[-1].lower() == "heic" for f in ds.files)):
                self.dataset = PrefetchImageLoader(ds)

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
            hot = same and self.model.stride == 32 and all(x.dtype == np.uint8 and x.ndim == 3 and x.shape[2] == 3 for x in im)
            if hot and self.pipelined:
                # the hot case: nothing is copied here; `inference` streams the frames through the pipelined path
                return HostFrames(im, shapes[0][:2], (H, W), dev)
            eng = self.model.engine(len(im), H, W)
            lib, st = A.lib(), torch.cuda.current_stream(dev).cuda_stream
            dst = eng.input_nhwc
            if hot:  # same-shape batch on the single-engine route (heads whose extras the chunked pipeline does not carry)
                stg = self.model.staging(len(im), shapes[0][:2])
                pinned = next(iter(stg.chunks(im)))  # cached pinned ring, filled by the copy pool (GIL-free memmoves)
                src = pinned.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                stg.uploaded(pinned, ev)  # the ring slot may be refilled once this copy has completed
                _, _, top, left, uh, uw = geo[0]
                A.check(lib.b2y_preprocess(src.data_ptr(), len(im), shapes[0][0], shapes[0][1], dst.data_ptr(), H, W, top, left,
                                           uh, uw, 0, st))
                self._keep = [src]
                return PreprocessedBatch(eng)
            keep = []  # ragged batch: LetterBox(auto=False) per image into that image's slot of the input buffer
            for i, x in enumerate(im):
                src = torch.from_numpy(np.ascontiguousarray(x)).pin_memory().to(dev, non_blocking=True)
                keep.append(src)
                _, _, top, left, uh, uw = geo[i]
                A.check(lib.b2y_preprocess(src.data_ptr(), 1, x.shape[0], x.shape[1], dst[i].data_ptr(), H, W, top, left,
                                           uh, uw, 0, st))
            self._keep = keep
            return PreprocessedBatch(eng)

        def inference(self, im, *args, **kwargs):
            if self.args.augment or self.args.visualize or self.args.embed:
                raise NotImplementedError("augment / visualize / embed are not on the B200 detect path")
            if not isinstance(im, HostFrames):
                return self.model(im)
            a = self.args
            n = len(im)
            B = min(n, self.max_chunk)
            new_shape = tuple(self.imgsz) if not isinstance(self.imgsz, int) else (self.imgsz, self.imgsz)
            pipe, staging = self.model.pipeline(B, im.src_hw, new_shape)
            pipe.conf, pipe.iou, pipe.max_det = a.conf, a.iou, a.max_det
            pipe.classes, pipe.agnostic = a.classes, bool(a.agnostic_nms)
            names, paths, origs = self.model.names, self.batch[0], im.frames
            results = []
            with torch.cuda.device(self.device):
                blk = im.pinned
                if blk is not None and all(lo + B <= blk.shape[0] for lo in range(0, n, B)):
                    # the loader already put the batch into pinned memory (ingest.py): upload it from there, chunk by chunk
                    stream = pipe.predict_stream((blk[lo: lo + B] for lo in range(0, n, B)), rescale=True, device_dets=True)
                else:
                    stream = pipe.predict_stream(staging.chunks(im.frames), rescale=True, device_dets=True,
                                                 uploaded=staging.uploaded)
                for c, (dets, counts) in enumerate(stream):
                    lo = c * B
                    cnt = counts.tolist()
                    # Results are built here, chunk by chunk, under the GPU work of the following chunks
                    results.extend(Results(origs[lo + i], path=paths[lo + i], names=names, boxes=dets[i, : cnt[i]])
                                   for i in range(min(B, n - lo)))
            return results

        def postprocess(self, preds, img, orig_imgs):
            if isinstance(img, HostFrames):  # already post-processed chunk by chunk in `inference`
                return preds
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

    class B200PosePredictor(B200DetectionPredictor):
        """PosePredictor (models/yolo/pose/predict.py:8-59) on the B200 path: backbone, neck, the Detect
        branches and the keypoint branch (Pose.cv4, head.py:239-246) run in the same CUDA-graphed plan; NMS is the Detect
        NMS (the reference drags the 51 decoded keypoint rows of every anchor through it, ops.py:245-273); Pose.kpts_decode
        (head.py:252-279), the gather of the kept rows, ops.scale_coords (ops.py:740-772) and the `.round()` of the boxes
        (predict.py:52) are ONE launch over the <= max_det kept rows (b2y_gather_extras).  ``Results.keypoints`` stay on
        the GPU.  Batches take the single-engine route (the chunked pipeline carries boxes only)."""

        pipelined = False
        head_kind = "pose"

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self.args.task = "pose"

        def postprocess(self, preds, img, orig_imgs):
            from .ops import coords_geometry

            if isinstance(preds, (list, tuple)):
                preds = preds[0]
            a = self.args
            dets, counts, kept = nms_raw(preds, a.conf, a.iou, a.max_det, a.classes, bool(a.agnostic_nms), return_anchors=True)
            if not isinstance(orig_imgs, list):
                from ultralytics.utils import ops as ref_ops
                orig_imgs = ref_ops.convert_torch2numpy_batch(orig_imgs)
            shapes = [o.shape[:2] for o in orig_imgs]
            scale_boxes_batched(dets, counts, tuple(img.shape[2:]), shapes)
            kpts = img.engine.keypoints(kept, counts, geom=coords_geometry(tuple(img.shape[2:]), shapes, dets.device),
                                        round_dets=dets)
            n = counts.tolist()  # the one device->host sync of the batch
            nk = kpts.shape[2] * kpts.shape[3]
            return [Results(o, path=p, names=self.model.names, boxes=dets[i, : n[i]],
                            keypoints=kpts[i, : n[i]] if n[i] else kpts[i, :0].reshape(0, nk))  # predict.py:51: (0, nk) when empty
                    for i, (o, p) in enumerate(zip(orig_imgs, self.batch[0]))]

    class B200SegmentationPredictor(B200DetectionPredictor):
        """SegmentationPredictor (models/yolo/segment/predict.py:8-57) on the B200 path: Proto (block.py:80-97;
        its ConvTranspose2d as one tensor-core 1x1 conv + a depth-to-space shuffle) and the mask-coefficient branch
        (Segment.cv4, head.py:186,193) run inside the CUDA-graphed plan; NMS is the Detect NMS; the 32 coefficients of the
        kept rows are gathered afterwards (b2y_gather_extras) and ops.process_mask (ops.py:663-693: coefficients @
        prototypes, crop to the box, bilinear x4, > 0) runs as ONE launch for the whole batch (b2y_process_mask) instead
        of a per-image loop.  ``Results.masks`` stay on the GPU.  retina_masks=True (process_mask_native) is not built."""

        pipelined = False
        head_kind = "segment"

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self.args.task = "segment"

        def postprocess(self, preds, img, orig_imgs):
            from .ops import process_mask_batched

            if self.args.retina_masks:
                raise NotImplementedError("retina_masks=True (ops.process_mask_native) is not on the B200 segment path")
            if isinstance(preds, (list, tuple)):
                preds = preds[0]
            a = self.args
            dets, counts, kept = nms_raw(preds, a.conf, a.iou, a.max_det, a.classes, bool(a.agnostic_nms), return_anchors=True)
            eng = img.engine
            coeffs = eng.mask_coeffs(kept, counts)
            if not isinstance(orig_imgs, list):
                from ultralytics.utils import ops as ref_ops
                orig_imgs = ref_ops.convert_torch2numpy_batch(orig_imgs)
            n = counts.tolist()  # the one device->host sync of the batch
            masks = process_mask_batched(eng.protos, coeffs, dets, n, tuple(img.shape[2:]))  # from the UNSCALED boxes (predict.py:54)
            scale_boxes_batched(dets, counts, tuple(img.shape[2:]), [o.shape[:2] for o in orig_imgs])
            return [Results(o, path=p, names=self.model.names, boxes=dets[i, : n[i]], masks=masks[i])
                    for i, (o, p) in enumerate(zip(orig_imgs, self.batch[0]))]

    return {"B200DetectionPredictor": B200DetectionPredictor, "B200PosePredictor": B200PosePredictor,
            "B200SegmentationPredictor": B200SegmentationPredictor}


def __getattr__(name):  # PEP 562: resolve the classes on first use so importing this module never imports ultralytics
    global _CLS
    if name in ("B200DetectionPredictor", "B200PosePredictor", "B200SegmentationPredictor"):
        if _CLS is None:
            _CLS = _build_class()
        return _CLS[name]
    raise AttributeError(name)

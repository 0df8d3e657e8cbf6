"""DetectPipeline: the public batch API of the B200 path -- uint8 BGR frames in, final detections out.

One pipeline = one GPU = one CUDA-graphed network (Engine) + preprocess + NMS (+ optional scale_boxes), i.e. the
device-side body of BasePredictor.stream_inference's hot loop (ultralytics/engine/predictor.py:241-266):
preprocess -> inference -> postprocess.  `predict_host` is the end-to-end entry: pinned host frames are copied in,
detections are copied back, with the copies double-buffered against compute on a second stream."""
from __future__ import annotations

import torch

from . import _cabi as A
from .engine import Engine, letterbox_geometry
from .ops import nms_raw, scale_boxes_batched


class DetectPipeline:
    def __init__(self, ir, batch: int, imgsz=(640, 640), src_shape=(640, 640), device: int = 0, conf=0.25, iou=0.7,
                 max_det=300, classes=None, agnostic=False, auto=True, use_graph=True, conv_impl=A.IMPL_AUTO):
        imgsz = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        self.src_shape = tuple(src_shape)
        H, W, *_ = letterbox_geometry(self.src_shape, imgsz, auto)
        self.auto = auto
        self.engine = Engine(ir, batch, H, W, device=device, use_graph=use_graph, conv_impl=conv_impl)
        self.device = self.engine.device
        self.B, self.H, self.W = batch, H, W
        self.conf, self.iou, self.max_det, self.classes, self.agnostic = conf, iou, max_det, classes, agnostic
        self._copy_stream = None
        self._staging = None

    @property
    def launches_per_batch(self) -> int:
        return self.engine.n_launches + 1 + 2  # network ops + preprocess + NMS (score, select)

    def __call__(self, src_u8: torch.Tensor, rescale: bool = False, consumed: torch.cuda.Event | None = None):
        """src_u8: (B,h,w,3) uint8 BGR on this pipeline's device. Returns (dets (B,max_det,6), counts (B,)) on device.
        ``consumed`` (optional) is recorded as soon as the frames have been read (after the preprocess kernel)."""
        self.engine.preprocess(src_u8, auto=self.auto)
        if consumed is not None:
            consumed.record(torch.cuda.current_stream(self.device))
        self.engine.forward()
        dets, counts = nms_raw(self.engine.y, self.conf, self.iou, self.max_det, self.classes, self.agnostic)
        if rescale:
            scale_boxes_batched(dets, counts, (self.H, self.W), [self.src_shape] * self.B)
        return dets, counts

    def predict_host(self, frames: torch.Tensor, rescale: bool = False, gather: bool = False):
        """frames: pinned host uint8 (B,h,w,3). Returns host (dets, counts); blocks until they are valid.
        gather=True (multi-GPU): every rank receives the detections of all ranks' shards (shard.gather_detections)."""
        if self._staging is None:
            self._staging = torch.empty(frames.shape, dtype=torch.uint8, device=self.device)
        self._staging.copy_(frames, non_blocking=True)
        dets, counts = self(self._staging, rescale)
        if gather:
            from .shard import gather_detections
            dets, counts = gather_detections(dets, counts)
        if getattr(self, "_out", None) is None or self._out[0].shape != dets.shape:
            self._out = (torch.empty(dets.shape, dtype=torch.float32).pin_memory(),
                         torch.empty(counts.shape, dtype=torch.int32).pin_memory())
        self._out[0].copy_(dets, non_blocking=True)
        self._out[1].copy_(counts, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._out

    def predict_stream(self, batches, rescale: bool = False, gather: bool = False):
        """Streaming end-to-end API (the analogue of ``predict(..., stream=True)``, engine/predictor.py:217-296): an
        iterable of pinned host uint8 batches (B,h,w,3) in, one host ``(dets, counts)`` pair per batch out, in order.
        The host->device copy of batch i+1 runs on a second stream while batch i is computed, and the result of batch i
        is handed out after batch i+1 has been enqueued, so neither the PCIe copy nor the host-side wait idles the GPU.
        The yielded tensors are pinned buffers that stay valid until two more batches have been requested."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        stage, copied, consumed, done, outs = [None, None], [None, None], [None, None], [None, None], [None, None]
        for k in range(2):
            copied[k], consumed[k], done[k] = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

        def upload(frames, k):
            if stage[k] is None or stage[k].shape != frames.shape:
                stage[k] = torch.empty(frames.shape, dtype=torch.uint8, device=dev)
            with torch.cuda.stream(cs):
                cs.wait_event(consumed[k])  # the preprocess kernel that last read this slot has finished
                stage[k].copy_(frames, non_blocking=True)
                copied[k].record(cs)

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        consumed[0].record(main)
        consumed[1].record(main)
        upload(nxt, 0)
        i, pending = 0, None
        while nxt is not None:
            k = i & 1
            nxt = next(it, None)
            if nxt is not None:
                upload(nxt, k ^ 1)  # overlaps the compute of batch i
            main.wait_event(copied[k])
            dets, counts = self(stage[k], rescale, consumed=consumed[k])
            if gather:
                from .shard import gather_detections
                dets, counts = gather_detections(dets, counts)
            if outs[k] is None or outs[k][0].shape != dets.shape:
                outs[k] = (torch.empty(dets.shape, dtype=torch.float32).pin_memory(),
                           torch.empty(counts.shape, dtype=torch.int32).pin_memory())
            outs[k][0].copy_(dets, non_blocking=True)
            outs[k][1].copy_(counts, non_blocking=True)
            done[k].record(main)
            if pending is not None:
                done[pending].synchronize()
                yield outs[pending]
            pending = k
            i += 1
        done[pending].synchronize()
        yield outs[pending]

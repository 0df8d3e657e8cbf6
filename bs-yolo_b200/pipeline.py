"""DetectPipeline: the public batch API of the B200 path -- uint8 BGR frames in, final detections out.

One pipeline = one GPU: preprocess + CUDA-graphed network + NMS (+ optional scale_boxes), i.e. the device-side body of
BasePredictor.stream_inference's hot loop (ultralytics/engine/predictor.py:241-266): preprocess -> inference ->
postprocess.  Batches are independent, so the pipeline is software-pipelined over two slots (two Engines = two activation
workspaces and graphs).  The network's persistent kernels keep every SM busy through the wide, bandwidth-bound first half
of the graph and leave most of them idle in the thin P4/P5 layers, so the side work is parked there: the preprocess of
batch i waits for the mid-point marker of graph i-1, and the NMS (+ gather, + device->host copy) of batch i is enqueued
only once graph i+1 has been launched, behind its marker (`Engine.wait_mid`).  In `predict_stream` the host<->device
copies overlap as well.

    __call__(frames)        one batch, stream-synchronous semantics: the results are ready for the caller's stream
    submit(frames)          one batch, pipelined: returns a Ticket; `ticket.result()` -> (dets, counts, event)
    flush()                 enqueue the post-process of the last submitted batch (nothing left deferred)
    predict_stream(batches) end to end with host buffers: pinned frames in, pinned detections out, everything overlapped
"""
from __future__ import annotations

import os

import torch

from . import _cabi as A
from .engine import Engine, letterbox_geometry
from .ops import nms_raw, scale_boxes_batched, scale_geometry


class Ticket:
    """Handle of one submitted batch.  Its post-process is enqueued by the NEXT submit() (or by flush()/result())."""

    def __init__(self, pipe, slot, rescale, gather, host_out):
        self.pipe, self.slot, self.rescale, self.gather, self.host_out = pipe, slot, rescale, gather, host_out
        self.dets = self.counts = self.event = None

    def result(self):
        """(dets (B,max_det,6), counts (B,), event): the tensors are valid once `event` has completed."""
        if self.event is None:
            self.pipe.flush()
        return self.dets, self.counts, self.event

    def __iter__(self):
        return iter(self.result())


class DetectPipeline:
    def __init__(self, ir, batch: int, imgsz=(640, 640), src_shape=(640, 640), device: int = 0, conf=0.25, iou=0.7,
                 max_det=300, classes=None, agnostic=False, auto=True, use_graph=True, conv_impl=A.IMPL_AUTO, slots: int = 2):
        imgsz = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        slots = int(os.environ.get("B2Y_SLOTS", slots))  # A/B knob: engines (= network graphs that may be in flight at once)
        self.src_shape = tuple(src_shape)
        H, W, *_ = letterbox_geometry(self.src_shape, imgsz, auto)
        self.auto = auto
        self.defer = slots >= 2 and os.environ.get("B2Y_DEFER_POST", "1") != "0"
        mid = int(os.environ.get("B2Y_MID_STRIDE", "16")) if self.defer else 0
        self.engines = [Engine(ir, batch, H, W, device=device, use_graph=use_graph, conv_impl=conv_impl, mid_stride=mid)
                        for _ in range(max(1, slots))]
        self.engine = self.engines[0]
        self.device = self.engine.device
        self.B, self.H, self.W = batch, H, W
        self.conf, self.iou, self.max_det, self.classes, self.agnostic = conf, iou, max_det, classes, agnostic
        self._slot = 0
        self._pending = None  # Ticket of the batch whose post-process has not been enqueued yet
        self._last = None     # engine of the most recent forward()
        self._post_stream = torch.cuda.Stream(self.device)
        self._pre_stream = torch.cuda.Stream(self.device)
        # The network graphs of consecutive batches (one per slot) run on their own streams, i.e. two batches are in flight
        # at once -- the thin, latency-bound P4/P5 and head layers of one graph fill the dependency bubbles of the other
        # (measured: +1.8 % on yolo11n batch 64, +3.9 % on BS-YOLO-n batch 256).  B2Y_DUAL_NET=0: both on the caller's stream
        # (batch i+1's network starts when batch i's ended).
        self.dual_net = len(self.engines) >= 2 and os.environ.get("B2Y_DUAL_NET", "1") != "0"
        self._net_streams = [torch.cuda.Stream(self.device) for _ in self.engines] if self.dual_net else None
        self._copy_stream = self._copy_stream2 = None
        self._fwd_done = [torch.cuda.Event() for _ in self.engines]
        self._pre_done = [torch.cuda.Event() for _ in self.engines]
        self._post_done = [torch.cuda.Event() for _ in self.engines]
        for e in self._post_done + self._fwd_done:
            e.record(torch.cuda.current_stream(self.device))
        # scale_boxes geometry is constant per pipeline (fixed source shape): one upload here, none on the post stream
        self._geom = scale_geometry((H, W), [self.src_shape] * batch, self.device)

    @property
    def launches_per_batch(self) -> int:
        return self.engine.n_launches + 1 + 2  # network ops + preprocess + NMS (score, select)

    def submit(self, src_u8: torch.Tensor, rescale: bool = False, consumed: torch.cuda.Event | None = None,
               gather: bool = False, host_out=None, src_ready=None) -> Ticket:
        """Enqueue one batch: src_u8 (B,h,w,3) uint8 BGR on this device.  The network goes to the current stream; the
        preprocess kernel runs on its own stream (it only needs the frames and this slot's input buffer) behind the
        mid-point marker of the previous batch's network; NMS (+ scale_boxes, + the multi-GPU gather, + the device->host
        copy into `host_out`) of the PREVIOUS batch is enqueued on the post-process stream behind this batch's marker.
        `src_ready`: an event after which the frames are complete, or True if they already are (resident inputs); None
        (default) orders the preprocess after everything enqueued on the current stream so far.
        Returns a Ticket; ``ticket.result()`` gives (dets (B,max_det,6), counts (B,), event) -- the tensors are valid once
        `event` has completed and must not be consumed on another stream before that.  ``consumed`` is recorded when the
        frames have been read."""
        k = self._slot
        self._slot = (k + 1) % len(self.engines)
        eng = self.engines[k]
        main = torch.cuda.current_stream(self.device)
        pre = self._pre_stream
        if src_ready is None:
            src_ready = torch.cuda.Event()
            src_ready.record(main)
        with torch.cuda.stream(pre):
            if src_ready is not True:
                pre.wait_event(src_ready)
            pre.wait_event(self._fwd_done[k])  # the network that last read this slot's input buffer has finished
            if self.defer and self._last is not None:
                self._last.wait_mid(pre)         # ... and the previous batch's network is past its wide layers
            src_u8.record_stream(pre)            # (allocator: the frames are read on this stream)
            eng.preprocess(src_u8, auto=self.auto)
            self._pre_done[k].record(pre)
            if consumed is not None:
                consumed.record(pre)
        if self._pending is not None and self._pending.slot == k:
            self._post(self._pending, None)      # (single slot: its y is about to be overwritten)
        net = main
        if self.dual_net:
            net = self._net_streams[k]
            ordered = torch.cuda.Event()
            ordered.record(main)
            net.wait_event(ordered)          # keeps the caller's stream order (everything enqueued before this submit)
        net.wait_event(self._pre_done[k])
        net.wait_event(self._post_done[k])   # the NMS that last read this slot's y has finished
        with torch.cuda.stream(net):
            eng.forward()
            self._fwd_done[k].record(net)
        self._last = eng
        ticket = Ticket(self, k, rescale, gather, host_out)
        if self.defer:
            if self._pending is not None:
                self._post(self._pending, eng)  # previous batch: under this network's thin layers
            self._pending = ticket
        else:
            self._post(ticket, None)
        return ticket

    def _post(self, t: Ticket, behind: Engine | None):
        """Enqueue the post-process of ticket `t` (behind the mid-point marker of `behind`'s latest forward, if given)."""
        k = t.slot
        eng = self.engines[k]
        post = self._post_stream
        with torch.cuda.stream(post):
            post.wait_event(self._fwd_done[k])
            if behind is not None:
                behind.wait_mid(post)
            dets, counts = nms_raw(eng.y, self.conf, self.iou, self.max_det, self.classes, self.agnostic)
            if t.rescale:
                scale_boxes_batched(dets, counts, geom=self._geom)
            if t.gather:
                from .shard import gather_detections
                dets, counts = gather_detections(dets, counts, dst=0 if t.gather == "root" else None)
            if t.host_out is not None:
                if t.host_out[0] is not None:
                    t.host_out[0].copy_(dets, non_blocking=True)
                t.host_out[1].copy_(counts, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(post)
            self._post_done[k].record(post)
        t.dets, t.counts, t.event = dets, counts, ev
        if self._pending is t:
            self._pending = None

    def flush(self):
        """Enqueue whatever post-process is still deferred (call before waiting for the last batch)."""
        if self._pending is not None:
            self._post(self._pending, None)

    def drain(self):
        """flush() and make the current stream wait for every slot's post-process (the end of a timed region)."""
        self.flush()
        main = torch.cuda.current_stream(self.device)
        for e in self._post_done + self._fwd_done:
            main.wait_event(e)

    def __call__(self, src_u8: torch.Tensor, rescale: bool = False, consumed: torch.cuda.Event | None = None):
        """One batch with plain stream semantics: returns (dets (B,max_det,6), counts (B,)) usable on the current stream."""
        dets, counts, ev = self.submit(src_u8, rescale, consumed).result()
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ev)
        dets.record_stream(main)
        counts.record_stream(main)
        return dets, counts

    def predict_host(self, frames: torch.Tensor, rescale: bool = False, gather=False):
        """frames: pinned host uint8 (B,h,w,3). Returns host (dets, counts); blocks until they are valid."""
        return next(self.predict_stream(iter([frames]), rescale, gather))

    def predict_stream(self, batches, rescale: bool = False, gather=False, device_dets: bool = False, uploaded=None):
        """Streaming end-to-end API (the analogue of ``predict(..., stream=True)``, engine/predictor.py:217-296): an
        iterable of pinned host uint8 batches (B,h,w,3) in, one host ``(dets, counts)`` pair per batch out, in order.
        The host->device copy of batch i+1 and the NMS + device->host copy of batch i-1 run on their own streams beside
        the network of batch i, and the result of batch i is handed out after batch i+1 has been enqueued, so neither
        PCIe nor the host-side wait idles the GPU (with two network graphs in flight the results run two batches behind).
        The yielded tensors are pinned buffers that stay valid until two more batches have been requested.
        gather (multi-GPU): False -- every rank yields its own shard; "root" -- rank 0 yields all ranks' detections
        (world*B rows, rank-major; one NCCL gather of the padded detections per batch), the other ranks their own shard;
        True / "all" -- every rank yields all ranks' detections (all-gather: world x the traffic, only if every rank
        really consumes them).
        device_dets: yield ``(dets on the device, counts on the host)`` -- for callers that keep boxes on the GPU (the
        drop-in predictor builds ``Results`` from device views); only the counts cross PCIe.
        uploaded: optional ``f(frames, event)`` called once the host->device copy of ``frames`` has been enqueued;
        ``event`` completes when the copy has finished and the pinned buffer may be refilled (staging rings)."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._copy_stream2 = torch.cuda.Stream(dev)
        cs, cs2 = self._copy_stream, self._copy_stream2
        split = int(os.environ.get("B2Y_H2D_SPLIT", "1")) >= 2  # A/B knob: the upload as two halves on two copy streams
        # results are handed out `lag` batches behind the submission: 1 keeps one network in flight while the host waits,
        # 2 (two network graphs in flight, dual_net) lets batch i+1's graph be enqueued while batch i's has barely started
        lag = 2 if self.dual_net else 1
        R = lag + 1
        stage, copied, consumed, tickets, outs, half = ([None] * R for _ in range(6))
        for k in range(R):
            copied[k], consumed[k], half[k] = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

        def upload(frames, k):
            if stage[k] is None or stage[k].shape != frames.shape:
                stage[k] = torch.empty(frames.shape, dtype=torch.uint8, device=dev)
            n = frames.shape[0] // 2
            if split and n:
                with torch.cuda.stream(cs2):
                    cs2.wait_event(consumed[k])
                    stage[k][n:].copy_(frames[n:], non_blocking=True)
                    half[k].record(cs2)
            with torch.cuda.stream(cs):
                cs.wait_event(consumed[k])  # the preprocess kernel that last read this slot has finished
                if split and n:
                    stage[k][:n].copy_(frames[:n], non_blocking=True)
                    cs.wait_event(half[k])
                else:
                    stage[k].copy_(frames, non_blocking=True)
                copied[k].record(cs)
                if uploaded is not None:
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    uploaded(frames, ev)

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        for e in consumed:
            e.record(main)
        upload(nxt, 0)
        rows = self.B
        if gather:
            import torch.distributed as dist
            if gather != "root" or dist.get_rank() == 0:
                rows = dist.get_world_size() * self.B
        i = 0
        while nxt is not None:
            k = i % R
            nxt = next(it, None)
            if nxt is not None:
                upload(nxt, (i + 1) % R)  # overlaps the compute of batch i
            if outs[k] is None:
                outs[k] = (None if device_dets else torch.empty(rows, self.max_det, 6, dtype=torch.float32).pin_memory(),
                           torch.empty(rows, dtype=torch.int32).pin_memory())
            tickets[k] = self.submit(stage[k], rescale, consumed=consumed[k], gather=gather, host_out=outs[k],
                                     src_ready=copied[k])
            if i >= lag:  # batch i-lag: its post-process was enqueued by an earlier submit
                j = (i - lag) % R
                yield self._resolve(tickets[j], outs[j], device_dets, main)
            i += 1
        for b in range(max(i - lag, 0), i):  # the tail, in order (the last one is flushed here)
            j = b % R
            yield self._resolve(tickets[j], outs[j], device_dets, main)

    @staticmethod
    def _resolve(ticket, out, device_dets, main):
        dets, _, ev = ticket.result()
        ev.synchronize()
        if not device_dets:
            return out
        dets.record_stream(main)  # allocated on the post-process stream, consumed by the caller's
        return dets, out[1]

"""Drop-in validator: ``model.val()`` for a detect model with every per-batch stage on the B200 path.

    from bs_yolo_b200.validator import B200DetectionValidator
    metrics = YOLO("best.pt").val(data="data.yaml", validator=B200DetectionValidator, device=0)   # engine/model.py:638-641

What the reference does per batch (engine/validator.py:160-190) and what runs here instead:

    preprocess      DetectionValidator.preprocess (detect/val.py:47-50: uint8 -> float / 255)   -> b2y_preprocess_chw
    inference       AutoBackend.forward -> DetectionModel._predict_once                          -> CUDA-graphed plan
    postprocess     ops.non_max_suppression(conf=0.001, multi_label=True) (detect/val.py:92-102) -> b2y_nms
    update_metrics  per image: scale_boxes to the original image (val.py:117-123), box_iou +
                    match_predictions (val.py:209-228, validator.py:222-258)                     -> b2y_scale_boxes +
                                                                                                    b2y_match_predictions,
                                                                                                    ONE launch each per batch

Everything around it is the reference's own code, inherited: dataset / dataloader construction, label parsing, the
AP / mAP arithmetic of ``DetMetrics`` (numpy, on the host), printing, callbacks, ``save_json`` / ``save_txt`` / ``plots``.
``BaseValidator.__call__`` builds an ``AutoBackend`` itself (validator.py:125-131), which is why this class restates that
method instead of overriding a hook.  Labels are few and stay on the host: they are converted to original-image xyxy with
the reference's own op sequence (val.py:112-114) on CPU tensors and uploaded once per batch.

No CPU fallback: a non-CUDA device, ``half=True``, ``augment``, ``save_hybrid`` and validation inside a training loop raise.
"""
from __future__ import annotations

import weakref

import torch

from . import _cabi as A
from .ops import match_predictions_batched, nms_raw, preprocess_chw

_CLS = None
_BACKENDS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()  # model -> (device, weights fingerprint, B200Backend)


def _fingerprint(model) -> int:
    """Changes whenever a parameter or buffer is replaced or written in place (optimizer step, load_state_dict, .half())."""
    return hash(tuple((t.data_ptr(), t._version) for t in list(model.parameters()) + list(model.buffers())))


def _backend_for(model, device):
    """Lowered plans are kept per live model object, so that a second ``model.val()`` does not lower, autotune and
    capture again; a model whose weights changed since is lowered afresh."""
    from .predictor import B200Backend

    fp = _fingerprint(model)
    hit = _BACKENDS.get(model)
    if hit is not None and hit[0] == device and hit[1] == fp:
        return hit[2]
    backend = B200Backend(model, device, head="detect", who="B200DetectionValidator")
    backend.max_engines = 8  # rect batches: one plan per (batch, H, W); keep more of them than the predictor does
    _BACKENDS[model] = (device, fp, backend)
    return backend


def _label_rows(batch, nb: int, imgsz, device):
    """All labels of a batch in original-image xyxy, padded per image: (B, max_labels, 5) fp32 [cls, x1, y1, x2, y2] and
    counts (B,) int32 on `device`, plus the per-image host tensors (cls, bbox) the reference's stats need.
    Same fp32 op order as DetectionValidator._prepare_batch (val.py:104-115): xywh2xyxy(bbox) * (w,h,w,h); then scale_boxes
    with the dataset's ratio_pad (ops.py:117-127): -= pad, /= gain, clip to the original shape."""
    h, w = imgsz
    idx = batch["batch_idx"].cpu().long()
    cls_all = batch["cls"].cpu().float().reshape(-1)
    xywh = batch["bboxes"].cpu().float().reshape(-1, 4)
    half = xywh[:, 2:] / 2
    xyxy = torch.cat([xywh[:, :2] - half, xywh[:, :2] + half], 1) * torch.tensor([w, h, w, h], dtype=torch.float32)
    per_image = []
    for si in range(nb):
        sel = idx == si
        b = xyxy[sel].clone()
        if len(b):
            gain, pad = batch["ratio_pad"][si][0][0], batch["ratio_pad"][si][1]
            oh, ow = batch["ori_shape"][si]
            b[:, 0] -= pad[0]
            b[:, 1] -= pad[1]
            b[:, 2] -= pad[0]
            b[:, 3] -= pad[1]
            b /= gain
            b[:, 0].clamp_(0, ow)
            b[:, 1].clamp_(0, oh)
            b[:, 2].clamp_(0, ow)
            b[:, 3].clamp_(0, oh)
        per_image.append((cls_all[sel], b))
    max_labels = max(1, max(len(c) for c, _ in per_image))
    rows = torch.zeros(nb, max_labels, 5, dtype=torch.float32)
    counts = torch.zeros(nb, dtype=torch.int32)
    if torch.device(device).type == "cuda":
        rows, counts = rows.pin_memory(), counts.pin_memory()
    for si, (c, b) in enumerate(per_image):
        if len(c):
            rows[si, : len(c), 0] = c
            rows[si, : len(c), 1:] = b
            counts[si] = len(c)
    return rows.to(device, non_blocking=True), counts.to(device, non_blocking=True), per_image


def _build_class():
    from ultralytics.models.yolo.detect.val import DetectionValidator
    from ultralytics.utils import LOGGER, TQDM, callbacks
    from ultralytics.utils.checks import check_imgsz
    from ultralytics.utils.ops import Profile
    from ultralytics.utils.torch_utils import select_device

    from .predictor import PreprocessedBatch

    class B200DetectionValidator(DetectionValidator):
        """DetectionValidator whose preprocess / inference / NMS / box scaling / TP matching run as sm_100a kernels."""

        capture = None  # tests: a list that receives (predn, tp, counts, per-image labels) of every batch (host tensors)

        def __call__(self, trainer=None, model=None):
            if trainer is not None:
                raise NotImplementedError("validation inside a training loop is not on the B200 path (training is out of scope)")
            a = self.args
            if a.half or a.augment or a.save_hybrid:
                raise NotImplementedError("half / augment / save_hybrid are not on the B200 validator path")
            with torch.inference_mode():
                return self._run(model)

        # ---- BaseValidator.__call__ (engine/validator.py:107-213), non-training branch, B200Backend for AutoBackend
        def _run(self, model):
            from ultralytics.data.utils import check_det_dataset

            a = self.args
            self.training = False
            callbacks.add_integration_callbacks(self)
            model = model if model is not None else a.model
            if isinstance(model, (str, bytes)) or hasattr(model, "__fspath__"):
                from ultralytics.nn.tasks import attempt_load_weights  # the reference's own loader, unchanged
                model = attempt_load_weights(str(model), device="cpu", inplace=True, fuse=False)
            backend = _backend_for(model, torch.device(select_device(a.device, a.batch)))
            self.backend, self.device = backend, backend.device
            a.half = False
            self.stride = backend.stride
            check_imgsz(a.imgsz, stride=self.stride)
            if str(a.data).split(".")[-1] not in {"yaml", "yml"}:
                raise FileNotFoundError(f"Dataset '{a.data}' for task={a.task} not found")
            self.data = check_det_dataset(a.data)
            self.dataloader = self.dataloader or self.get_dataloader(self.data.get(a.split), a.batch)

            self.run_callbacks("on_val_start")
            dt = tuple(Profile(device=self.device) for _ in range(4))
            self.init_metrics(backend)
            self.jdict = []
            for batch_i, batch in enumerate(TQDM(self.dataloader, desc=self.get_desc(), total=len(self.dataloader))):
                self.run_callbacks("on_val_batch_start")
                self.batch_i = batch_i
                with dt[0]:
                    batch = self.preprocess(batch)
                with dt[1]:
                    preds = backend(self._net_input)
                with dt[3]:
                    preds = self.postprocess(preds)
                self.update_metrics(preds, batch)
                if a.plots and batch_i < 3:
                    self.plot_val_samples(batch, batch_i)
                    self.plot_predictions(batch, self._pred_list(preds), batch_i)
                self.run_callbacks("on_val_batch_end")
            stats = self.get_stats()
            self.check_stats(stats)
            self.speed = dict(zip(self.speed.keys(), (x.t / len(self.dataloader.dataset) * 1e3 for x in dt)))
            self.finalize_metrics()
            self.print_results()
            self.run_callbacks("on_val_end")
            LOGGER.info("Speed: {:.1f}ms preprocess, {:.1f}ms inference, {:.1f}ms loss, {:.1f}ms postprocess per image".format(
                *tuple(self.speed.values())))
            if a.save_json and self.jdict:
                import json

                with open(str(self.save_dir / "predictions.json"), "w") as f:
                    json.dump(self.jdict, f)
                stats = self.eval_json(stats)
            return stats

        def preprocess(self, batch):
            """detect/val.py:47-50.  batch["img"] stays the uint8 (B,3,H,W) tensor (now on the device; the plotting helpers
            take it as is), the normalised bf16 copy is written straight into the plan's input buffer.  Labels stay on the host."""
            img = batch["img"]
            if img.dtype != torch.uint8:
                raise TypeError("B200DetectionValidator expects the dataloader's uint8 image batch")
            if not img.is_pinned():
                img = img.pin_memory()
            img = img.to(self.device, non_blocking=True)
            batch["img"] = img
            B, _, H, W = (int(v) for v in img.shape)
            eng = self.backend.engine(B, H, W)
            with torch.cuda.device(self.device):
                preprocess_chw(img, eng.input_nhwc)
            self._net_input = PreprocessedBatch(eng)
            return batch

        def postprocess(self, preds):
            """detect/val.py:92-102: NMS with multi_label=True at the validator's conf (0.001 by default).  Returns the
            padded (dets (B, max_det, 6), counts (B,)) pair; nothing crosses PCIe here."""
            y = preds[0] if isinstance(preds, (list, tuple)) else preds
            a = self.args
            with torch.cuda.device(self.device):
                return nms_raw(y, a.conf, a.iou, a.max_det, None, bool(a.single_cls or a.agnostic_nms), multi_label=True)

        @staticmethod
        def _pred_list(preds):
            dets, counts = preds
            return [dets[i, :n] for i, n in enumerate(counts.tolist())]

        def update_metrics(self, preds, batch):
            """detect/val.py:125-172 for the whole batch: predictions -> original-image coordinates (b2y_scale_boxes with the
            dataset's ratio_pad), true-positive matrix (b2y_match_predictions), one device->host copy, then the reference's
            per-image bookkeeping on host tensors."""
            dets, counts = preds
            B, max_det, _ = dets.shape
            imgsz = tuple(int(v) for v in batch["img"].shape[2:])
            with torch.cuda.device(self.device):
                if self.args.single_cls:
                    dets[:, :, 5] = 0
                geom = torch.tensor([[rp[0][0], rp[1][0], rp[1][1], s[1], s[0]]
                                     for rp, s in zip(batch["ratio_pad"], batch["ori_shape"])], dtype=torch.float32)
                predn = dets.clone()  # the letterbox-space boxes stay intact for plot_predictions / callbacks
                A.check(A.lib().b2y_scale_boxes(predn.data_ptr(), counts.data_ptr(),
                                                geom.pin_memory().to(self.device, non_blocking=True).data_ptr(), B, max_det,
                                                torch.cuda.current_stream(self.device).cuda_stream))
                rows, nl_dev, per_image = _label_rows(batch, B, imgsz, self.device)
                tp = match_predictions_batched(predn, counts, rows, nl_dev, self.iouv)
                predn_h, tp_h, n_h = predn.cpu(), tp.cpu(), counts.tolist()  # the one device->host sync of the batch
            if self.capture is not None:
                self.capture.append((predn_h, tp_h, n_h, per_image))
            for si in range(B):
                self.seen += 1
                npr = n_h[si]
                cls, bbox = per_image[si]
                nl = len(cls)
                pn = predn_h[si, :npr]
                stat = dict(conf=pn[:, 4], pred_cls=pn[:, 5], tp=tp_h[si, :npr], target_cls=cls, target_img=cls.unique())
                if npr == 0:
                    if nl:
                        stat["conf"], stat["pred_cls"] = torch.zeros(0), torch.zeros(0)
                        for k in self.stats.keys():
                            self.stats[k].append(stat[k])
                        if self.args.plots:
                            self.confusion_matrix.process_batch(detections=None, gt_bboxes=bbox, gt_cls=cls)
                    continue
                if self.args.plots:
                    self.confusion_matrix.process_batch(pn, bbox, cls)
                for k in self.stats.keys():
                    self.stats[k].append(stat[k])
                if self.args.save_json:
                    self.pred_to_json(pn, batch["im_file"][si])
                if self.args.save_txt:
                    from pathlib import Path

                    self.save_one_txt(pn, self.args.save_conf, batch["ori_shape"][si],
                                      self.save_dir / "labels" / f'{Path(batch["im_file"][si]).stem}.txt')

    return B200DetectionValidator


def __getattr__(name):  # PEP 562: resolve the class on first use so importing this module never imports ultralytics
    global _CLS
    if name == "B200DetectionValidator":
        if _CLS is None:
            _CLS = _build_class()
        return _CLS
    raise AttributeError(name)

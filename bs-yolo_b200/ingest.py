"""Ingest for file sources: ``predict("dir/of/images", batch=64)`` without the host becoming the step.

The reference's ``LoadImagesAndVideos`` (ultralytics/data/loaders.py:284-449) decodes the files of a batch one after the
other on the caller's thread (``imread``, utils/patches.py: np.fromfile + cv2.imdecode), and ``BasePredictor.preprocess``
(engine/predictor.py:116-161) then letterboxes them one by one in Python, stacks and uploads from pageable memory.
``PrefetchImageLoader`` stands where that loader stands -- same iteration protocol ``(paths, im0s, info)``, same batch
boundaries, same info strings, same skipping of unreadable files, same attributes (files, nf, ni, bs, mode, count,
video_flag, cap, source_type) -- and

  * decodes the files of up to two batches AHEAD on a thread pool (np.fromfile / cv2.imdecode release the GIL), with the
    reference's own ``imread``;
  * when the images of a batch share one shape, copies them -- still on the pool, off the caller's thread -- into a ring of
    pinned ``(bs, h, w, 3)`` uint8 blocks, and hands the batch out as a ``PinnedFrames`` list: the decoded arrays (what
    ``Results.orig_img`` keeps, exactly as with the reference loader) plus the pinned block, which
    ``B200DetectionPredictor`` uploads directly (no staging copy on the caller's thread, no stack, no pageable H2D).

LetterBox, BGR->RGB, /255 and the layout change happen on the GPU (b2y_preprocess), so nothing of predictor.py:116-161 is
left on the host.  Video files and HEIC images keep the reference's loader (sequential decoders).
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

RING = 3  # pinned blocks; the producer stays at most RING - 1 batches ahead of the batch being consumed


class PinnedFrames(list):
    """A batch of decoded BGR frames (a plain list for every consumer that expects one) whose pixels also sit, in order,
    in rows [0, len) of the pinned uint8 block ``pinned`` ((bs, h, w, 3)); ``pinned`` is None for ragged batches."""

    pinned: torch.Tensor | None = None


class PrefetchImageLoader:
    def __init__(self, loader, imread=None, workers: int | None = None):
        """`loader`: a reference LoadImagesAndVideos over image files only (its file list and batch size are taken over)."""
        if any(loader.video_flag) or any(f.rsplit(".", 1)[-1].lower() == "heic" for f in loader.files):
            raise ValueError("PrefetchImageLoader handles image files only (videos / HEIC keep the reference loader)")
        if imread is None:
            from ultralytics.utils.patches import imread  # the reference's own decoder call
        self._imread = imread
        self.files, self.nf, self.ni = list(loader.files), loader.nf, loader.ni
        self.video_flag, self.mode, self.vid_stride, self.bs, self.cap = [False] * self.nf, "image", loader.vid_stride, loader.bs, None
        if hasattr(loader, "source_type"):
            self.source_type = loader.source_type
        self.count = 0
        self._pool = ThreadPoolExecutor(max_workers=workers or max(2, min(32, (os.cpu_count() or 2))))
        self._ring: list[torch.Tensor | None] = [None] * RING
        self._pin = torch.cuda.is_available()
        self._q: queue.Queue | None = None
        self._go: threading.Semaphore | None = None
        self._stop = threading.Event()
        self._done = False

    def __len__(self):
        return -(-self.nf // self.bs)

    # ------------------------------------------------------------------ producer (background thread)
    def _block(self, slot: int, shape) -> torch.Tensor:
        b = self._ring[slot]
        if b is None or tuple(b.shape[1:]) != tuple(shape):
            b = torch.empty((self.bs, *shape), dtype=torch.uint8)
            b = b.pin_memory() if self._pin else b
            self._ring[slot] = b
        return b

    def _produce(self, q: queue.Queue, go: threading.Semaphore, stop: threading.Event):
        try:
            pos, k = 0, 0
            pending: list = []  # (file index, image) decoded but not yet handed out, in file order
            while not stop.is_set():
                while pos < self.nf:  # decode windows of files in parallel until bs readable images are waiting
                    ready = sum(im is not None for _, im in pending)
                    if ready >= self.bs:
                        break
                    hi = min(pos + self.bs - ready, self.nf)
                    for idx, im in zip(range(pos, hi), self._pool.map(self._imread, self.files[pos:hi])):
                        pending.append((idx, im))  # unreadable files stay in the list as None: they carry the count forward
                    pos = hi
                if not pending:
                    break
                batch, rest, n_ok = [], [], 0
                for p in pending:  # the next bs readable images; the files skipped on the way are consumed with them
                    (batch if n_ok < self.bs else rest).append(p)
                    n_ok += p[1] is not None and n_ok < self.bs
                pending = rest
                frames = PinnedFrames(im for _, im in batch if im is not None)
                paths = [self.files[i] for i, im in batch if im is not None]
                info = [f"image {i + 1}/{self.nf} {self.files[i]}: " for i, im in batch if im is not None]
                bad = [self.files[i] for i, im in batch if im is None]
                go.acquire()  # the ring slot of batch k is free once batch k - RING + 1 has been handed out
                if stop.is_set():
                    break
                if frames and len({f.shape for f in frames}) == 1 and frames[0].dtype == np.uint8 and frames[0].ndim == 3:
                    blk = self._block(k % RING, frames[0].shape)
                    views = blk.numpy()
                    list(self._pool.map(lambda a: np.copyto(views[a[0]], a[1]), enumerate(frames)))
                    frames.pinned = blk
                q.put((batch[-1][0] + 1, paths, frames, info, bad))
                k += 1
            q.put(None)
        except BaseException as e:  # noqa: BLE001 -- surfaced on the consumer's thread
            q.put(e)

    # ------------------------------------------------------------------ iteration protocol of the reference loader
    def __iter__(self):
        self.close()
        self.count, self._done = 0, False
        self._stop = threading.Event()
        self._q = queue.Queue()
        self._go = threading.Semaphore(RING - 1)
        threading.Thread(target=self._produce, args=(self._q, self._go, self._stop), daemon=True).start()
        return self

    def __next__(self):
        if self._done:
            raise StopIteration
        if self._q is None:
            iter(self)
        while True:
            item = self._q.get()
            if item is None or isinstance(item, BaseException):
                self._q, self._done = None, True
                if item is None:
                    raise StopIteration
                raise item
            self._go.release()  # asking for this batch means the previous one has been fully consumed
            self.count, paths, frames, info, bad = item
            for f in bad:
                from ultralytics.utils import LOGGER

                LOGGER.warning(f"WARNING ⚠️ Image Read Error {f}")
            if frames:
                return paths, frames, info

    def close(self):
        if self._q is not None:
            self._stop.set()
            if self._go is not None:
                self._go.release()
            self._q = None

    def __del__(self):
        try:
            self.close()
            self._pool.shutdown(wait=False)
        except Exception:
            pass

"""Image I/O on the edges of the b200 path (mirrors fpie/io.py:10-39).

Same three functions, same results -- ``read_image`` (BGR uint8, grey images replicated to three
channels, an alpha channel dropped, ``FileNotFoundError`` for an unreadable file: io.py:10-19),
``write_image`` (io.py:22-24) and ``read_images`` (an absent mask file means "the whole source":
io.py:27-39) -- arranged for the device path on either side of them:

* the three files are decoded CONCURRENTLY (``cv2.imread`` releases the GIL) and each decoded image is
  staged into a recycled PAGE-LOCKED buffer, so that ``Processor.reset``'s upload is one DMA at PCIe
  speed instead of the driver's bounce-buffer copy of pageable memory;
* ``ImageWriter`` encodes in a worker thread from a snapshot of the result (the Processor hands out the
  same canvas on every ``step``, process.py:278-279), so that the PNG encoder of ``-p`` progress images
  (hundreds of milliseconds for a 4096^2 canvas) runs beside the next ``step`` instead of between two.

Nothing here touches the device: without a CUDA runtime the staging buffers are ordinary arrays.
"""

from __future__ import annotations

import atexit
import os
import queue
import threading
import warnings

import numpy as np

from . import _lib

_POOL_LOCK = threading.Lock()
_POOL: dict = {}  # (shape) -> page-locked arrays handed out before, reused once their holder let go
_POOL_KEEP = 6


def _staging(shape) -> np.ndarray:
    """A page-locked uint8 array of ``shape`` (pageable when page-locked memory is unavailable)."""
    import sys

    shape = tuple(int(v) for v in shape)
    with _POOL_LOCK:
        for buf in _POOL.get(shape, ()):
            base = buf.base
            # references: the pool's list + `buf` + getrefcount's argument (see process._reset_with_canvas)
            free = sys.getrefcount(buf) <= 3 and (base is None or sys.getrefcount(base) <= 3)
            del base
            if free:
                return buf
        try:
            buf = _lib.pinned_empty(shape, np.uint8)
        except (RuntimeError, OSError):
            return np.empty(shape, np.uint8)
        held = _POOL.setdefault(shape, [])
        held.append(buf)
        del held[:-_POOL_KEEP]
        return buf


def _three_channels(img: np.ndarray) -> np.ndarray:
    if img.ndim == 2:  # io.py:15-16
        return np.stack([img, img, img], axis=-1)
    if img.ndim == 3 and img.shape[-1] == 4:  # io.py:17-18
        return img[..., :-1]
    return img


def read_image(name: str, pinned: bool = True) -> np.ndarray:
    """``fpie.io.read_image`` (io.py:10-19); ``pinned`` stages the result in page-locked memory."""
    import cv2

    img = cv2.imread(name)
    if img is None:
        raise FileNotFoundError(f"Failed to read image: {name}")
    img = _three_channels(img)
    if not pinned:
        return np.ascontiguousarray(img)
    out = _staging(img.shape)
    np.copyto(out, img)
    return out


def write_image(name: str, image: np.ndarray) -> None:
    """``fpie.io.write_image`` (io.py:22-24), synchronous."""
    import cv2

    cv2.imwrite(name, image)


def read_images(src_name: str, mask_name: str, tgt_name: str, pinned: bool = True):
    """``fpie.io.read_images`` (io.py:27-39): the files are decoded and staged side by side."""
    names = [src_name, tgt_name] + ([mask_name] if os.path.exists(mask_name) else [])
    out: list = [None] * len(names)
    errors: list = []

    def work(i):
        try:
            out[i] = read_image(names[i], pinned)
        except BaseException as exc:  # re-raised in the caller's thread, in the reference's order
            errors.append((i, exc))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(names))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise min(errors, key=lambda e: e[0])[1]
    src, tgt = out[0], out[1]
    if len(names) == 3:
        mask = out[2]
    else:
        warnings.warn("No mask file found, use default setting", stacklevel=2)
        mask = _staging(src.shape) if pinned else np.empty(src.shape, np.uint8)
        mask[...] = 255  # io.py:38 `np.zeros_like(src) + 255`
    return src, mask, tgt


class ImageWriter:
    """``write_image`` whose encoder runs beside the caller: ``write`` snapshots ``image`` and returns;
    ``flush`` waits for everything queued (and re-raises the first failure)."""

    def __init__(self, depth: int = 2):
        self._q: queue.Queue = queue.Queue(maxsize=depth)
        self._err = None
        self._thread = None
        self._lock = threading.Lock()

    def _run(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                if self._err is None:
                    write_image(*item)
            except BaseException as exc:
                self._err = exc
            finally:
                self._q.task_done()

    def write(self, name: str, image: np.ndarray) -> None:
        with self._lock:
            if self._thread is None or not self._thread.is_alive():
                self._thread = threading.Thread(target=self._run, daemon=True)
                self._thread.start()
        self._q.put((name, np.array(image, copy=True)))  # (blocks when `depth` images are waiting)

    def flush(self) -> None:
        self._q.join()
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def close(self) -> None:
        self.flush()
        with self._lock:
            if self._thread is not None and self._thread.is_alive():
                self._q.put(None)
                self._thread.join()
            self._thread = None


_WRITER = ImageWriter()
atexit.register(_WRITER.close)


def write_image_async(name: str, image: np.ndarray) -> None:
    """Queue ``image`` for writing; ``flush_writes()`` (also run at interpreter exit) completes it."""
    _WRITER.write(name, image)


def flush_writes() -> None:
    _WRITER.flush()

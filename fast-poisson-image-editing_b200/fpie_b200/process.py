"""Processor layer for the b200 backend.

Mirrors ``fpie.process.EquProcessor`` / ``GridProcessor`` (fpie/process.py:146-395):
same constructor argument order (fpie/cli.py:16-32 passes them positionally),
``reset(src, mask, tgt, mask_on_src, mask_on_tgt) -> int``, ``sync()`` and
``step(iteration) -> (uint8 target image, err[3])``, returning the same
full-size target buffer on every call (process.py:278-279, 393-394).

The difference is where ``reset`` runs: the reference builds the linear
system / gradient field on the host with numpy and hands fp32 arrays to the
core; here ``reset`` uploads the three uint8 images once and every step of
the preprocessing (mask threshold, frame clear, bounding box, id scan, A/X/B
or gradient build) is a CUDA kernel (``*_reset_from_images`` in
include/fpie_b200.h).
"""

from __future__ import annotations

import sys
import threading

import numpy as np

from . import _lib
from .solver import EquSolver, GridSolver

BACKEND = "b200"
GRADIENTS = ("max", "src", "avg")


class BaseProcessor:
    """Common state of both processors (fpie/process.py:86-143)."""

    def __init__(self, gradient: str, backend: str, core):
        if gradient not in GRADIENTS:
            raise ValueError(f"gradient must be one of {GRADIENTS}, got {gradient!r}")
        if backend != BACKEND:
            raise AssertionError(f"Invalid backend {backend}.")  # process.py:105
        self.gradient = gradient
        self.backend = backend
        self.core = core
        self.rank = 0
        self.root = True
        self._canvas_stale = False
        self._tgt = None
        self.box = None
        self._canvas_pool = []

    def _require_reset(self) -> None:
        if self._tgt is None or self.box is None:  # the same error the core raises (GridSolver::require_ready)
            raise RuntimeError(f"{type(self).__name__}: step called before reset")

    def sync(self) -> None:
        self.core.sync()

    def solve(self, max_iters: int, tol: float, check_every: int = 100):
        """Convergence-driven ``step``: sweep until ``max(err) <= tol`` or ``max_iters``; returns
        ``(tgt, err, iterations_run)`` (not part of the reference Processor)."""
        _, done = self.core.solve(max_iters, tol, check_every)
        out, err = self.step(0)
        return out, err, done

    def _reset_with_canvas(self, tgt, device_reset):
        """Run the device-side reset while a worker thread makes the Processor's private copy of
        the target (process.py:268 / 384); both release the GIL, so they overlap."""
        # page-locked, so that every step's device-to-host copy of the result runs at PCIe speed.  The
        # reference hands out a fresh copy per reset (`tgt.copy()`); page-locking is too slow for that (tens of
        # milliseconds for a 50 MB image), so canvases are recycled from a small pool -- but only ones that
        # nobody outside this object can still see (a caller that keeps the previous result keeps it intact)
        self.tgt = None
        canvas = None
        old = None
        for old in self._canvas_pool:
            base = old.base
            # references: the pool's + `old` + getrefcount's argument; base: the canvas + `base` + the argument
            private = sys.getrefcount(old) <= 3 and (base is None or sys.getrefcount(base) <= 3)
            del base
            if private and old.shape == tgt.shape:
                canvas = old
                break
        del old
        if canvas is None:
            try:
                canvas = _lib.pinned_empty(tgt.shape, np.uint8)
            except RuntimeError:  # page-locked memory exhausted: an ordinary array works, only slower
                canvas = np.empty(tgt.shape, np.uint8)
            self._canvas_pool.append(canvas)
            del self._canvas_pool[:-3]  # (a dropped canvas lives on for as long as its holder keeps it)
        rows, cols = tgt.shape[0], tgt.shape[1]
        parts = 4 if tgt.size >= (8 << 20) else 1  # large images: fault the fresh pages in from several threads
        bounds = [rows * i // parts for i in range(parts + 1)]
        block = max(1, (1 << 20) // max(1, tgt.strides[0]))  # rows per copy call: about 1 MB
        known = [None]  # the blend's bounding box (x0, x1, y0, y1) in the target, once the device has found it

        def copy(lo, hi):
            # The device returns the blended crop into canvas[x0:x1, y0:y1] (at zero sweeps: the target's own pixels),
            # so only what lies OUTSIDE the box has to come from `tgt` -- a 4096^2 target whose box is the whole image
            # costs nothing beyond the device-side reset instead of outlasting it by a millisecond.
            r = lo
            while r < hi:
                e = min(hi, r + block)
                box = known[0]
                a, b = (max(r, box[0]), min(e, box[1])) if box else (e, e)
                if a >= b:
                    np.copyto(canvas[r:e], tgt[r:e], casting="unsafe")
                else:
                    if r < a:
                        np.copyto(canvas[r:a], tgt[r:a], casting="unsafe")
                    if b < e:
                        np.copyto(canvas[b:e], tgt[b:e], casting="unsafe")
                    if box[2] > 0:
                        np.copyto(canvas[a:b, : box[2]], tgt[a:b, : box[2]], casting="unsafe")
                    if box[3] < cols:
                        np.copyto(canvas[a:b, box[3] :], tgt[a:b, box[3] :], casting="unsafe")
                r = e

        workers = []

        def start(box=None):
            # Called by the core as soon as the box is known (the mask travels first, the images after it): with the
            # box covering the whole target there is nothing to copy; otherwise the workers run beside the rest of
            # the device-side reset.  (Also called with no box by the fallback below.)
            if workers:
                return
            known[0] = box
            if box and box[0] <= 0 and box[1] >= rows and box[2] <= 0 and box[3] >= cols:
                workers.append(None)
                return
            for i in range(parts):
                wk = threading.Thread(target=copy, args=(bounds[i], bounds[i + 1]))
                wk.start()
                workers.append(wk)

        hook = getattr(self.core, "on_box", None)
        if hook is not None:
            hook(start)
        else:
            start()
        try:
            result = device_reset()
            if not workers:  # (a core that never announced the box)
                start(tuple(int(v) for v in result[1]))
        finally:
            if hook is not None:
                hook(None)
            for wk in workers:  # (every read of the caller's `tgt` has completed when reset returns)
                if wk is not None:
                    wk.join()
        self._tgt = canvas
        self._canvas_stale = True  # the inside of the box is still to come from the device: `tgt` / `step` fetch it
        return result

    def _fetch_into_canvas(self, iteration: int):
        raise NotImplementedError

    @property
    def tgt(self):
        """The Processor's private copy of the target (process.py:268 / 384), holding the blend after ``step``."""
        if self._canvas_stale and self._tgt is not None and self.box is not None:
            self._fetch_into_canvas(0)
        return self._tgt

    @tgt.setter
    def tgt(self, value) -> None:
        self._tgt = value
        self._canvas_stale = False

    @staticmethod
    def _check_images(src, mask, tgt):
        for name, img in (("src", src), ("tgt", tgt)):
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError(f"{name} must be [rows, cols, 3]")
        if mask.ndim not in (2, 3):
            raise ValueError("mask must be [rows, cols] or [rows, cols, channels]")


class EquProcessor(BaseProcessor):
    """PIE Jacobi equation processor on the b200 core (fpie/process.py:146-280)."""

    def __init__(self, gradient: str = "max", backend: str = BACKEND, n_cpu: int = 0, min_interval: int = 100,
                 block_size: int = 256, device: int | None = None, mode: str = "jacobi"):
        super().__init__(gradient, backend, EquSolver(block_size, device=device, mode=mode))

    def reset(self, src, mask, tgt, mask_on_src=(0, 0), mask_on_tgt=(0, 0)) -> int:
        src, mask, tgt = np.asarray(src), np.asarray(mask), np.asarray(tgt)
        self._check_images(src, mask, tgt)
        n, box = self._reset_with_canvas(
            tgt, lambda: self.core.reset_from_images(src, mask, tgt, mask_on_src, mask_on_tgt, self.gradient))
        self.box = box
        return n

    def step(self, iteration: int):
        # process.py:278 `self.tgt[self.tgt_index] = x[1:]`: the device scatters the K solved pixels into
        # its copy of the crop, the device-to-host copy lands it in self.tgt[x0:x1, y0:y1]
        self._require_reset()
        return self._tgt, self._fetch_into_canvas(iteration)

    def _fetch_into_canvas(self, iteration: int):
        x0, _, y0, _ = self.box
        err = self.core.step_paste_into(iteration, self._tgt, x0, y0)
        self._canvas_stale = False
        return err


class GridProcessor(BaseProcessor):
    """PIE grid processor on the b200 core (fpie/process.py:283-395)."""

    def __init__(self, gradient: str = "max", backend: str = BACKEND, n_cpu: int = 0, min_interval: int = 100,
                 block_size: int = 256, grid_x: int = 8, grid_y: int = 8, device: int | None = None,
                 block_k: int = 0):
        super().__init__(gradient, backend, GridSolver(grid_x, grid_y, device=device, block_k=block_k))

    def reset(self, src, mask, tgt, mask_on_src=(0, 0), mask_on_tgt=(0, 0)) -> int:
        src, mask, tgt = np.asarray(src), np.asarray(mask), np.asarray(tgt)
        self._check_images(src, mask, tgt)
        n, box = self._reset_with_canvas(
            tgt, lambda: self.core.reset_from_images(src, mask, tgt, mask_on_src, mask_on_tgt, self.gradient))
        self.box = box
        self.x0, self.x1, self.y0, self.y1 = box
        return n

    def step(self, iteration: int):
        # process.py:393 `self.tgt[x0:x1, y0:y1] = tgt`, done by the device-to-host copy itself
        self._require_reset()
        return self._tgt, self._fetch_into_canvas(iteration)

    def _fetch_into_canvas(self, iteration: int):
        err = self.core.step_into(iteration, self._tgt, self.x0, self.y0)
        self._canvas_stale = False
        return err


class BatchGridProcessor(BaseProcessor):
    """Many small, independent edits at once (the GUI's reset + step per click,
    fpie/gui.py:96-99, batched): ``reset`` takes stacks of images, ``step``
    returns the blended stack and one ``err`` triple per edit."""

    def __init__(self, gradient: str = "max", backend: str = BACKEND, device: int | None = None, block_k: int = 0):
        super().__init__(gradient, backend, GridSolver(8, 8, device=device, block_k=block_k))

    def reset(self, src, mask, tgt) -> int:
        """``src``, ``tgt``: uint8 ``[B, rows, cols, 3]``; ``mask``: uint8 ``[B, rows, cols]``;
        the mask of edit ``b`` applies at offset (0, 0) of ``src[b]`` and ``tgt[b]``."""
        self.core.reset_batch(src, mask, tgt, self.gradient)
        return int(np.prod(np.asarray(src).shape[:3]))

    def step(self, iteration: int):
        return self.core.step(iteration)

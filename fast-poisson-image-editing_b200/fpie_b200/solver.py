"""Core solvers with the reference's interface.

``EquSolver`` / ``GridSolver`` mirror the classes every fpie backend exports
(pybind11 classes in fpie/core/cuda/solver.cc:3-15, Python classes in
fpie/np_solver.py): ``partition`` (Equ only), ``reset``, ``sync``,
``step(iteration) -> (uint8 image, float32 err[3])``.  All compute happens in
hand-written sm_100a CUDA behind the C ABI of ``include/fpie_b200.h``; numpy
arrays are only the boundary format.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

GRAD_CODE = {"src": 0, "avg": 1, "max": 2}


def _ptr(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _mask_view(mask) -> np.ndarray:
    """int32 2-D view whose rows are contiguous (the Processor hands over a
    non-contiguous slice, fpie/process.py:224/351; rows stay strided)."""
    m = np.asarray(mask)
    if m.ndim != 2:
        raise ValueError("mask must be a 2-D array")
    if m.dtype != np.int32 or m.strides[1] != 4 or m.strides[0] % 4 != 0 or m.strides[0] < 4 * m.shape[1]:
        m = np.ascontiguousarray(m, dtype=np.int32)
    return m


def _as_u8(arr, what: str) -> np.ndarray:
    """uint8 passes through; any other dtype must already hold whole numbers in 0..255 (the reference
    does its arithmetic in float32 on whatever it is given, fpie/process.py:227-246 -- values that a
    uint8 cannot hold would silently wrap here, so they are rejected instead)."""
    a = np.asarray(arr)
    if a.dtype != np.uint8:
        if a.dtype == np.bool_ or not (np.issubdtype(a.dtype, np.integer) or np.issubdtype(a.dtype, np.floating)):
            raise ValueError(f"{what} must be a uint8 array (got {a.dtype}); scale boolean / normalised data to 0..255")
        if a.size and (a.min() < 0 or a.max() > 255 or (np.issubdtype(a.dtype, np.floating) and np.any(a != np.rint(a)))):
            raise ValueError(f"{what} must hold whole numbers in 0..255 to be taken as uint8 (got {a.dtype} outside that set)")
    return np.ascontiguousarray(a, dtype=np.uint8)


def _as_u8_image(img, what: str) -> np.ndarray:
    a = _as_u8(img, what)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"{what} must be a uint8 [rows, cols, 3] image")
    return a


def _as_u8_mask(mask) -> np.ndarray:
    a = _as_u8(mask, "mask")
    if a.ndim == 2:
        return a[:, :, None]
    if a.ndim == 3 and 1 <= a.shape[2] <= 16:
        return a  # any channel count: the reference thresholds mask.mean(-1) (fpie/process.py:209-211)
    raise ValueError("mask must be uint8 [rows, cols] or [rows, cols, channels]")


def default_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except ImportError:
        pass
    return 0


def _box_callback(fn):
    """``fn(box tuple)`` as a C callback (a NULL function pointer for ``None``)."""
    if fn is None:
        return _lib.BOX_FN()
    return _lib.BOX_FN(lambda user, p: fn((int(p[0]), int(p[1]), int(p[2]), int(p[3]))))


class _Handle:
    def __init__(self):
        self._h = ctypes.c_void_p()
        self._lib = _lib.load()

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("solver has been closed")
        return self._h


class GridSolver(_Handle):
    """Drop-in for ``core_cuda.GridSolver(grid_x, grid_y)`` (fpie/process.py:312-313).

    ``grid_x`` / ``grid_y`` are accepted for signature compatibility; the tile
    geometry of the temporally blocked kernel is fixed by the register layout.
    ``block_k`` = sweeps fused per pass over HBM (default 8).
    """

    def __init__(self, grid_x: int = 8, grid_y: int = 8, device: int | None = None, block_k: int = 0,
                 variant: int = 0):
        super().__init__()
        self.grid_x, self.grid_y = int(grid_x), int(grid_y)
        self.device = default_device() if device is None else int(device)
        self.shape = None
        # the torch stream that was current at construction; the Stream object is kept alive with the solver
        # (a collected torch stream would leave the core holding a dangling cudaStream_t)
        self._stream_owner, stream = _lib.current_stream(self.device)
        self.stream_handle = stream
        _lib.check(self._lib.fpie_b200_grid_create(self.device, ctypes.c_void_p(stream), int(block_k), int(variant),
                                                   ctypes.byref(self._h)))

    def close(self) -> None:
        if self._h:
            self._lib.fpie_b200_grid_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference interface -------------------------------------------------
    def reset(self, N, mask, tgt, grad) -> None:
        """``reset(N, mask i32[n,m], tgt f32[n,m,3], grad f32[n,m,3])``; ``N`` is
        ignored exactly as in the reference (fpie/core/base_solver.h:101-113)."""
        m = _mask_view(mask)
        t = np.ascontiguousarray(tgt, dtype=np.float32)
        g = np.ascontiguousarray(grad, dtype=np.float32)
        n, w = m.shape
        if t.shape != (n, w, 3) or g.shape != (n, w, 3):
            raise ValueError("tgt and grad must have shape mask.shape + (3,)")
        _lib.check(self._lib.fpie_b200_grid_reset(self.handle, n, w, _ptr(m, ctypes.c_int32), m.strides[0] // 4, 1,
                                                  _ptr(t, ctypes.c_float), _ptr(g, ctypes.c_float)))
        self.shape = (n, w)
        self.batch_shape = None

    def sync(self) -> None:
        """No-op in the reference for every backend but mpi (base_solver.h:142)."""

    def step(self, iteration: int):
        if self.shape is None and getattr(self, "batch_shape", None):
            b, n, w = self.batch_shape
            img = np.empty((b, n, w, 3), np.uint8)
            err = np.empty((b, 3), np.float32)
        else:
            n, w = self._need_shape()
            img = np.empty((n, w, 3), np.uint8)
            err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_grid_step(self.handle, int(iteration), _ptr(img, ctypes.c_uint8),
                                                 _ptr(err, ctypes.c_float)))
        return img, err

    # -- extras ---------------------------------------------------------------
    EDGE, INTERIOR = 0, 1

    def set_edge_rows(self, rows: int) -> None:
        """Split the tile list: tiles storing any of the first / last ``rows`` grid rows are the EDGE
        part, the others the INTERIOR part (for ``pass_async``)."""
        _lib.check(self._lib.fpie_b200_grid_set_edge_rows(self.handle, int(rows)))

    def pass_async(self, nsweeps: int, part: int) -> None:
        """One pass (``1..block_k`` sweeps) over one part of the tile list, no buffer flip."""
        _lib.check(self._lib.fpie_b200_grid_pass_async(self.handle, int(nsweeps), int(part)))

    def flip(self) -> None:
        """Make the output of the pass just enqueued (both parts) the current state."""
        _lib.check(self._lib.fpie_b200_grid_flip(self.handle))

    def set_formulation(self, equ: bool) -> None:
        """Image-level resets that follow build the EquSolver's system on the grid
        (``equ=True``: X / B of process.py:227-266, zero outside the mask) instead of
        the GridSolver's -- the row-band shardable form of the EquSolver."""
        _lib.check(self._lib.fpie_b200_grid_set_formulation(self.handle, 1 if equ else 0))

    def solve(self, max_iters: int, tol: float, check_every: int = 100):
        """Sweep until every channel of ``err`` is ``<= tol`` (looked at every
        ``check_every`` sweeps) or ``max_iters`` sweeps have run; returns
        ``(err, iterations_run)``.  The state stays on the device: ``step(0)``
        returns the image.  Never implied by ``step`` -- the reference always runs
        the iteration count it is given."""
        self._need_shape()
        err = np.empty(3, np.float32)
        done = ctypes.c_int(0)
        _lib.check(self._lib.fpie_b200_grid_solve(self.handle, int(max_iters), int(check_every), float(tol),
                                                  _ptr(err, ctypes.c_float), ctypes.byref(done)))
        return err, done.value

    def step_into(self, iteration: int, canvas: np.ndarray, x0: int, y0: int) -> np.ndarray:
        """``step`` whose uint8 result lands directly in ``canvas[x0:x0+n, y0:y0+m]``
        (a C-contiguous uint8 ``[rows, cols, 3]`` image); returns ``err``."""
        n, w = self._need_shape()
        if canvas.dtype != np.uint8 or canvas.ndim != 3 or canvas.shape[2] != 3 or not canvas.flags["C_CONTIGUOUS"]:
            raise ValueError("canvas must be a C-contiguous uint8 [rows, cols, 3] image")
        if x0 < 0 or y0 < 0 or x0 + n > canvas.shape[0] or y0 + w > canvas.shape[1]:
            raise ValueError("the solved crop does not fit into the canvas")
        view = canvas[x0 : x0 + n, y0 : y0 + w]
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_grid_step_into(self.handle, int(iteration), _ptr(view, ctypes.c_uint8),
                                                      canvas.strides[0], _ptr(err, ctypes.c_float)))
        return err

    def reset_from_images(self, src, mask, tgt, mask_on_src, mask_on_tgt, gradient: str = "max"):
        """Fused ``GridProcessor.reset`` on the device (fpie/process.py:321-386).
        Returns ``(n_vars, (x0, x1, y0, y1))`` with the box in target coordinates."""
        s, t, mk = _as_u8_image(src, "src"), _as_u8_image(tgt, "tgt"), _as_u8_mask(mask)
        out_n = ctypes.c_int64()
        box = np.zeros(4, np.int32)
        _lib.check(self._lib.fpie_b200_grid_reset_from_images(
            self.handle, _ptr(s, ctypes.c_uint8), s.shape[0], s.shape[1], _ptr(mk, ctypes.c_uint8), mk.shape[0],
            mk.shape[1], mk.shape[2], _ptr(t, ctypes.c_uint8), t.shape[0], t.shape[1], int(mask_on_src[0]),
            int(mask_on_src[1]), int(mask_on_tgt[0]), int(mask_on_tgt[1]), GRAD_CODE[gradient], ctypes.byref(out_n),
            _ptr(box, ctypes.c_int32)))
        self.shape = (int(box[1] - box[0]), int(box[3] - box[2]))
        return int(out_n.value), tuple(int(v) for v in box)

    def on_box(self, fn) -> None:
        """``fn((x0, x1, y0, y1))`` is called from inside ``reset_from_images`` as soon as the blend's bounding box in
        the target is known (before the images travel); ``None`` removes it."""
        self._box_cb = _box_callback(fn)  # (kept alive for as long as the core may call it)
        _lib.check(self._lib.fpie_b200_grid_on_box(self.handle, self._box_cb, None))

    def reset_batch(self, src, mask, tgt, gradient: str = "max") -> None:
        """Batched small edits: ``src`` / ``tgt`` uint8 ``[B, rows, cols, 3]``, ``mask`` uint8
        ``[B, rows, cols]`` (or ``[B, rows, cols, 1|3]``).  Afterwards ``step`` returns
        ``(uint8 [B, rows, cols, 3], err [B, 3])``."""
        s = np.ascontiguousarray(src, dtype=np.uint8)
        t = np.ascontiguousarray(tgt, dtype=np.uint8)
        mk = np.ascontiguousarray(mask, dtype=np.uint8)
        if mk.ndim == 3:
            mk = mk[..., None]
        if s.ndim != 4 or s.shape[3] != 3 or t.shape != s.shape or mk.shape[:3] != s.shape[:3] or not 1 <= mk.shape[3] <= 16:
            raise ValueError("expected src/tgt [B, rows, cols, 3] and mask [B, rows, cols(, 1|3)]")
        b, n, w = s.shape[:3]
        _lib.check(self._lib.fpie_b200_grid_reset_batch(
            self.handle, _ptr(s, ctypes.c_uint8), _ptr(mk, ctypes.c_uint8), _ptr(t, ctypes.c_uint8), b, n, w,
            mk.shape[3], GRAD_CODE[gradient]))
        self.shape = None
        self.batch_shape = (b, n, w)

    def reset_slab(self, src, mask, tgt, gradient: str = "max") -> None:
        """Load one slab of a row-band sharded problem (see ``fpie_b200.band``): three
        uint8 images of identical size; the whole slab is the grid, its frame is fixed."""
        s, t, mk = _as_u8_image(src, "src"), _as_u8_image(tgt, "tgt"), _as_u8_mask(mask)
        if not (s.shape == t.shape and mk.shape[:2] == s.shape[:2]):
            raise ValueError("slab images must have identical rows x cols")
        _lib.check(self._lib.fpie_b200_grid_reset_slab(
            self.handle, _ptr(s, ctypes.c_uint8), _ptr(mk, ctypes.c_uint8), _ptr(t, ctypes.c_uint8), s.shape[0],
            s.shape[1], mk.shape[2], GRAD_CODE[gradient]))
        self.shape = (s.shape[0], s.shape[1])

    def state(self) -> np.ndarray:
        n, w = self._need_shape()
        out = np.empty((n, w, 3), np.float32)
        _lib.check(self._lib.fpie_b200_grid_state(self.handle, _ptr(out, ctypes.c_float)))
        return out

    def batch_state(self) -> np.ndarray:
        """fp32 state of a batched solve as ``[B, rows, cols, 3]`` (de-mosaicked on the host)."""
        if not getattr(self, "batch_shape", None):
            raise RuntimeError("batch_state needs reset_batch")
        b, n, w = self.batch_shape
        bcols = 1
        while bcols * w < 4096 and bcols < b:  # mosaic rule of GridSolver::reset_batch (csrc/grid.cu)
            bcols *= 2
        brows = -(-b // bcols)
        out = np.empty((brows * n, bcols * w, 3), np.float32)
        _lib.check(self._lib.fpie_b200_grid_state(self.handle, _ptr(out, ctypes.c_float)))
        tiles = out.reshape(brows, n, bcols, w, 3).transpose(0, 2, 1, 3, 4).reshape(brows * bcols, n, w, 3)
        return np.ascontiguousarray(tiles[:b])

    def sweeps_async(self, iteration: int) -> None:
        _lib.check(self._lib.fpie_b200_grid_sweeps_async(self.handle, int(iteration)))

    def finish_async(self) -> None:
        _lib.check(self._lib.fpie_b200_grid_finish_async(self.handle))

    def wait(self) -> None:
        _lib.check(self._lib.fpie_b200_grid_sync(self.handle))

    def fetch(self, img: np.ndarray | None = None):
        n, w = self._need_shape()
        if img is None:
            img = np.empty((n, w, 3), np.uint8)
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_grid_fetch(self.handle, _ptr(img, ctypes.c_uint8), _ptr(err, ctypes.c_float)))
        return img, err

    def fetch_rows(self, lo: int, hi: int, img: np.ndarray | None = None):
        """``fetch`` of rows ``[lo, hi)`` only (a row band's own rows, without its halo rows)."""
        n, w = self._need_shape()
        if not 0 <= lo <= hi <= n:
            raise ValueError("row range outside the grid")
        if img is None:
            img = np.empty((hi - lo, w, 3), np.uint8)
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_grid_fetch_rows(self.handle, int(lo), int(hi), _ptr(img, ctypes.c_uint8),
                                                       _ptr(err, ctypes.c_float)))
        return img, err

    def info(self) -> dict:
        unk, launches, act, tot = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        k = ctypes.c_int()
        _lib.check(self._lib.fpie_b200_grid_info(self.handle, ctypes.byref(unk), ctypes.byref(launches),
                                                 ctypes.byref(k), ctypes.byref(act), ctypes.byref(tot)))
        v, rows, warps, occ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self._lib.fpie_b200_grid_config(self.handle, ctypes.byref(v), ctypes.byref(rows), ctypes.byref(warps),
                                                   ctypes.byref(occ)))
        return dict(unknowns=unk.value, launches=launches.value, block_k=k.value, active_tiles=act.value,
                    total_tiles=tot.value, variant=v.value, tile=(rows.value * warps.value, 128),
                    rows_per_thread=rows.value, warps=warps.value, ctas_per_sm=occ.value)

    def patch_info(self) -> dict:
        """The persistent small-image kernel for the current problem: usable?, thread tile, cluster size, launches."""
        u, r, c, cl = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        n = ctypes.c_int64()
        _lib.check(self._lib.fpie_b200_grid_patch_info(self.handle, ctypes.byref(u), ctypes.byref(r), ctypes.byref(c),
                                                       ctypes.byref(cl), ctypes.byref(n)))
        return dict(usable=bool(u.value), rows_per_thread=r.value, cols_per_thread=c.value, cluster=cl.value,
                    launches=n.value & ((1 << 40) - 1), clusters=n.value >> 40)

    def set_row_window(self, lo: int, hi: int) -> None:
        _lib.check(self._lib.fpie_b200_grid_set_row_window(self.handle, int(lo), int(hi)))

    def band_view(self, which: int):
        base = ctypes.c_void_p()
        plane, pitch = ctypes.c_int64(), ctypes.c_int64()
        padr, padc = ctypes.c_int(), ctypes.c_int()
        _lib.check(self._lib.fpie_b200_grid_band_view(self.handle, int(which), ctypes.byref(base), ctypes.byref(plane),
                                                      ctypes.byref(pitch), ctypes.byref(padr), ctypes.byref(padc)))
        return dict(base=base.value, plane=plane.value, pitch=pitch.value, pad_rows=padr.value, pad_cols=padc.value)

    # -- halo link (row bands; the exchange runs behind the C ABI, csrc/halo.cu) ---------------
    HALO_BLOB_BYTES = 128
    UP, DOWN = 0, 1

    def halo_config(self, band_lo: int, band_hi: int, force: bool = False) -> bool:
        """This slab's own rows are ``[band_lo, band_hi)``; returns True when the link was rebuilt
        (then ``halo_export`` / ``halo_connect`` are due on both ends)."""
        changed = ctypes.c_int()
        _lib.check(self._lib.fpie_b200_grid_halo_config(self.handle, int(band_lo), int(band_hi), int(bool(force)),
                                                        ctypes.byref(changed)))
        return bool(changed.value)

    def halo_export(self, side: int) -> bytes:
        blob = np.zeros(self.HALO_BLOB_BYTES, np.uint8)
        _lib.check(self._lib.fpie_b200_grid_halo_export(self.handle, int(side), _ptr(blob, ctypes.c_uint8)))
        return blob.tobytes()

    def halo_connect(self, side: int, blob: bytes, same_process: bool) -> None:
        buf = np.frombuffer(bytes(blob), np.uint8).copy()
        if buf.size != self.HALO_BLOB_BYTES:
            raise ValueError("not a halo blob")
        _lib.check(self._lib.fpie_b200_grid_halo_connect(self.handle, int(side), _ptr(buf, ctypes.c_uint8),
                                                         int(bool(same_process))))

    def band_sweeps_async(self, iteration: int) -> None:
        _lib.check(self._lib.fpie_b200_grid_band_sweeps_async(self.handle, int(iteration)))

    def halo_exchanges(self) -> int:
        n = ctypes.c_int64()
        _lib.check(self._lib.fpie_b200_grid_halo_stats(self.handle, ctypes.byref(n)))
        return n.value

    def halo_debug(self) -> dict:
        """Counters of the halo link (safe to call from another thread while a step is blocked)."""
        v = (ctypes.c_int64 * 16)()
        _lib.check(self._lib.fpie_b200_grid_halo_debug(self.handle, v))
        v = list(v)
        return dict(rows=v[0:2], sent=v[2:4], received=v[4:6], flags_up=v[6:8], flags_down=v[8:10], current=v[10],
                    block_k=v[11], variant=v[12], edge_tiles=v[13], interior_tiles=v[14], pending=v[15])

    def halo_trace_begin(self, max_intervals: int) -> None:
        _lib.check(self._lib.fpie_b200_grid_halo_trace_begin(self.handle, int(max_intervals)))

    def halo_trace(self):
        """``[(tag, ms), ...]`` of the traced intervals (see include/fpie_b200.h); synchronises."""
        out = np.zeros(8192, np.float32)
        n = ctypes.c_int()
        _lib.check(self._lib.fpie_b200_grid_halo_trace_read(self.handle, _ptr(out, ctypes.c_float), out.size,
                                                            ctypes.byref(n)))
        return [(int(out[i]), float(out[i + 1])) for i in range(0, n.value, 2)]

    def current_buffer(self) -> int:
        which = ctypes.c_int()
        _lib.check(self._lib.fpie_b200_grid_band_current(self.handle, ctypes.byref(which)))
        return which.value

    def _need_shape(self):
        if self.shape is None:
            raise RuntimeError("GridSolver: step/state called before reset")
        return self.shape


class EquSolver(_Handle):
    """Drop-in for ``core_cuda.EquSolver(block_size)`` (fpie/process.py:173-174)."""

    MODES = {"jacobi": 0, "redblack": 1, "gather": 2}
    PATHS = {0: "gather-int4", 1: "gather-compact", 2: "tiled", 3: "redblack"}

    def __init__(self, block_size: int = 256, device: int | None = None, mode: str = "jacobi"):
        """``mode="jacobi"`` (default): true Jacobi; when the system is provably the 4-neighbour structure
        of a mask this solver labelled (``partition`` / ``reset_from_images``) it runs on the temporally
        blocked grid kernel (same bits, several times faster), otherwise on the gather kernels.
        ``mode="gather"``: Jacobi on the index-mapped gather kernels only.
        ``mode="redblack"``: the reference OpenMP backend's red-black Gauss-Seidel
        (fpie/core/openmp/equ.cc:22-56, 107-118); ``partition`` then labels odd pixels before even
        ones, exactly as ``core_openmp.EquSolver.partition`` does."""
        super().__init__()
        if mode not in self.MODES:
            raise ValueError(f"mode must be one of {sorted(self.MODES)}")
        self.device = default_device() if device is None else int(device)
        self.mode = mode
        self.N = 0
        self.crop_shape = None
        self._stream_owner, stream = _lib.current_stream(self.device)
        self.stream_handle = stream
        _lib.check(self._lib.fpie_b200_equ_create(self.device, ctypes.c_void_p(stream), int(block_size),
                                                  ctypes.byref(self._h)))
        if mode != "jacobi":
            _lib.check(self._lib.fpie_b200_equ_set_mode(self.handle, self.MODES[mode]))

    def close(self) -> None:
        if self._h:
            self._lib.fpie_b200_equ_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference interface -------------------------------------------------
    def partition(self, mask) -> np.ndarray:
        """Row-major ids of masked pixels by a device prefix scan
        (fpie/core/cuda/equ.cu:36-54; fpie/np_solver.py:14-16)."""
        m = _mask_view(mask)
        ids = np.empty(m.shape, np.int32)
        _lib.check(self._lib.fpie_b200_equ_partition(self.handle, m.shape[0], m.shape[1], _ptr(m, ctypes.c_int32),
                                                     m.strides[0] // 4, 1, _ptr(ids, ctypes.c_int32)))
        return ids

    def reset(self, N, A, X, B) -> None:
        N = int(N)
        a = np.ascontiguousarray(A, dtype=np.int32)
        x = np.ascontiguousarray(X, dtype=np.float32)
        b = np.ascontiguousarray(B, dtype=np.float32)
        if a.shape != (N, 4) or x.shape != (N, 3) or b.shape != (N, 3):
            raise ValueError("expected A[N,4], X[N,3], B[N,3]")
        _lib.check(self._lib.fpie_b200_equ_reset(self.handle, N, _ptr(a, ctypes.c_int32), _ptr(x, ctypes.c_float),
                                                 _ptr(b, ctypes.c_float)))
        self.N = N
        self.crop_shape = None

    def sync(self) -> None:
        """No-op, as in the reference (base_solver.h:61)."""

    def step(self, iteration: int):
        self._need_reset()
        img = np.empty((self.N, 3), np.uint8)
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_equ_step(self.handle, int(iteration), _ptr(img, ctypes.c_uint8),
                                                _ptr(err, ctypes.c_float)))
        return img, err

    # -- extras ---------------------------------------------------------------
    def reset_from_images(self, src, mask, tgt, mask_on_src, mask_on_tgt, gradient: str = "max"):
        """Fused ``EquProcessor.reset`` on the device (fpie/process.py:192-271).
        Returns ``(N, (x0, x1, y0, y1))`` with the box in target coordinates."""
        s, t, mk = _as_u8_image(src, "src"), _as_u8_image(tgt, "tgt"), _as_u8_mask(mask)
        out_n = ctypes.c_int64()
        box = np.zeros(4, np.int32)
        _lib.check(self._lib.fpie_b200_equ_reset_from_images(
            self.handle, _ptr(s, ctypes.c_uint8), s.shape[0], s.shape[1], _ptr(mk, ctypes.c_uint8), mk.shape[0],
            mk.shape[1], mk.shape[2], _ptr(t, ctypes.c_uint8), t.shape[0], t.shape[1], int(mask_on_src[0]),
            int(mask_on_src[1]), int(mask_on_tgt[0]), int(mask_on_tgt[1]), GRAD_CODE[gradient], ctypes.byref(out_n),
            _ptr(box, ctypes.c_int32)))
        self.N = int(out_n.value)
        self.crop_shape = (int(box[1] - box[0]), int(box[3] - box[2]))
        return self.N, tuple(int(v) for v in box)

    def on_box(self, fn) -> None:
        """As ``GridSolver.on_box``."""
        self._box_cb = _box_callback(fn)
        _lib.check(self._lib.fpie_b200_equ_on_box(self.handle, self._box_cb, None))

    def step_paste(self, iteration: int):
        """``step`` + the Processor's scatter (process.py:273-280), on the device:
        returns the uint8 crop ``[x1-x0, y1-y0, 3]`` and ``err``."""
        if self.crop_shape is None:
            raise RuntimeError("step_paste needs reset_from_images")
        crop = np.empty(self.crop_shape + (3,), np.uint8)
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_equ_step_paste(self.handle, int(iteration), _ptr(crop, ctypes.c_uint8),
                                                      _ptr(err, ctypes.c_float)))
        return crop, err

    def solve(self, max_iters: int, tol: float, check_every: int = 100):
        """As ``GridSolver.solve``: sweep until ``max(err) <= tol`` or ``max_iters``;
        returns ``(err, iterations_run)``."""
        err = np.empty(3, np.float32)
        done = ctypes.c_int(0)
        _lib.check(self._lib.fpie_b200_equ_solve(self.handle, int(max_iters), int(check_every), float(tol),
                                                 _ptr(err, ctypes.c_float), ctypes.byref(done)))
        return err, done.value

    def step_paste_into(self, iteration: int, canvas: np.ndarray, x0: int, y0: int) -> np.ndarray:
        """``step_paste`` landing directly in ``canvas[x0:x0+n, y0:y0+m]`` (C-contiguous uint8 image)."""
        if self.crop_shape is None:
            raise RuntimeError("step_paste_into needs reset_from_images")
        n, w = self.crop_shape
        if canvas.dtype != np.uint8 or canvas.ndim != 3 or canvas.shape[2] != 3 or not canvas.flags["C_CONTIGUOUS"]:
            raise ValueError("canvas must be a C-contiguous uint8 [rows, cols, 3] image")
        if x0 < 0 or y0 < 0 or x0 + n > canvas.shape[0] or y0 + w > canvas.shape[1]:
            raise ValueError("the solved crop does not fit into the canvas")
        view = canvas[x0 : x0 + n, y0 : y0 + w]
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_equ_step_paste_into(self.handle, int(iteration), _ptr(view, ctypes.c_uint8),
                                                           canvas.strides[0], _ptr(err, ctypes.c_float)))
        return err

    def state(self) -> np.ndarray:
        self._need_reset()
        out = np.empty((self.N, 3), np.float32)
        _lib.check(self._lib.fpie_b200_equ_state(self.handle, _ptr(out, ctypes.c_float)))
        return out

    def system(self):
        """(A, X, B) as the device holds them (parity checks of the fused reset)."""
        self._need_reset()
        A = np.empty((self.N, 4), np.int32)
        X = np.empty((self.N, 3), np.float32)
        B = np.empty((self.N, 3), np.float32)
        _lib.check(self._lib.fpie_b200_equ_system(self.handle, _ptr(A, ctypes.c_int32), _ptr(X, ctypes.c_float),
                                                  _ptr(B, ctypes.c_float)))
        return A, X, B

    def sweeps_async(self, iteration: int) -> None:
        _lib.check(self._lib.fpie_b200_equ_sweeps_async(self.handle, int(iteration)))

    def finish_async(self) -> None:
        _lib.check(self._lib.fpie_b200_equ_finish_async(self.handle))

    def wait(self) -> None:
        _lib.check(self._lib.fpie_b200_equ_sync(self.handle))

    def fetch(self, img: np.ndarray | None = None):
        self._need_reset()
        if img is None:
            img = np.empty((self.N, 3), np.uint8)
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_equ_fetch(self.handle, _ptr(img, ctypes.c_uint8), _ptr(err, ctypes.c_float)))
        return img, err

    # -- id-range sharding (fpie_b200/shard.py) ------------------------------------
    def set_window(self, lo: int, hi: int) -> None:
        """The residual of ``finish`` / ``step`` sums rows ``[lo, hi)`` only."""
        _lib.check(self._lib.fpie_b200_equ_set_window(self.handle, int(lo), int(hi)))

    def fetch_rows(self, lo: int, hi: int, img: np.ndarray | None = None):
        """uint8 rows ``[lo, hi)`` of the last ``finish_async`` and ``err``."""
        self._need_reset()
        if img is None:
            img = np.empty((hi - lo, 3), np.uint8)
        err = np.empty(3, np.float32)
        _lib.check(self._lib.fpie_b200_equ_fetch_rows(self.handle, int(lo), int(hi), _ptr(img, ctypes.c_uint8),
                                                      _ptr(err, ctypes.c_float)))
        return img, err

    def gather_rows(self, idx_ptr: int, n: int, out_ptr: int) -> None:
        """``out[j, :] = X[idx[j], :]`` for DEVICE buffers (``idx`` int32, ``out`` fp32 ``[n, 3]``), on the solver's stream."""
        _lib.check(self._lib.fpie_b200_equ_gather_rows(self.handle, int(idx_ptr), int(n), int(out_ptr)))

    def scatter_rows(self, idx_ptr: int, n: int, in_ptr: int) -> None:
        """``X[idx[j], :] = in[j, :]`` for DEVICE buffers, on the solver's stream."""
        _lib.check(self._lib.fpie_b200_equ_scatter_rows(self.handle, int(idx_ptr), int(n), int(in_ptr)))

    def rows_checked(self, on: bool = True) -> None:
        """The index lists of ``gather_rows`` / ``scatter_rows`` were validated: skip the per-call read-back."""
        _lib.check(self._lib.fpie_b200_equ_rows_checked(self.handle, int(bool(on))))

    def info(self) -> dict:
        unk, launches, path = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        _lib.check(self._lib.fpie_b200_equ_info(self.handle, ctypes.byref(unk), ctypes.byref(launches),
                                                ctypes.byref(path)))
        code = path.value
        # gather-compact with its table / B-stream form in `table`: the path NAME stays what callers test for
        table = {1: "int2 + fp32 B", 9: "delta16 + fp32 B", 25: "delta16 + fp16 B"}.get(code)
        return dict(unknowns=unk.value, launches=launches.value, path=self.PATHS[code & 7], table=table)

    def _need_reset(self):
        if self.N <= 0:
            raise RuntimeError("EquSolver: step/state called before reset")


def device_count() -> int:
    return int(_lib.load().fpie_b200_device_count())


def device_info(device: int = 0) -> dict:
    lib = _lib.load()
    name = ctypes.create_string_buffer(256)
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.fpie_b200_device_info(int(device), name, 256, ctypes.byref(sm), ctypes.byref(major),
                                         ctypes.byref(minor)))
    return dict(name=name.value.decode(), sm_count=sm.value, cc=(major.value, minor.value))

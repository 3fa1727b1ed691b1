"""ctypes binding of ``libfpie_b200.so`` (C ABI: include/fpie_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be
loaded, importing a solver raises ``ImportError`` -- the same signal the
reference uses to drop a backend (fpie/process.py:77-83).
"""

from __future__ import annotations

import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# FPIE_B200_LIB points tools/ at an experimental build of the same library (never a different implementation)
LIB_PATH = os.environ.get("FPIE_B200_LIB") or os.path.join(PKG_DIR, "libfpie_b200.so")

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_void_p = ctypes.c_void_p
P = ctypes.POINTER
i32p, f32p, u8p, i64p, intp = P(ctypes.c_int32), P(ctypes.c_float), P(ctypes.c_uint8), P(c_i64), P(c_int)

BOX_FN = ctypes.CFUNCTYPE(None, c_void_p, P(ctypes.c_int32))  # fpie_b200_box_fn

# name -> (argtypes); every function returns int except the two noted below.
SIGNATURES = {
    "fpie_b200_abi_version": [],
    "fpie_b200_device_count": [],
    "fpie_b200_host_alloc": [c_i64, ctypes.POINTER(c_void_p)],
    "fpie_b200_host_free": [c_void_p],
    "fpie_b200_device_info": [c_int, ctypes.c_char_p, c_int, intp, intp, intp],
    "fpie_b200_grid_create": [c_int, c_void_p, c_int, c_int, P(c_void_p)],
    "fpie_b200_grid_destroy": [c_void_p],
    "fpie_b200_grid_reset": [c_void_p, c_int, c_int, i32p, c_i64, c_i64, f32p, f32p],
    "fpie_b200_grid_step": [c_void_p, c_int, u8p, f32p],
    "fpie_b200_grid_step_into": [c_void_p, c_int, u8p, c_i64, f32p],
    "fpie_b200_grid_config": [c_void_p, intp, intp, intp, intp],
    "fpie_b200_grid_set_formulation": [c_void_p, c_int],
    "fpie_b200_grid_set_edge_rows": [c_void_p, c_int],
    "fpie_b200_grid_pass_async": [c_void_p, c_int, c_int],
    "fpie_b200_grid_flip": [c_void_p],
    "fpie_b200_grid_solve": [c_void_p, c_int, c_int, ctypes.c_float, f32p, intp],
    "fpie_b200_grid_state": [c_void_p, f32p],
    "fpie_b200_grid_sweeps_async": [c_void_p, c_int],
    "fpie_b200_grid_finish_async": [c_void_p],
    "fpie_b200_grid_sync": [c_void_p],
    "fpie_b200_grid_fetch": [c_void_p, u8p, f32p],
    "fpie_b200_grid_fetch_rows": [c_void_p, c_int, c_int, u8p, f32p],
    "fpie_b200_grid_patch_info": [c_void_p, intp, intp, intp, intp, i64p],
    "fpie_b200_grid_halo_config": [c_void_p, c_int, c_int, c_int, intp],
    "fpie_b200_grid_halo_export": [c_void_p, c_int, u8p],
    "fpie_b200_grid_halo_connect": [c_void_p, c_int, u8p, c_int],
    "fpie_b200_grid_band_sweeps_async": [c_void_p, c_int],
    "fpie_b200_grid_halo_stats": [c_void_p, i64p],
    "fpie_b200_grid_halo_debug": [c_void_p, i64p],
    "fpie_b200_grid_halo_trace_begin": [c_void_p, c_int],
    "fpie_b200_grid_halo_trace_read": [c_void_p, f32p, c_int, intp],
    "fpie_b200_grid_info": [c_void_p, i64p, i64p, intp, i64p, i64p],
    "fpie_b200_grid_reset_from_images": [c_void_p, u8p, c_int, c_int, u8p, c_int, c_int, c_int, u8p, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, i64p, i32p],
    "fpie_b200_grid_reset_batch": [c_void_p, u8p, u8p, u8p, c_int, c_int, c_int, c_int, c_int],
    "fpie_b200_grid_reset_slab": [c_void_p, u8p, u8p, u8p, c_int, c_int, c_int, c_int],
    "fpie_b200_grid_band_view": [c_void_p, c_int, P(c_void_p), i64p, i64p, intp, intp],
    "fpie_b200_grid_band_current": [c_void_p, intp],
    "fpie_b200_grid_set_row_window": [c_void_p, c_int, c_int],
    "fpie_b200_equ_create": [c_int, c_void_p, c_int, P(c_void_p)],
    "fpie_b200_equ_destroy": [c_void_p],
    "fpie_b200_equ_set_mode": [c_void_p, c_int],
    "fpie_b200_equ_partition": [c_void_p, c_int, c_int, i32p, c_i64, c_i64, i32p],
    "fpie_b200_equ_reset": [c_void_p, c_i64, i32p, f32p, f32p],
    "fpie_b200_equ_step": [c_void_p, c_int, u8p, f32p],
    "fpie_b200_equ_solve": [c_void_p, c_int, c_int, ctypes.c_float, f32p, intp],
    "fpie_b200_equ_state": [c_void_p, f32p],
    "fpie_b200_equ_sweeps_async": [c_void_p, c_int],
    "fpie_b200_equ_finish_async": [c_void_p],
    "fpie_b200_equ_sync": [c_void_p],
    "fpie_b200_equ_fetch": [c_void_p, u8p, f32p],
    "fpie_b200_equ_info": [c_void_p, i64p, i64p, intp],
    "fpie_b200_equ_reset_from_images": [c_void_p, u8p, c_int, c_int, u8p, c_int, c_int, c_int, u8p, c_int, c_int,
                                        c_int, c_int, c_int, c_int, c_int, i64p, i32p],
    "fpie_b200_equ_step_paste": [c_void_p, c_int, u8p, f32p],
    "fpie_b200_equ_step_paste_into": [c_void_p, c_int, u8p, c_i64, f32p],
    "fpie_b200_equ_system": [c_void_p, i32p, f32p, f32p],
    "fpie_b200_grid_on_box": [c_void_p, BOX_FN, c_void_p],
    "fpie_b200_equ_on_box": [c_void_p, BOX_FN, c_void_p],
    "fpie_b200_equ_set_window": [c_void_p, c_i64, c_i64],
    "fpie_b200_equ_fetch_rows": [c_void_p, c_i64, c_i64, u8p, f32p],
    "fpie_b200_equ_gather_rows": [c_void_p, c_void_p, c_i64, c_void_p],
    "fpie_b200_equ_scatter_rows": [c_void_p, c_void_p, c_i64, c_void_p],
    "fpie_b200_equ_rows_checked": [c_void_p, c_int],
}

_lib = None


def load():
    """Load the shared library once; raise ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m fpie_b200._build` "
            "(needs nvcc; fpie_b200 has no CPU fallback)"
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # e.g. libcudart not found
        raise ImportError(f"cannot load {LIB_PATH}: {exc}") from exc
    lib.fpie_b200_last_error.restype = ctypes.c_char_p
    lib.fpie_b200_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Turn a non-zero status into RuntimeError, as pybind11 does for the
    reference's std::runtime_error (fpie/core/base_solver.h:63-74)."""
    if rc != 0:
        msg = load().fpie_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(msg or f"fpie_b200 call failed with status {rc}")


class _PinnedBlock:
    """Page-locked host bytes owned by the C ABI (``fpie_b200_host_alloc``).  numpy arrays made from it
    keep it as their ``base``; the memory is released when the last of them is gone."""

    def __init__(self, nbytes: int):
        self._lib = load()
        self._ptr = c_void_p()
        check(self._lib.fpie_b200_host_alloc(int(nbytes), ctypes.byref(self._ptr)))
        self.__array_interface__ = {"shape": (max(int(nbytes), 1),), "typestr": "|u1",
                                    "data": (self._ptr.value, False), "version": 3}

    def __del__(self):
        try:
            if self._ptr:
                self._lib.fpie_b200_host_free(self._ptr)
                self._ptr = c_void_p()
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """``np.empty(shape, dtype)`` in page-locked host memory."""
    import numpy as np

    shape = tuple(int(v) for v in shape)
    count = int(np.prod(shape, dtype=np.int64))
    nbytes = count * np.dtype(dtype).itemsize
    raw = np.asarray(_PinnedBlock(nbytes))
    return raw[:nbytes].view(dtype).reshape(shape)


def current_stream(device: int):
    """``(torch Stream object or None, raw cudaStream_t)`` of torch's current stream on ``device``
    (``(None, 0)`` = the default stream when torch is not importable).  Whoever hands the raw handle
    to the C ABI must keep the Stream object alive for as long as the handle is in use."""
    try:
        import torch
    except ImportError:
        return None, 0
    if not torch.cuda.is_available():
        return None, 0
    stream = torch.cuda.current_stream(device)
    return stream, int(stream.cuda_stream)


def current_stream_handle(device: int) -> int:
    return current_stream(device)[1]

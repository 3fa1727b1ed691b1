"""ctypes front end for ``oracle/libjacobi_oracle.so`` (the C restatement) and
loader for the compiled reference cores under ``oracle/_ref``.

TEST INFRASTRUCTURE ONLY -- see the header of ``oracle/np_oracle.py``.
"""

from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjacobi_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (and, when the reference checkout is present,
    the reference's own cores into ``oracle/_ref``)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
        os.path.join(HERE, "jacobi_oracle.c")
    ):
        subprocess.check_call(["make", "-C", HERE, "-B", "libjacobi_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def build_reference(cuda: bool = False) -> bool:
    """Compile the unmodified reference cores from /root/reference, if present."""
    ref = os.environ.get("FPIE_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "fpie", "core")):
        return False
    targets = ["ref"] + (["ref-cuda"] if cuda else [])
    subprocess.check_call(["make", "-C", HERE, f"REF={ref}", *targets], stdout=subprocess.DEVNULL)
    return True


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        i64, i32p, f32p, f64p, u8p, ci = (
            ctypes.c_int64,
            ctypes.POINTER(ctypes.c_int32),
            ctypes.POINTER(ctypes.c_float),
            ctypes.POINTER(ctypes.c_double),
            ctypes.POINTER(ctypes.c_uint8),
            ctypes.c_int,
        )
        L.oracle_max_threads.restype = ci
        L.oracle_grid_sweeps.argtypes = [i64, i64, i32p, f32p, f32p, f32p, ci, ci]
        L.oracle_grid_residual.argtypes = [i64, i64, i32p, f32p, f32p, f32p, f64p]
        L.oracle_equ_sweeps.argtypes = [i64, i32p, f32p, f32p, f32p, ci, ci]
        L.oracle_equ_residual.argtypes = [i64, i32p, f32p, f32p, f32p, f64p]
        L.oracle_clip_u8.argtypes = [i64, f32p, u8p]
        for fn in ("oracle_grid_sweeps", "oracle_grid_residual", "oracle_equ_sweeps", "oracle_equ_residual", "oracle_clip_u8"):
            getattr(L, fn).restype = None
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def grid_sweeps(mask, tgt, grad, iters: int, threads: int = 0) -> np.ndarray:
    """Return the fp32 grid state after ``iters`` Jacobi sweeps (inputs untouched)."""
    mask = np.ascontiguousarray(mask, np.int32)
    t = np.array(tgt, np.float32, order="C", copy=True)
    g = np.ascontiguousarray(grad, np.float32)
    scratch = np.empty_like(t)
    n, m = mask.shape
    lib().oracle_grid_sweeps(n, m, _p(mask, ctypes.c_int32), _p(t, ctypes.c_float), _p(g, ctypes.c_float),
                             _p(scratch, ctypes.c_float), int(iters), int(threads))
    return t


def grid_residual(mask, tgt, grad):
    """Return ``(err_f32_sequential[3], err_f64[3])``."""
    mask = np.ascontiguousarray(mask, np.int32)
    t = np.ascontiguousarray(tgt, np.float32)
    g = np.ascontiguousarray(grad, np.float32)
    e32 = np.zeros(3, np.float32)
    e64 = np.zeros(3, np.float64)
    n, m = mask.shape
    lib().oracle_grid_residual(n, m, _p(mask, ctypes.c_int32), _p(t, ctypes.c_float), _p(g, ctypes.c_float),
                               _p(e32, ctypes.c_float), _p(e64, ctypes.c_double))
    return e32, e64


def equ_sweeps(A, X, B, iters: int, threads: int = 0) -> np.ndarray:
    A = np.ascontiguousarray(A, np.int32)
    x = np.array(X, np.float32, order="C", copy=True)
    b = np.ascontiguousarray(B, np.float32)
    scratch = np.empty_like(x)
    lib().oracle_equ_sweeps(A.shape[0], _p(A, ctypes.c_int32), _p(x, ctypes.c_float), _p(b, ctypes.c_float),
                            _p(scratch, ctypes.c_float), int(iters), int(threads))
    return x


def equ_residual(A, X, B):
    A = np.ascontiguousarray(A, np.int32)
    x = np.ascontiguousarray(X, np.float32)
    b = np.ascontiguousarray(B, np.float32)
    e32 = np.zeros(3, np.float32)
    e64 = np.zeros(3, np.float64)
    lib().oracle_equ_residual(A.shape[0], _p(A, ctypes.c_int32), _p(x, ctypes.c_float), _p(b, ctypes.c_float),
                              _p(e32, ctypes.c_float), _p(e64, ctypes.c_double))
    return e32, e64


def clip_u8(state) -> np.ndarray:
    s = np.ascontiguousarray(state, np.float32)
    out = np.empty(s.shape, np.uint8)
    lib().oracle_clip_u8(s.size, _p(s, ctypes.c_float), _p(out, ctypes.c_uint8))
    return out


def load_reference_core(name: str):
    """Import ``core_openmp`` / ``core_gcc`` / ``core_cuda`` from ``oracle/_ref``
    (binaries compiled from the unmodified reference).  Returns None if absent."""
    if not os.path.isdir(REF_DIR):
        return None
    for fn in os.listdir(REF_DIR):
        if fn.startswith(name + ".") and fn.endswith(".so"):
            if name in sys.modules:
                return sys.modules[name]
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, fn))
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
            except ImportError:
                return None
            sys.modules[name] = mod
            return mod
    return None

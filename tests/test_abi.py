"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/fpie_b200.h declares, and refuses to run without a GPU
(no silent CPU fallback)."""

import ctypes
import os
import re

import pytest
from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "fpie_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpie_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for required in (
        "fpie_b200_grid_create", "fpie_b200_grid_reset", "fpie_b200_grid_step", "fpie_b200_grid_state",
        "fpie_b200_grid_destroy", "fpie_b200_equ_create", "fpie_b200_equ_partition", "fpie_b200_equ_reset",
        "fpie_b200_equ_step", "fpie_b200_equ_state", "fpie_b200_equ_destroy", "fpie_b200_last_error",
    ):
        assert required in syms


def test_library_exports_every_declared_symbol():
    from fpie_b200 import _lib

    lib = _lib.load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/fpie_b200.h but not exported: {missing}"
    # and the ctypes table binds all of them
    unbound = [s for s in declared_symbols() if s not in _lib.SIGNATURES and s != "fpie_b200_last_error"]
    assert not unbound, f"no ctypes signature for: {unbound}"
    assert lib.fpie_b200_abi_version() == 1


def test_only_declared_symbols_are_exported():
    from fpie_b200 import _lib
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    extra = {s for s in exported if not s.startswith("fpie_b200_") and not s.startswith("_")}
    assert not extra, f"unexpected exported symbols: {sorted(extra)}"
    assert set(declared_symbols()) <= exported


def _has_gpu():
    from fpie_b200 import _lib

    return _lib.load().fpie_b200_device_count() > 0


def test_no_cpu_fallback_without_gpu():
    if _has_gpu():
        pytest.skip("a GPU is present")
    import fpie_b200

    with pytest.raises(RuntimeError):
        fpie_b200.GridSolver(8, 8)
    with pytest.raises(RuntimeError):
        fpie_b200.EquSolver(256)
    with pytest.raises(RuntimeError):
        fpie_b200.GridProcessor("max", "b200")


def test_null_handle_is_an_error_not_a_crash():
    from fpie_b200 import _lib

    lib = _lib.load()
    rc = lib.fpie_b200_grid_step(None, 1, None, None)
    assert rc != 0 and b"null" in lib.fpie_b200_last_error()
    rc = lib.fpie_b200_equ_step(None, 1, None, None)
    assert rc != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fast-poisson-image-editing_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "oracle" not in text.lower() or f == "_build.py", f"{f} mentions the oracle"

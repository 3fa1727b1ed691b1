"""Builds ``libfpie_b200.so`` in-tree with nvcc for sm_100a.

``python -m fpie_b200._build`` (or ``__graft_entry__.build()``).  The library
is a plain C-ABI shared object (include/fpie_b200.h); it links only cudart.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DIST_ROOT = os.path.dirname(PKG_DIR)
REPO_ROOT = os.path.dirname(DIST_ROOT)
CSRC = os.path.join(DIST_ROOT, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_PATH = os.path.join(PKG_DIR, "libfpie_b200.so")
OBJ_DIR = os.path.join(DIST_ROOT, "build")

SOURCES = ("api.cu", "grid.cu", "halo.cu", "equ.cu", "prep.cu")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: every fused multiply-add in the kernels is an explicit intrinsic;
# nothing else may be contracted, or fp32 results drift from the reference's.
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: fpie_b200 has no CPU fallback and cannot be built without the CUDA toolkit")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), lib_path: str = LIB_PATH, obj_dir: str = OBJ_DIR) -> str:
    # FPIE_B200_ALL_VARIANTS=1: also compile the measured-and-dominated tile shapes (tuning / full test builds)
    if os.environ.get("FPIE_B200_ALL_VARIANTS", "") not in ("", "0") and "FPIE_ALL_VARIANTS" not in defines:
        defines = (*defines, "FPIE_ALL_VARIANTS")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "fpie_b200.h"))
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path, *headers]):
            cmd = [nvcc, *ARCH, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", path, "-o", obj]
            log = open(obj + ".log", "w")
            procs.append((src, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log))
    failed = []
    for src, p, log in procs:
        rc = p.wait()
        log.close()
        text = open(log.name).read()
        if verbose or rc != 0:
            sys.stderr.write(f"--- nvcc {src} (exit {rc})\n{text}\n")
        if rc != 0:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    if force or procs or _stale(lib_path, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", lib_path, *objs, "-lcudart"]
        subprocess.check_call(cmd)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

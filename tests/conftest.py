import os
import sys

import numpy as np
import pytest

# Several bands of one problem live in ONE process in the GPU tests, and the halo link waits on flag words at
# stream level: streams that alias one hardware queue would turn such a wait into a false dependency on the
# neighbour's sends.  Give every stream its own queue (must be set before CUDA initialises).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and a kernel's first launch must not fall into such a wait: CUDA loads kernels lazily and a load may
# synchronise the context (the library also pre-loads what a band step launches, GridSolver::preload_kernels)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "fast-poisson-image-editing_b200")
for p in (ROOT, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built artefacts (they are git-ignored): build the C-ABI library once
    # (nvcc cross-compiles without a GPU); the C oracle builds itself on first use
    try:
        from fpie_b200 import _build

        if not os.path.exists(_build.LIB_PATH):
            _build.build()
    except Exception as exc:  # the ABI tests will then fail loudly with the import error
        print(f"[conftest] could not build libfpie_b200.so: {exc}")


def _cuda_devices() -> int:
    try:
        from fpie_b200 import _lib

        return int(_lib.load().fpie_b200_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device: gpu-marked tests are skipped, not failed."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "fpie_numpy_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


GOLDEN_CASES = ("smoke6", "rng24", "ring_off", "holes_full", "disk_sat")
MODES = ("max", "src", "avg")


def golden_case(golden, name):
    return dict(
        src=golden[f"{name}/src"],
        mask=golden[f"{name}/mask"],
        tgt=golden[f"{name}/tgt"],
        off_src=tuple(int(v) for v in golden[f"{name}/off_src"]),
        off_tgt=tuple(int(v) for v in golden[f"{name}/off_tgt"]),
        steps=[int(v) for v in golden[f"{name}/steps"]],
    )

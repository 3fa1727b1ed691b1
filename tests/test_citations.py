"""Every ``path:line`` citation of the reference in the C header, the docs, the oracle and the host
layer must point at a file that exists in the reference checkout and at line numbers it has.  Runs
only where the checkout is present (the build container); skips on the GPU box."""

import os
import re

import pytest
from conftest import PKG_ROOT, ROOT

REF = os.environ.get("FPIE_REFERENCE", "/root/reference")
CITE = re.compile(r"(?<![\w/.-])((?:fpie/|tests/|docs/)?[\w/]+\.(?:py|cc|cu|h|md|txt)):(\d+)(?:-(\d+))?")

SOURCES = [
    os.path.join(ROOT, "include", "fpie_b200.h"),
    os.path.join(ROOT, "DESIGN.md"),
    os.path.join(ROOT, "INTEGRATION.md"),
    os.path.join(ROOT, "oracle", "np_oracle.py"),
    os.path.join(ROOT, "oracle", "jacobi_oracle.c"),
    os.path.join(PKG_ROOT, "fpie_b200", "process.py"),
    os.path.join(PKG_ROOT, "fpie_b200", "solver.py"),
    os.path.join(PKG_ROOT, "fpie_b200", "band.py"),
    os.path.join(PKG_ROOT, "fpie_b200", "register.py"),
    os.path.join(PKG_ROOT, "csrc", "grid.cu"),
    os.path.join(PKG_ROOT, "csrc", "equ.cu"),
    os.path.join(PKG_ROOT, "csrc", "prep.cu"),
]


def _resolve(path):
    """Citations are written relative to the checkout, to fpie/, or to fpie/core/."""
    for prefix in ("", "fpie", os.path.join("fpie", "core")):
        full = os.path.join(REF, prefix, path)
        if os.path.isfile(full):
            return full
    return None


def test_reference_citations_resolve():
    if not os.path.isdir(os.path.join(REF, "fpie")):
        pytest.skip("reference checkout not present")
    checked, bad = 0, []
    for src in SOURCES:
        text = open(src, encoding="utf-8").read()
        for m in CITE.finditer(text):
            path, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            if path.startswith(("tests/test_", "tests/golden", "tests/band")) and not os.path.isfile(os.path.join(REF, path)):
                continue  # our own tests, not the reference's
            full = _resolve(path)
            if full is None:
                if path.split("/")[0] in ("fpie", "docs") or path.startswith(("core/", "cuda/", "openmp/", "mpi/", "gcc/")):
                    bad.append(f"{os.path.relpath(src, ROOT)}: {m.group(0)} -> no such file")
                continue
            n_lines = sum(1 for _ in open(full, encoding="utf-8", errors="replace"))
            checked += 1
            if not (1 <= lo <= hi <= n_lines):
                bad.append(f"{os.path.relpath(src, ROOT)}: {m.group(0)} -> file has {n_lines} lines")
    assert checked > 100, f"only {checked} citations found: the pattern no longer matches the documents"
    assert not bad, "\n".join(bad)

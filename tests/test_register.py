"""``-b b200`` registration against the reference's own CLI plumbing
(fpie/args.py:25-31, fpie/cli.py:16-32).  Needs the reference checkout, which
exists only in the build container; skipped elsewhere."""

import os
import subprocess
import sys

import pytest
from conftest import PKG_ROOT

REF = os.environ.get("FPIE_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "fpie")), reason="reference checkout not present")


def _run(code: str) -> subprocess.CompletedProcess:
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG_ROOT, REF]))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)


@needs_ref
def test_backend_becomes_selectable():
    r = _run(
        "import sys, fpie_b200; fpie_b200.register(); "
        "from fpie.cli import main; sys.argv = ['fpie', '--check-backend']; main()"
    )
    assert r.returncode == 0, r.stderr
    assert "b200" in r.stdout and "numpy" in r.stdout


@needs_ref
def test_other_backends_untouched_and_b200_dispatches():
    r = _run(
        "import numpy as np, fpie_b200\n"
        "Equ, Grid = fpie_b200.register()\n"
        "import fpie.process as fp\n"
        "p = fp.GridProcessor('max', 'numpy')\n"
        "assert type(p.core).__module__ == 'fpie.np_solver', type(p.core)\n"
        "src = np.zeros((6, 6, 3), np.uint8); mask = np.zeros((6, 6), np.uint8); mask[2:4, 2:4] = 255\n"
        "tgt = np.ones((6, 6, 3), np.uint8) * 10\n"
        "assert p.reset(src, mask, tgt, (0, 0), (0, 0)) == 16\n"
        "out, err = p.step(2); assert out.dtype == np.uint8 and err.shape == (3,)\n"
        "try:\n"
        "    fp.EquProcessor('max', 'b200')\n"
        "except RuntimeError as e:\n"
        "    print('loud:', e)\n"
        "else:\n"
        "    print('constructed')\n"
    )
    assert r.returncode == 0, r.stderr
    assert "loud:" in r.stdout or "constructed" in r.stdout


@needs_ref
def test_argparse_accepts_b200():
    r = _run(
        "import sys, fpie_b200; fpie_b200.register(); from fpie.args import get_args; "
        "f = sys.executable; sys.argv = ['fpie', '-b', 'b200', '--method', 'grid', '-s', f, '-t', f, '-m', f, '-o', 'c']; "
        "a = get_args('cli'); print(a.backend, a.method)"
    )
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == ["b200", "grid"]

"""The drop-in, exercised AS a drop-in: the reference's own command line (fpie/cli.py:16-62) and
the reference's own Processors (fpie/process.py:146-395) running over the b200 core on a GPU.

The unmodified reference is installed into ``baseline/_ref`` by ``baseline/install_ref.sh``
(git-ignored, shipped to the GPU box by gpurun); nothing here reads ``/root/reference``.
Every case runs the stock ``fpie.cli.main()`` twice in a fresh interpreter -- once with the
reference's numpy backend, once with ``-b b200`` after ``fpie_b200.register()`` -- on the same
PNG files and compares the PNGs the two runs wrote, byte for byte.
"""

import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import PKG_ROOT, ROOT

REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF_INSTALL, "fpie", "cli.py")),
                               reason="baseline/_ref (the installed reference) is absent: run baseline/install_ref.sh")
pytestmark = [pytest.mark.gpu, needs_ref]

# One interpreter per wiring runs every job (importing torch once per job would dominate the run time):
# argv lists come in as JSON, every job calls the reference's `main()` in its own working directory.
CLI = """
import contextlib, io, json, os, sys
mode, jobs = sys.argv[1], json.loads(sys.argv[2])
if mode != "stock":
    import fpie_b200
    fpie_b200.register(fused=(mode != "core"), stage_io=(mode == "staged"))
from fpie.cli import main
logs = {}
for name, cwd, argv in jobs:
    os.chdir(cwd)
    sys.argv = ["fpie", *argv]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        main()
    if mode == "staged":
        import fpie_b200.io
        fpie_b200.io.flush_writes()  # (python -m fpie_b200.register does this itself)
    logs[name] = buf.getvalue()
print(json.dumps(logs))
"""


def run_cli(mode, jobs):
    """jobs: [(name, cwd, argv)]; returns {name: stdout of fpie.cli.main()}."""
    import json

    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG_ROOT, REF_INSTALL]))
    r = subprocess.run([sys.executable, "-c", CLI, mode, json.dumps(jobs)], capture_output=True, text=True, env=env,
                       timeout=900)
    assert r.returncode == 0, f"{mode}\n{r.stdout}\n{r.stderr}"
    return json.loads(r.stdout.strip().splitlines()[-1])


def read_png(path):
    import cv2

    img = cv2.imread(path)
    assert img is not None, path
    return img


def errors_of(stdout):
    """The `Iter N, abs error [...]` lines the CLI prints (cli.py:52)."""
    out = []
    for line in stdout.splitlines():
        if line.startswith("Iter "):
            vals = line.split("abs error")[1].strip().strip("[]").split()
            out.append(np.array([float(v) for v in vals]))
    return out


CASES = [
    # method, gradient, mask kind, h, w
    ("grid", "max", "circle", 72, 96),
    ("grid", "src", "holes", 64, 80),
    ("grid", "avg", "ring", 96, 72),
    ("equ", "max", "circle", 72, 96),
    ("equ", "src", "star", 80, 80),
    ("equ", "avg", "holes", 64, 80),
]
OFFSET_FLAGS = ["-n", "90", "-p", "30", "-h0", "12", "-w0", "20", "-h1", "40", "-w1", "55"]
# "staged": the fused wiring with fpie.io swapped for fpie_b200.io (page-locked staging, PNG encoding in a worker)
MODES3 = (("stock", "numpy"), ("fused", "b200"), ("core", "b200"), ("staged", "b200"))


@pytest.fixture(scope="module")
def cli_runs(tmp_path_factory):
    """Every CLI job of this file, run once per wiring: {(mode, job name): (stdout, directory)}."""
    import cv2

    from fpie_b200 import synth

    root = tmp_path_factory.mktemp("dropin")
    jobs = {m: [] for m, _ in MODES3}
    where = {}

    def add(name, images, flags):
        for mode, backend in MODES3:
            d = root / f"{name}-{mode}"
            d.mkdir()
            for fname, img in images.items():
                cv2.imwrite(str(d / f"{fname}.png"), img)
            jobs[mode].append((name, str(d), ["-b", backend, *flags, "-s", "src.png", "-m", "mask.png", "-t", "tgt.png",
                                              "-o", "out.png"]))
            where[(mode, name)] = d

    for method, gradient, kind, h, w in CASES:
        src, mask, tgt = synth.make_problem(kind, h, w, seed=3)
        add(f"{method}-{gradient}-{kind}", dict(src=src, mask=mask, tgt=tgt), ["--method", method, "-g", gradient, "-n", "150"])
    # offsets, progress images, a 3-channel soft mask on images of three different sizes
    rng = np.random.default_rng(11)
    src = rng.integers(0, 256, (90, 110, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, (120, 140, 3), dtype=np.uint8)
    mask = synth.make_mask("star", 60, 64)
    mask3 = np.clip(mask[..., None].astype(np.int32) + rng.integers(-120, 120, (60, 64, 3)), 0, 255).astype(np.uint8)
    for method in ("grid", "equ"):
        add(f"offsets-{method}", dict(src=src, mask=mask3, tgt=tgt), ["--method", method, "-g", "max", *OFFSET_FLAGS])
    out = {}
    for mode, _ in MODES3:
        logs = run_cli(mode, jobs[mode])
        for name, text in logs.items():
            out[(mode, name)] = (text, where[(mode, name)])
    return out


@pytest.mark.parametrize("wiring", ["fused", "core"])
@pytest.mark.parametrize("method,gradient,kind,h,w", CASES)
def test_cli_b200_equals_cli_numpy(cli_runs, method, gradient, kind, h, w, wiring):
    name = f"{method}-{gradient}-{kind}"
    ref_out, ref_dir = cli_runs[("stock", name)]
    got_out, got_dir = cli_runs[(wiring, name)]
    assert "with b200 backend" in got_out and "with numpy backend" in ref_out
    assert np.array_equal(read_png(str(got_dir / "out.png")), read_png(str(ref_dir / "out.png")))
    # same `# of vars` line (process.py:190, 352) and the same residuals
    nvars = [ln for ln in ref_out.splitlines() if ln.startswith("# of vars")]
    assert nvars and nvars == [ln for ln in got_out.splitlines() if ln.startswith("# of vars")]
    for a, b in zip(errors_of(ref_out), errors_of(got_out), strict=True):
        np.testing.assert_allclose(b, a, rtol=1e-4)


@pytest.mark.parametrize("wiring", ["fused", "core", "staged"])
@pytest.mark.parametrize("method", ["grid", "equ"])
def test_cli_offsets_progress_images_and_soft_mask(cli_runs, method, wiring):
    """`-h0/-w0/-h1/-w1` offsets, `-p` progress images (cli.py:48-57: repeated step calls on one
    solver, a PNG every P sweeps) and a 3-channel soft mask -- all PNGs equal the numpy backend's."""
    _, ref_dir = cli_runs[("stock", f"offsets-{method}")]
    got_out, got_dir = cli_runs[(wiring, f"offsets-{method}")]
    assert len(errors_of(got_out)) == 3
    for f in ("iter00030.png", "iter00060.png", "out.png"):
        assert np.array_equal(read_png(str(got_dir / f)), read_png(str(ref_dir / f))), f


def test_reference_processor_over_b200_core_matches_openmp_grid(tmp_path):
    """`register(fused=False)`: the reference's GridProcessor (host numpy preprocessing,
    process.py:321-386) over the b200 core vs the same Processor over the reference's own compiled
    OpenMP core (true Jacobi, openmp/grid.cc:80-103) at a size numpy would take minutes for."""
    code = """
import numpy as np, fpie_b200
from fpie_b200 import synth
fpie_b200.register(fused=False)
import fpie.process as fp
assert "openmp" in fp.ALL_BACKEND, fp.ALL_BACKEND
src, mask, tgt = synth.make_problem("circle", 700, 900, seed=5)
a = fp.GridProcessor("max", "openmp", 4, 100, 1024, 2, 16)
b = fp.GridProcessor("max", "b200", 4, 100, 1024, 2, 16)
assert isinstance(b, fp.BaseProcessor) and type(b).reset is type(a).reset  # the reference's own host preprocessing
assert a.reset(src, mask, tgt, (0, 0), (0, 0)) == b.reset(src, mask, tgt, (0, 0), (0, 0))
a.sync(); b.sync()
for _ in range(2):
    ia, ea = a.step(400)
    ib, eb = b.step(400)
    assert ia.dtype == ib.dtype == np.uint8 and ia.shape == ib.shape
    # the OpenMP core adds in another order (g+U+L+R+D, divided in double: openmp/grid.cc:29-44) than numpy and this backend
    # (np_solver.py:83-88): last-bit differences in fp32, hence the stated tolerances -- uint8 within 1, err 1e-4
    assert int(np.abs(ia.astype(int) - ib.astype(int)).max()) <= 1
    assert float((ia != ib).mean()) < 1e-3
    np.testing.assert_allclose(eb, ea, rtol=1e-4)
print("ok", b.core.info()["launches"])
"""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG_ROOT, REF_INSTALL]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path),
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok")

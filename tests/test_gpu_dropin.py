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

CLI = """
import sys
mode = sys.argv.pop(1)
if mode != "stock":
    import fpie_b200
    fpie_b200.register(fused=(mode == "fused"))
from fpie.cli import main
main()
"""


def run_cli(mode, argv, cwd):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG_ROOT, REF_INSTALL]))
    r = subprocess.run([sys.executable, "-c", CLI, mode, *argv], capture_output=True, text=True, env=env, cwd=cwd,
                       timeout=600)
    assert r.returncode == 0, f"{mode} {argv}\n{r.stdout}\n{r.stderr}"
    return r.stdout


def write_problem(tmp, kind, h, w, seed, mask_channels=1):
    import cv2

    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem(kind, h, w, seed)
    if mask_channels == 3:  # soft, coloured mask: the Processor thresholds the channel mean (process.py:209-215)
        rng = np.random.default_rng(seed + 7)
        mask = np.clip(mask[..., None].astype(np.int32) + rng.integers(-140, 140, (h, w, 3)), 0, 255).astype(np.uint8)
    cv2.imwrite(os.path.join(tmp, "src.png"), src)
    cv2.imwrite(os.path.join(tmp, "mask.png"), mask)
    cv2.imwrite(os.path.join(tmp, "tgt.png"), tgt)


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
    # method, gradient, mask kind, h, w, extra CLI flags
    ("grid", "max", "circle", 72, 96, []),
    ("grid", "src", "holes", 64, 80, []),
    ("grid", "avg", "ring", 96, 72, []),
    ("equ", "max", "circle", 72, 96, []),
    ("equ", "src", "star", 80, 80, []),
    ("equ", "avg", "holes", 64, 80, []),
]


@pytest.mark.parametrize("fused", ["fused", "core"])
@pytest.mark.parametrize("method,gradient,kind,h,w,extra", CASES)
def test_cli_b200_equals_cli_numpy(tmp_path, method, gradient, kind, h, w, extra, fused):
    tmp = str(tmp_path)
    write_problem(tmp, kind, h, w, seed=3)
    common = ["--method", method, "-g", gradient, "-s", "src.png", "-m", "mask.png", "-t", "tgt.png", "-n", "150",
              *extra]
    ref_out = run_cli("stock", ["-b", "numpy", *common, "-o", "ref.png"], tmp)
    got_out = run_cli(fused, ["-b", "b200", *common, "-o", "b200.png"], tmp)
    assert "with b200 backend" in got_out
    assert np.array_equal(read_png(os.path.join(tmp, "b200.png")), read_png(os.path.join(tmp, "ref.png")))
    # same `# of vars` line (process.py:190, 352) and the same residuals
    nvars = [ln for ln in ref_out.splitlines() if ln.startswith("# of vars")]
    assert nvars and nvars == [ln for ln in got_out.splitlines() if ln.startswith("# of vars")]
    for a, b in zip(errors_of(ref_out), errors_of(got_out), strict=True):
        np.testing.assert_allclose(b, a, rtol=1e-4)


@pytest.mark.parametrize("method", ["grid", "equ"])
def test_cli_offsets_progress_images_and_soft_mask(tmp_path, method):
    """`-h0/-w0/-h1/-w1` offsets, `-p` progress images (cli.py:48-57: repeated step calls on one
    solver, a PNG every P sweeps) and a 3-channel soft mask -- all PNGs equal the numpy backend's."""
    import cv2

    from fpie_b200 import synth

    tmp = str(tmp_path)
    rng = np.random.default_rng(11)
    src = rng.integers(0, 256, (90, 110, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, (120, 140, 3), dtype=np.uint8)
    mask = synth.make_mask("star", 60, 64)
    mask3 = np.clip(mask[..., None].astype(np.int32) + rng.integers(-120, 120, (60, 64, 3)), 0, 255).astype(np.uint8)
    for name, img in (("src", src), ("mask", mask3), ("tgt", tgt)):
        cv2.imwrite(os.path.join(tmp, f"{name}.png"), img)
    common = ["--method", method, "-g", "max", "-s", "src.png", "-m", "mask.png", "-t", "tgt.png", "-n", "90", "-p",
              "30", "-h0", "12", "-w0", "20", "-h1", "40", "-w1", "55"]
    outs = {}
    for mode, backend in (("stock", "numpy"), ("fused", "b200"), ("core", "b200")):
        sub = os.path.join(tmp, mode)
        os.makedirs(sub)
        for name in ("src", "mask", "tgt"):
            os.symlink(os.path.join(tmp, f"{name}.png"), os.path.join(sub, f"{name}.png"))
        run_cli(mode, ["-b", backend, *common, "-o", "out.png"], sub)
        outs[mode] = {f: read_png(os.path.join(sub, f)) for f in ("iter00030.png", "iter00060.png", "out.png")}
    for mode in ("fused", "core"):
        for f, img in outs["stock"].items():
            assert np.array_equal(outs[mode][f], img), (mode, f)


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
    assert np.array_equal(ia, ib), int(np.abs(ia.astype(int) - ib.astype(int)).max())
    np.testing.assert_allclose(eb, ea, rtol=1e-4)
print("ok", b.core.info()["launches"])
"""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG_ROOT, REF_INSTALL]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path),
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok")

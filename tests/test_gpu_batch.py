"""Batched small edits (BASELINE config 5): many independent patches solved as
one mosaic.  Every patch must equal the oracle's solve of that patch alone."""

import numpy as np
import pytest
from conftest import MODES

from oracle import c_oracle, np_oracle

pytestmark = pytest.mark.gpu


def _batch(b, n, m, seed):
    from fpie_b200 import synth

    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, (b, n, m, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, (b, n, m, 3), dtype=np.uint8)
    mask = np.stack([synth.make_mask(synth.MASK_KINDS[i % 5], n, m, seed=i) for i in range(b)])
    return src, mask, tgt


def _oracle_patch(src, mask, tgt, mode, iters):
    """Solve one patch alone; returns (uint8 canvas, err64, fp32 state canvas)."""
    m, t, g, (x0, x1, y0, y1) = np_oracle.grid_system(src, mask, tgt, (0, 0), (0, 0), mode)
    state = c_oracle.grid_sweeps(m, t, g, iters)
    canvas = tgt.copy()
    canvas[x0:x1, y0:y1] = c_oracle.clip_u8(state)
    full = tgt.astype(np.float32)
    full[x0:x1, y0:y1] = state
    return canvas, c_oracle.grid_residual(m, state, g)[1], full


@pytest.mark.parametrize("b,n,m,iters", [(1, 40, 36, 9), (7, 64, 48, 30), (37, 96, 128, 41)])
@pytest.mark.parametrize("mode", MODES)
def test_batch_matches_per_patch_oracle(b, n, m, iters, mode):
    import fpie_b200

    src, mask, tgt = _batch(b, n, m, seed=b * 7 + n)
    proc = fpie_b200.BatchGridProcessor(mode, "b200")
    assert proc.reset(src, mask, tgt) == b * n * m
    out, err = proc.step(iters)
    assert out.shape == (b, n, m, 3) and out.dtype == np.uint8 and err.shape == (b, 3)
    state = proc.core.batch_state()
    for i in range(b):
        canvas, e64, full = _oracle_patch(src[i], mask[i], tgt[i], mode, iters)
        np.testing.assert_array_equal(state[i], full)
        np.testing.assert_array_equal(out[i], canvas)
        np.testing.assert_allclose(err[i], e64, rtol=1e-4, atol=1e-3)


def test_batch_steps_accumulate_and_patches_do_not_interact():
    import fpie_b200

    src, mask, tgt = _batch(12, 64, 64, seed=1)
    proc = fpie_b200.BatchGridProcessor("max", "b200")
    proc.reset(src, mask, tgt)
    proc.step(8)
    a, ea = proc.step(17)
    proc.reset(src, mask, tgt)
    b, eb = proc.step(25)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(ea, eb, rtol=1e-6)
    # changing one patch leaves every other patch bit-identical
    src2 = src.copy()
    src2[5] = 255 - src2[5]
    proc.reset(src2, mask, tgt)
    c, _ = proc.step(25)
    keep = [i for i in range(12) if i != 5]
    np.testing.assert_array_equal(c[keep], b[keep])
    assert not np.array_equal(c[5], b[5])


def test_config5_sized_batch():
    """256x256 patches as in BASELINE config 5 (a slice of the 512-patch batch), grad src."""
    import fpie_b200

    src, mask, tgt = _batch(24, 256, 256, seed=3)
    proc = fpie_b200.BatchGridProcessor("src", "b200")
    proc.reset(src, mask, tgt)
    out, err = proc.step(120)
    for i in (0, 7, 13, 23):
        canvas, e64, _ = _oracle_patch(src[i], mask[i], tgt[i], "src", 120)
        np.testing.assert_array_equal(out[i], canvas)
        np.testing.assert_allclose(err[i], e64, rtol=1e-4, atol=1e-3)


# ---- the persistent small-image kernel (csrc/patch.cuh) ---------------------------------------------
# Steps of >= 32 sweeps on patches / images of at most 512 x 256 pixels run as ONE launch in which a
# cluster keeps each (patch, plane) in registers.  Same arithmetic as the tiled kernel: every result
# below must be bit-identical to the oracle and to the mosaic path.


def _full_batch(b, n, m, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, (b, n, m, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, (b, n, m, 3), dtype=np.uint8)
    return src, np.full((b, n, m), 255, np.uint8), tgt


@pytest.mark.parametrize("b,n,m,iters,full", [
    (5, 256, 256, 64, True),    # config 5 shape, full-square masks: the select-free (frame-only) stream, cluster of 4
    (5, 256, 256, 45, False),   # ... arbitrary masks: per-pixel selects
    (3, 64, 64, 100, True),     # 4 columns and 4 rows per thread, cluster of 2
    (4, 100, 128, 33, False),   # rows do not fill the cluster (padding rows), 4 columns per thread
    (3, 200, 252, 40, False),   # columns do not fill the warp
    (2, 300, 256, 37, True),    # 8 rows per thread (more than 256 rows)
    (2, 512, 200, 35, False),   # the tallest supported patch
    (40, 32, 32, 50, False),    # more clusters than the device holds at once: the persistent loop
    (40, 256, 256, 40, True),   # a batch too large for 4 rows per thread to be resident at once: 8 rows, cluster of 4
])
def test_persistent_kernel_matches_oracle_and_mosaic(b, n, m, iters, full, monkeypatch):
    import fpie_b200

    monkeypatch.setenv("FPIE_B200_PATCH", "2")  # (any value but "0": on -- the kernel is the policy for every small image it fits)
    src, mask, tgt = _full_batch(b, n, m, seed=n + m) if full else _batch(b, n, m, seed=n + m)
    proc = fpie_b200.BatchGridProcessor("max", "b200")
    proc.reset(src, mask, tgt)
    info = proc.core.patch_info()
    assert info["usable"] and info["cluster"] == -(-n // (8 * info["rows_per_thread"]))
    # 4 rows per thread (a plane on up to 8 CTAs) while every item's cluster is resident at once, else 8
    sms = fpie_b200.device_info(0)["sm_count"]
    cl4 = -(-n // 32)
    assert info["rows_per_thread"] == (4 if (n <= 128 or (cl4 <= 8 and 3 * b * cl4 <= 2 * sms)) else 8)
    assert info["cols_per_thread"] == (4 if m <= 128 else 8)
    out, err = proc.step(iters)
    assert proc.core.patch_info()["launches"] == 1 and proc.core.info()["launches"] < 12
    state = proc.core.batch_state()
    for i in range(b):
        canvas, e64, fullstate = _oracle_patch(src[i], mask[i], tgt[i], "max", iters)
        np.testing.assert_array_equal(state[i], fullstate)
        np.testing.assert_array_equal(out[i], canvas)
        np.testing.assert_allclose(err[i], e64, rtol=1e-4, atol=1e-3)
    # a second step continues from the state the kernel wrote back; the mosaic path (persistent kernel off)
    # gives the same bits
    out2, err2 = proc.step(iters + 1)
    monkeypatch.setenv("FPIE_B200_PATCH", "0")
    ref = fpie_b200.BatchGridProcessor("max", "b200")
    ref.reset(src, mask, tgt)
    assert not ref.core.patch_info()["usable"]
    ref.step(iters)
    want2, werr2 = ref.step(iters + 1)
    np.testing.assert_array_equal(out2, want2)
    np.testing.assert_array_equal(proc.core.batch_state(), ref.core.batch_state())
    np.testing.assert_allclose(err2, werr2, rtol=1e-6)


@pytest.mark.parametrize("kind,n,m", [("circle", 256, 256), ("star", 180, 140), ("holes", 96, 250), ("square", 400, 256)])
@pytest.mark.parametrize("rows", ["4", "8"])
def test_persistent_kernel_single_images(kind, n, m, rows, monkeypatch):
    """One small image through the ordinary GridProcessor / EquProcessor: the GUI's reset + step per click."""
    import fpie_b200
    from fpie_b200 import synth

    monkeypatch.setenv("FPIE_B200_PATCH_ROWS", rows)
    monkeypatch.setenv("FPIE_B200_PATCH", "2")  # (on; single small images take the persistent kernel by policy)
    src, mask, tgt = synth.make_problem(kind, n, m, seed=7)
    proc = fpie_b200.GridProcessor("avg", "b200")
    proc.reset(src, mask, tgt, (0, 0), (0, 0))
    out, err = proc.step(90)
    assert proc.core.patch_info()["launches"] == 1
    proc.step(10)  # (a short step stays on the tiled kernel)
    assert proc.core.patch_info()["launches"] == 1
    out, err = proc.step(50)
    want = np_oracle.GridOracle("avg")
    want.reset(src, mask, tgt)
    want.t = c_oracle.grid_sweeps(want.mask, want.t, want.g, 150)
    wout, werr = want.step(0)
    np.testing.assert_array_equal(proc.core.state(), want.t)
    np.testing.assert_array_equal(out, wout)
    np.testing.assert_allclose(err, werr, rtol=1e-4)
    equ = fpie_b200.EquProcessor("avg", "b200")
    equ.reset(src, mask, tgt, (0, 0), (0, 0))
    eout, _ = equ.step(150)
    np.testing.assert_array_equal(eout, wout)


def test_persistent_kernel_is_not_used_when_a_tile_shape_is_requested():
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("circle", 128, 128, seed=1)
    core = fpie_b200.GridSolver(8, 8, block_k=8, variant=24)
    core.reset_from_images(src, mask, tgt, (0, 0), (0, 0), "max")
    core.step(64)
    assert not core.patch_info()["usable"] and core.patch_info()["launches"] == 0


def test_persistent_kernel_policy():
    """Used by itself for batches and single images of up to 256 columns and 512 rows (any mask); 4 rows per thread
    while all clusters are resident at once and a plane needs at most 8 CTAs."""
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = _full_batch(4, 64, 64, seed=2)
    proc = fpie_b200.BatchGridProcessor("max", "b200")
    proc.reset(src, mask, tgt)
    assert proc.core.patch_info()["usable"]
    proc.reset(src[:1], mask[:1], tgt[:1])
    assert proc.core.patch_info()["usable"]
    one = fpie_b200.GridProcessor("max", "b200")
    one.reset(*synth.make_problem("circle", 128, 128, seed=1), (0, 0), (0, 0))
    one.step(64)
    assert one.core.patch_info()["usable"] and one.core.patch_info()["launches"] == 1
    one.reset(*synth.make_problem("circle", 200, 200, seed=1), (0, 0), (0, 0))
    one.step(64)
    assert one.core.patch_info()["usable"] and one.core.patch_info()["launches"] == 2
    one.reset(*synth.make_problem("holes", 200, 300, seed=1), (0, 0), (0, 0))  # wider than 256 columns: tiled
    one.step(64)
    assert not one.core.patch_info()["usable"] and one.core.patch_info()["launches"] == 2
    one.reset(*synth.make_problem("square", 256, 256, seed=1), (0, 0), (0, 0))
    info = one.core.patch_info()
    assert info["usable"] and (info["rows_per_thread"], info["cluster"]) == (4, 8)
    one.reset(*synth.make_problem("square", 300, 256, seed=1), (0, 0), (0, 0))
    info = one.core.patch_info()
    assert info["usable"] and (info["rows_per_thread"], info["cluster"]) == (8, 5)  # (more than 8 CTAs of 32 rows)

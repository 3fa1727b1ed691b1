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

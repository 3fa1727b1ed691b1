"""Pins the CPU oracle to the reference.

* numpy restatement vs golden vectors recorded from the reference's own numpy
  backend (tests/golden/make_golden.py), bit-for-bit;
* the known answers listed in SURVEY.md section 4 for the reference's own test
  fixture (tests/test_smoke.py:48-66);
* the C restatement vs the numpy restatement, bit-for-bit;
* the compiled reference OpenMP GridSolver (oracle/_ref, when built) vs the
  oracle -- the reference's own parity test, restated.
"""

import hashlib

import numpy as np
import pytest
from conftest import GOLDEN_CASES, MODES, golden_case

from oracle import c_oracle, np_oracle


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("mode", MODES)
def test_grid_system_matches_reference(golden, name, mode):
    c = golden_case(golden, name)
    mask, tgt, grad, _ = np_oracle.grid_system(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"], mode)
    key = f"{name}/grid/{mode}"
    np.testing.assert_array_equal(mask, golden[f"{key}/mask_crop"])
    np.testing.assert_array_equal(tgt, golden[f"{key}/tgt_crop"])
    np.testing.assert_array_equal(grad, golden[f"{key}/grad"])
    assert mask.size == int(golden[f"{key}/n"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("mode", MODES)
def test_equ_system_matches_reference(golden, name, mode):
    c = golden_case(golden, name)
    n, A, X, B, _ = np_oracle.equ_system(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"], mode)
    key = f"{name}/equ/{mode}"
    assert n == int(golden[f"{key}/n"])
    np.testing.assert_array_equal(A, golden[f"{key}/A"])
    np.testing.assert_array_equal(X, golden[f"{key}/X0"])
    np.testing.assert_array_equal(B, golden[f"{key}/B"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kind", ("equ", "grid"))
def test_processor_run_matches_reference(golden, name, mode, kind):
    c = golden_case(golden, name)
    orc = (np_oracle.EquOracle if kind == "equ" else np_oracle.GridOracle)(mode)
    orc.reset(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"])
    key = f"{name}/{kind}/{mode}"
    for si, it in enumerate(c["steps"]):
        img, err = orc.step(it)
        state = orc.X if kind == "equ" else orc.t
        np.testing.assert_array_equal(state, golden[f"{key}/state{si}"])  # fp32 bit-exact
        np.testing.assert_array_equal(img, golden[f"{key}/img{si}"])
        np.testing.assert_array_equal(err, golden[f"{key}/err{si}"])


# values recorded from the reference numpy backend for
# the fixture of tests/test_smoke.py:48-66 after step(5).
KNOWN = {
    "max": ([6024.8467, 5061.247, 5035.378], "55e65aee1fcd3065", 224648),
    "src": ([10017.986, 9229.196, 6978.8574], "0bc4374bd2f53876", 222387),
    "avg": ([5008.993, 4614.598, 3489.4287], "e261239f2d5a5575", 222493),
}


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kind", ("equ", "grid"))
def test_known_answers_rng24(mode, kind):
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    mask = np.zeros((24, 24), np.uint8)
    mask[2:-2, 2:-2] = (rng.random((20, 20)) > 0.35).astype(np.uint8) * 255
    orc = (np_oracle.EquOracle if kind == "equ" else np_oracle.GridOracle)(mode)
    n = orc.reset(src, mask, tgt)
    assert n == (267 if kind == "equ" else 484)
    img, err = orc.step(5)
    want_err, want_sha, want_sum = KNOWN[mode]
    np.testing.assert_allclose(err, want_err, rtol=2e-7)
    assert hashlib.sha1(img.tobytes()).hexdigest()[:16] == want_sha
    assert int(img.sum()) == want_sum


def test_smoke6_core_inputs():
    """SURVEY.md section 4, fixture 1: what the core sees for the 6x6 case."""
    src = np.zeros((6, 6, 3), np.uint8)
    mask = np.zeros((6, 6), np.uint8)
    mask[2:4, 2:4] = 255
    tgt = np.ones((6, 6, 3), np.uint8) * 10
    n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "max")
    assert n == 5
    assert A.tolist() == [[0, 0, 0, 0], [0, 3, 0, 2], [0, 4, 1, 0], [1, 0, 0, 4], [2, 0, 3, 0]]
    assert (X[1:] == 10).all() and (B[1:] == 20).all() and (X[0] == 0).all() and (B[0] == 0).all()
    m, t, g, _ = np_oracle.grid_system(src, mask, tgt, (0, 0), (0, 0), "max")
    assert m.size == 16


def test_empty_mask_raises():
    z = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(ValueError):
        np_oracle.grid_system(z, np.zeros((8, 8), np.uint8), z, (0, 0), (0, 0), "max")
    # a mask that only touches the frame is empty after the frame is cleared
    m = np.zeros((8, 8), np.uint8)
    m[0, :] = 255
    with pytest.raises(ValueError):
        np_oracle.equ_system(z, m, z, (0, 0), (0, 0), "max")


def test_box_outside_image_raises():
    src = np.zeros((10, 10, 3), np.uint8)
    tgt = np.zeros((10, 10, 3), np.uint8)
    mask = np.full((10, 10), 255, np.uint8)
    with pytest.raises(ValueError):
        np_oracle.grid_system(src, mask, tgt, (0, 0), (3, 0), "max")
    with pytest.raises(ValueError):
        np_oracle.equ_system(src, mask, tgt, (-2, 0), (0, 0), "max")


def _random_grid(n, m, seed, density=0.7):
    rng = np.random.default_rng(seed)
    mask = np.zeros((n, m), np.int32)
    mask[1:-1, 1:-1] = rng.random((n - 2, m - 2)) < density
    tgt = rng.integers(0, 256, (n, m, 3)).astype(np.float32)
    grad = (rng.integers(-2040, 2041, (n, m, 3)) / 2).astype(np.float32)
    grad[mask == 0] = 0
    return mask, tgt, grad


@pytest.mark.parametrize("shape", [(3, 3), (5, 9), (33, 47), (64, 64), (97, 83)])
def test_c_grid_matches_numpy_bitexact(shape):
    mask, tgt, grad = _random_grid(*shape, seed=shape[0] * 131 + shape[1])
    for iters in (0, 1, 7, 40):
        want = np_oracle.grid_sweeps(mask, tgt, grad, iters)
        got = c_oracle.grid_sweeps(mask, tgt, grad, iters, threads=3)
        np.testing.assert_array_equal(got, want)
        e32, e64 = c_oracle.grid_residual(mask, got, grad)
        np.testing.assert_array_equal(e32, np_oracle.grid_residual(mask, want, grad))
        np.testing.assert_allclose(e64, np_oracle.grid_residual_f64(mask, want, grad), rtol=1e-12)
        np.testing.assert_array_equal(c_oracle.clip_u8(got), np_oracle.clip_u8(want))


@pytest.mark.parametrize("name", ("rng24", "ring_off", "holes_full", "disk_sat"))
def test_c_equ_matches_numpy_bitexact(golden, name):
    A, X, B = (golden[f"{name}/equ/avg/{k}"] for k in ("A", "X0", "B"))
    for iters in (0, 1, 9, 33):
        want = np_oracle.equ_sweeps(A, X, B, iters)
        got = c_oracle.equ_sweeps(A, X, B, iters, threads=2)
        np.testing.assert_array_equal(got, want)
        e32, e64 = c_oracle.equ_residual(A, got, B)
        np.testing.assert_array_equal(e32, np_oracle.equ_residual(A, want, B))
        np.testing.assert_allclose(e64, np_oracle.equ_residual_f64(A, want, B), rtol=1e-12)


def test_clip_u8_truncates():
    v = np.array([-3.5, -0.0, 0.0, 0.999, 1.0, 127.5, 254.999, 255.0, 255.5, 1e9], np.float32)
    want = np.array([0, 0, 0, 0, 1, 127, 254, 255, 255, 255], np.uint8)
    np.testing.assert_array_equal(np_oracle.clip_u8(v), want)
    np.testing.assert_array_equal(c_oracle.clip_u8(v), want)


@pytest.mark.parametrize("parts,depth", [(2, 4), (4, 8), (8, 5), (8, 1), (3, 16)])
def test_banded_model_is_exact(parts, depth):
    """Deep-halo row bands reproduce global Jacobi bit-for-bit."""
    mask, tgt, grad = _random_grid(97, 83, seed=5, density=0.65)
    want = np_oracle.grid_sweeps(mask, tgt, grad, 37)
    got = np_oracle.grid_sweeps_banded(mask, tgt, grad, 37, parts, depth)
    np.testing.assert_array_equal(got, want)


def test_band_offsets_rule():
    assert np_oracle.band_offsets(10, 3) == [0, 4, 7, 10]
    assert np_oracle.band_offsets(8, 8) == list(range(9))
    assert np_oracle.band_offsets(5, 8) == [0, 1, 2, 3, 4, 5, 5, 5, 5]


def test_reference_openmp_grid_agrees():
    """The reference's own parity test (tests/test_smoke.py:48-66) against the
    compiled reference core: u8 identical, err within rtol 1e-5."""
    core = c_oracle.load_reference_core("core_openmp")
    if core is None:
        pytest.skip("oracle/_ref/core_openmp not built (needs /root/reference)")
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    mask = np.zeros((24, 24), np.uint8)
    mask[2:-2, 2:-2] = (rng.random((20, 20)) > 0.35).astype(np.uint8) * 255
    for mode in MODES:
        m, t, g, _ = np_oracle.grid_system(src, mask, tgt, (0, 0), (0, 0), mode)
        solver = core.GridSolver(1, 1, 4)
        solver.reset(m.size, m, t, g)
        img, err = solver.step(5)
        state = np_oracle.grid_sweeps(m, t, g, 5)
        np.testing.assert_array_equal(img, np_oracle.clip_u8(state))
        np.testing.assert_allclose(err, np_oracle.grid_residual(m, state, g), rtol=1e-5, atol=1e-5)


def test_reference_openmp_grid_medium():
    core = c_oracle.load_reference_core("core_openmp")
    if core is None:
        pytest.skip("oracle/_ref/core_openmp not built (needs /root/reference)")
    mask, tgt, grad = _random_grid(130, 171, seed=3)
    solver = core.GridSolver(2, 16, 4)
    solver.reset(mask.size, mask, tgt, grad)
    img, err = solver.step(25)
    state = c_oracle.grid_sweeps(mask, tgt, grad, 25)
    # openmp adds in the order g+up+left+right+down (openmp/grid.cc:35-43): +-1 on u8
    diff = np.abs(img.astype(np.int16) - c_oracle.clip_u8(state).astype(np.int16))
    assert diff.max() <= 1
    _, e64 = c_oracle.grid_residual(mask, state, grad)
    np.testing.assert_allclose(err, e64, rtol=1e-4)


def test_redblack_restatement_matches_reference_openmp(golden):
    """The red-black Gauss-Seidel restatement vs the compiled reference core_openmp.EquSolver
    (fpie/core/openmp/equ.cc): ids, n_mid behaviour, uint8 output, err."""
    core = c_oracle.load_reference_core("core_openmp")
    if core is None:
        pytest.skip("oracle/_ref/core_openmp not built (needs /root/reference)")
    for name, mode in (("rng24", "max"), ("ring_off", "avg"), ("holes_full", "src")):
        c = golden_case(golden, name)
        m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(c["mask"])
        crop = np.ascontiguousarray(m_full[x0:x1, y0:y1])
        solver = core.EquSolver(3)
        ref_ids = solver.partition(crop)
        ids, n_mid = np_oracle.partition_redblack(crop)
        np.testing.assert_array_equal(ref_ids, ids)
        n, A, X, B, _ = np_oracle.equ_system(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"], mode, ids=ids)
        solver.reset(n, A, X, B)
        img, err = solver.step(9)
        want = np_oracle.equ_sweeps_redblack(A, X, B, 9, n_mid)
        np.testing.assert_array_equal(img, np_oracle.clip_u8(want))
        np.testing.assert_allclose(err, np_oracle.equ_residual_f64(A, want, B), rtol=1e-5, atol=1e-3)
        # thread-count independent (deterministic red-black)
        solver2 = core.EquSolver(1)
        solver2.partition(crop)
        solver2.reset(n, A, X, B)
        np.testing.assert_array_equal(solver2.step(9)[0], img)

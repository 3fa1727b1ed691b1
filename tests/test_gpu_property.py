"""Property-based parity (hypothesis): random shapes, masks, offsets, gradient modes and sweep counts
through the Processor API, GPU vs the numpy oracle -- fp32 state bit-exact, uint8 image identical."""

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import c_oracle, np_oracle

pytestmark = pytest.mark.gpu


@st.composite
def blends(draw):
    mh = draw(st.integers(3, 70))
    mw = draw(st.integers(3, 90))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    density = draw(st.sampled_from([0.15, 0.5, 0.9, 1.0]))
    mask = (rng.random((mh, mw)) < density).astype(np.uint8) * draw(st.sampled_from([128, 200, 255]))
    if draw(st.booleans()):
        mask = np.repeat(mask[:, :, None], 3, axis=2)
    pad_s = [draw(st.integers(0, 6)) for _ in range(4)]
    pad_t = [draw(st.integers(0, 6)) for _ in range(4)]
    src = rng.integers(0, 256, (mh + pad_s[0] + pad_s[1], mw + pad_s[2] + pad_s[3], 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, (mh + pad_t[0] + pad_t[1], mw + pad_t[2] + pad_t[3], 3), dtype=np.uint8)
    mode = draw(st.sampled_from(["max", "src", "avg"]))
    steps = draw(st.lists(st.integers(0, 23), min_size=1, max_size=3))
    return src, mask, tgt, (pad_s[0], pad_s[2]), (pad_t[0], pad_t[2]), mode, steps


def _has_unknowns(mask):
    try:
        np_oracle.canonical_mask(mask)
        return True
    except ValueError:
        return False


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(blends(), st.sampled_from(["grid", "equ", "equ-gather", "equ-redblack"]))
def test_processor_matches_oracle(blend, kind):
    import fpie_b200

    src, mask, tgt, off_s, off_t, mode, steps = blend
    if kind == "grid":
        proc, orc = fpie_b200.GridProcessor(mode, "b200"), np_oracle.GridOracle(mode)
    elif kind == "equ-redblack":
        proc, orc = fpie_b200.EquProcessor(mode, "b200", mode="redblack"), None
    else:
        proc = fpie_b200.EquProcessor(mode, "b200", mode="gather" if kind == "equ-gather" else "jacobi")
        orc = np_oracle.EquOracle(mode)
    if not _has_unknowns(mask):
        with pytest.raises(RuntimeError, match="empty"):
            proc.reset(src, mask, tgt, off_s, off_t)
        return
    n = proc.reset(src, mask, tgt, off_s, off_t)
    if orc is None:  # red-black: oracle system with the odd/even labelling
        m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
        ids, n_mid = np_oracle.partition_redblack(m_full[x0:x1, y0:y1])
        N, A, X, B, index = np_oracle.equ_system(src, mask, tgt, off_s, off_t, mode, ids=ids)
        assert n == N
        canvas = tgt.copy()
        for it in steps:
            out, err = proc.step(it)
            X = np_oracle.equ_sweeps_redblack(A, X, B, it, n_mid)
            np.testing.assert_array_equal(proc.core.state(), X)
            canvas[index] = np_oracle.clip_u8(X)[1:]
            np.testing.assert_array_equal(out, canvas)
        return
    assert n == orc.reset(src, mask, tgt, off_s, off_t)
    for it in steps:
        out, err = proc.step(it)
        wout, werr = orc.step(it)
        state = proc.core.state()
        np.testing.assert_array_equal(state, orc.t if kind == "grid" else orc.X)
        np.testing.assert_array_equal(out, wout)
        np.testing.assert_allclose(err, werr, rtol=1e-4, atol=1e-3)


# -- convergence-driven stepping --------------------------
def _oracle_solve(step_fn, err_fn, max_iters, check_every, tol):
    """What solve() is defined to do, on the oracle: check, then sweep in chunks."""
    done = 0
    while True:
        if err_fn().max() <= tol or done >= max_iters:
            return done
        s = min(check_every, max_iters - done)
        step_fn(s)
        done += s


@pytest.mark.parametrize("solver_kind", ["grid", "equ", "equ-gather", "equ-redblack"])
def test_solve_stops_where_the_oracle_stops(solver_kind):
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("circle", 150, 170, seed=11)
    check_every, max_iters = 25, 400
    if solver_kind == "grid":
        proc = fpie_b200.GridProcessor("max", "b200")
        ora = np_oracle.GridOracle("max")
        ora.reset(src, mask, tgt, (0, 0), (0, 0))
        sweep = lambda s: setattr(ora, "t", c_oracle.grid_sweeps(ora.mask, ora.t, ora.g, s))  # noqa: E731
        err_of = lambda: np_oracle.grid_residual_f64(ora.mask, ora.t, ora.g)  # noqa: E731
    else:
        mode = {"equ": "jacobi", "equ-gather": "gather", "equ-redblack": "redblack"}[solver_kind]
        proc = fpie_b200.EquProcessor("max", "b200", mode=mode)
    proc.reset(src, mask, tgt, (0, 0), (0, 0))
    if solver_kind != "grid":
        A, X, B = proc.core.system()
        state = {"x": X.copy()}
        if mode == "redblack":
            m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
            n_mid = np_oracle.partition_redblack(m_full[x0:x1, y0:y1])[1]
            sweep = lambda s: state.__setitem__("x", np_oracle.equ_sweeps_redblack(A, state["x"], B, s, n_mid))  # noqa: E731
        else:
            sweep = lambda s: state.__setitem__("x", c_oracle.equ_sweeps(A, state["x"], B, s))  # noqa: E731
        err_of = lambda: np_oracle.equ_residual_f64(A, state["x"], B)  # noqa: E731

    # residuals at every check point; pick tolerances between two of them so that rounding in the
    # last digit of err cannot move the stopping point
    errs = [err_of().max()]
    for _ in range(max_iters // check_every):
        sweep(check_every)
        errs.append(err_of().max())
    assert errs[3] > errs[6] > errs[10]
    for stop_at in (3, 6, 10):
        tol = float(np.sqrt(errs[stop_at] * errs[stop_at - 1]))
        proc.reset(src, mask, tgt, (0, 0), (0, 0))
        if solver_kind != "grid":
            state["x"] = X.copy()
        out, err, done = proc.solve(max_iters, tol, check_every)
        assert done == stop_at * check_every
        np.testing.assert_allclose(err.max(), errs[stop_at], rtol=1e-4)
        assert err.max() <= tol
        out2, err2 = proc.step(0)
        np.testing.assert_array_equal(out, out2)
    # never converging within the budget: runs exactly max_iters; zero budget: runs nothing
    proc.reset(src, mask, tgt, (0, 0), (0, 0))
    assert proc.solve(60, 0.0, 25)[2] == 60
    proc.reset(src, mask, tgt, (0, 0), (0, 0))
    assert proc.solve(0, 0.0, 25)[2] == 0
    with pytest.raises(RuntimeError):
        proc.core.solve(10, 1.0, 0)


@pytest.mark.parametrize("kind", ["grid", "equ"])
@pytest.mark.parametrize("where", ["small box in a large target", "box = the whole target"])
def test_canvas_is_the_target_outside_the_box_and_the_device_crop_inside(kind, where):
    """`reset` copies the caller's target only OUTSIDE the blend's bounding box (several threads for a large image,
    the box unknown when they start); the inside comes from the device.  Whatever the split: `proc.tgt` equals the
    target before any step (process.py:268 / 384 `self.tgt = tgt.copy()`), the oracle's image after, is the same
    buffer every time, and the caller's array can be overwritten as soon as `reset` has returned."""
    import fpie_b200
    from fpie_b200 import synth

    rng = np.random.default_rng(3)
    if where.startswith("small"):
        tgt = rng.integers(0, 256, (2100, 1500, 3), dtype=np.uint8)  # > 8 MB: the threaded copy
        src = rng.integers(0, 256, (300, 260, 3), dtype=np.uint8)
        mask = synth.make_mask("circle", 200, 180)
        on_src, on_tgt = (40, 30), (1234, 777)
    else:
        tgt = rng.integers(0, 256, (1800, 1700, 3), dtype=np.uint8)
        src = rng.integers(0, 256, (1800, 1700, 3), dtype=np.uint8)
        mask = np.full((1800, 1700), 255, np.uint8)
        on_src, on_tgt = (0, 0), (0, 0)
    Proc = fpie_b200.GridProcessor if kind == "grid" else fpie_b200.EquProcessor
    want = (np_oracle.GridOracle if kind == "grid" else np_oracle.EquOracle)("max")
    n_want = want.reset(src, mask, tgt, on_src, on_tgt)
    for round_ in range(2):  # (the second reset recycles the page-locked canvas of the first)
        proc = Proc("max", "b200") if round_ == 0 else proc
        mine = tgt.copy()
        assert proc.reset(src, mask, mine, on_src, on_tgt) == n_want
        mine[...] = 0  # the caller's buffer is its own again
        before = proc.tgt
        np.testing.assert_array_equal(before, tgt)
        out, _ = proc.step(33)
        assert out is before and out is proc.tgt
    want_out, _ = want.step(33)
    np.testing.assert_array_equal(out, want_out)

"""EquSolver parity on the GPU, through the C ABI.

Same tolerances as test_gpu_grid.py; the gather kernel uses numpy's add order
((((B + X[up]) + X[down]) + X[left]) + X[right]) / 4, so states are required
to be bit-exact against the oracle.
"""

import numpy as np
import pytest
from conftest import GOLDEN_CASES, MODES, golden_case

from oracle import c_oracle, np_oracle

pytestmark = pytest.mark.gpu

ERR_RTOL = 1e-4


def _solver(block=256):
    import fpie_b200

    return fpie_b200.EquSolver(block)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_partition_matches_numpy_core(golden, name):
    s = _solver()
    mask = golden[f"{name}/grid/max/mask_crop"]
    ids = s.partition(mask)
    assert ids.shape == mask.shape
    np.testing.assert_array_equal(ids, np_oracle.partition_rowmajor(mask))


@pytest.mark.parametrize("shape", [(1, 1), (1, 37), (64, 64), (63, 65), (300, 4099), (2048, 2051)])
def test_partition_scan_sizes(shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    mask = (rng.random(shape) < 0.6).astype(np.int32) * rng.integers(1, 5, shape).astype(np.int32)
    s = _solver()
    np.testing.assert_array_equal(s.partition(mask), np_oracle.partition_rowmajor(mask))
    # non-contiguous view (fpie/process.py:224)
    if shape[0] > 4 and shape[1] > 4:
        view = mask[1:-2, 2:-1]
        np.testing.assert_array_equal(s.partition(view), np_oracle.partition_rowmajor(view))
    # negative entries count as unmasked (mask > 0, fpie/core/cuda/equ.cu:46)
    neg = mask.copy()
    neg[neg == 2] = -1
    np.testing.assert_array_equal(s.partition(neg), np_oracle.partition_rowmajor(neg))


@pytest.mark.parametrize("block", [32, 256, 1024, 100256])  # 100000 + z forces the generic int4 table
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_core_matches_reference_golden(golden, name, block):
    c = golden_case(golden, name)
    s = _solver(block)
    for mode in MODES:
        key = f"{name}/equ/{mode}"
        A, X, B = golden[f"{key}/A"], golden[f"{key}/X0"], golden[f"{key}/B"]
        s.reset(A.shape[0], A, X, B)
        for si, it in enumerate(c["steps"]):
            img, err = s.step(it)
            want = golden[f"{key}/state{si}"]
            np.testing.assert_array_equal(s.state(), want)
            np.testing.assert_array_equal(img, np_oracle.clip_u8(want))
            assert img.shape == (A.shape[0], 3) and img.dtype == np.uint8 and err.dtype == np.float32
            np.testing.assert_allclose(err, golden[f"{key}/err{si}"], rtol=ERR_RTOL, atol=1e-3)
            np.testing.assert_allclose(err, np_oracle.equ_residual_f64(A, want, B), rtol=ERR_RTOL, atol=1e-3)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("mode", MODES)
def test_processor_matches_reference_golden(golden, name, mode):
    """EquProcessor.reset on the device: scan + A/X/B build + paste."""
    import fpie_b200

    c = golden_case(golden, name)
    proc = fpie_b200.EquProcessor(mode, "b200")
    n = proc.reset(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"])
    key = f"{name}/equ/{mode}"
    assert n == int(golden[f"{key}/n"])
    A, X, B = proc.core.system()
    np.testing.assert_array_equal(A, golden[f"{key}/A"])
    np.testing.assert_array_equal(X, golden[f"{key}/X0"])
    np.testing.assert_array_equal(B, golden[f"{key}/B"])
    first = None
    for si, it in enumerate(c["steps"]):
        out, err = proc.step(it)
        first = out if first is None else first
        assert out is first
        np.testing.assert_array_equal(out, golden[f"{key}/img{si}"])
        np.testing.assert_array_equal(proc.core.state(), golden[f"{key}/state{si}"])
        np.testing.assert_allclose(err, golden[f"{key}/err{si}"], rtol=ERR_RTOL, atol=1e-3)


@pytest.mark.parametrize("kind,size", [("ring", 300), ("star", 511), ("holes", 257), ("circle", 1024), ("square", 640)])
def test_synthetic_masks_vs_c_oracle(kind, size):
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem(kind, size, size + 3, seed=9)
    mode = MODES[size % 3]
    n, A, X, B, index = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), mode)
    # core-level: oracle-built system through reset(N, A, X, B)
    s = _solver()
    s.reset(n, A, X, B)
    img, err = s.step(60)
    want = c_oracle.equ_sweeps(A, X, B, 60)
    np.testing.assert_array_equal(s.state(), want)
    np.testing.assert_array_equal(img, c_oracle.clip_u8(want))
    e32, e64 = c_oracle.equ_residual(A, want, B)
    np.testing.assert_allclose(err, e64, rtol=ERR_RTOL, atol=1e-3)
    np.testing.assert_allclose(err, e32, rtol=1e-3, atol=1e-3)  # the oracle's sequential fp32 sum drifts by itself
    # processor-level: device-built system is identical, pasted image too
    proc = fpie_b200.EquProcessor(mode, "b200")
    assert proc.reset(src, mask, tgt, (0, 0), (0, 0)) == n
    A2, X2, B2 = proc.core.system()
    np.testing.assert_array_equal(A2, A)
    np.testing.assert_array_equal(X2, X)
    np.testing.assert_array_equal(B2, B)
    out, err2 = proc.step(60)
    canvas = tgt.copy()
    canvas[index] = c_oracle.clip_u8(want)[1:]
    np.testing.assert_array_equal(out, canvas)
    np.testing.assert_allclose(err2, e64, rtol=ERR_RTOL, atol=1e-3)


def test_partition_through_reference_flow():
    """The exact call sequence of EquProcessor.mask2index (process.py:180-190) with our partition."""
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("star", 200, 240, seed=2)
    m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
    crop = m_full[x0:x1, y0:y1]  # non-contiguous view
    s = _solver()
    ids = s.partition(crop)
    n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "avg", ids=ids)
    n2, A2, X2, B2, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "avg")
    assert n == n2
    np.testing.assert_array_equal(A, A2)
    np.testing.assert_array_equal(B, B2)


def test_step_calls_accumulate_and_reset_reuses_solver(golden):
    s = _solver()
    for name in ("ring_off", "holes_full"):
        A, X, B = (golden[f"{name}/equ/max/{k}"] for k in ("A", "X0", "B"))
        s.reset(A.shape[0], A, X, B)
        s.step(5)
        s.step(0)
        s.step(8)
        a = s.state()
        s.reset(A.shape[0], A, X, B)
        s.step(13)
        np.testing.assert_array_equal(a, s.state())
        np.testing.assert_array_equal(a, np_oracle.equ_sweeps(A, X, B, 13))


def test_edge_and_error_behaviour():
    import fpie_b200

    s = _solver()
    with pytest.raises(RuntimeError):
        s.step(1)
    # N = 1: only the constant row
    s.reset(1, np.zeros((1, 4), np.int32), np.zeros((1, 3), np.float32), np.zeros((1, 3), np.float32))
    img, err = s.step(3)
    assert img.shape == (1, 3) and (img == 0).all() and (err == 0).all()
    # out-of-range neighbour index is rejected (the reference would read out of bounds)
    A = np.zeros((4, 4), np.int32)
    A[2, 1] = 4
    with pytest.raises(RuntimeError, match="outside"):
        s.reset(4, A, np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32))
    A[2, 1] = -1
    with pytest.raises(RuntimeError, match="outside"):
        s.reset(4, A, np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32))
    with pytest.raises(ValueError):
        s.reset(4, np.zeros((4, 3), np.int32), np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32))
    p = fpie_b200.EquProcessor("src", "b200")
    z = np.zeros((10, 10, 3), np.uint8)
    with pytest.raises(RuntimeError, match="empty"):
        p.reset(z, np.zeros((10, 10), np.uint8), z, (0, 0), (0, 0))
    with pytest.raises(RuntimeError, match="outside the target"):
        p.reset(z, np.full((10, 10), 255, np.uint8), z, (0, 0), (0, 5))
    with pytest.raises(ValueError):
        fpie_b200.EquProcessor("median", "b200")


def test_large_irregular_mask():
    """2048^2 ring (about 2.2 M unknowns), 40 sweeps, bit-exact vs the C restatement."""
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("ring", 2048, 2048, seed=1)
    proc = fpie_b200.EquProcessor("avg", "b200")
    n = proc.reset(src, mask, tgt, (0, 0), (0, 0))
    A, X, B = proc.core.system()
    out, err = proc.step(40)
    want = c_oracle.equ_sweeps(A, X, B, 40)
    np.testing.assert_array_equal(proc.core.state(), want)
    assert n == A.shape[0] and n > 2_000_000
    np.testing.assert_allclose(err, c_oracle.equ_residual(A, want, B)[1], rtol=ERR_RTOL)
    n0, A0, X0, B0, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "avg")
    np.testing.assert_array_equal(A, A0)
    np.testing.assert_array_equal(B, B0)


def test_config3_full_size_ring():
    """BASELINE config 3 size: 8192^2 ring mask (about 35 M unknowns), grad avg.  The device-built
    system must equal the oracle's, and 12 gather sweeps must match the C restatement bit for bit."""
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("ring", 8192, 8192, seed=0)
    proc = fpie_b200.EquProcessor("avg", "b200", mode="gather")  # the index-mapped gather path named by config 3
    n = proc.reset(src, mask, tgt, (0, 0), (0, 0))
    assert n > 35_000_000 and proc.core.info()["path"] == "gather-compact"
    A, X, B = proc.core.system()
    # ids are row-major: left / right neighbours are i-1 / i+1 wherever they are masked
    assert ((A[1:, 2] == 0) | (A[1:, 2] == np.arange(1, n) - 1)).all()
    assert ((A[1:, 3] == 0) | (A[1:, 3] == np.arange(1, n) + 1)).all()
    out, err = proc.step(12)
    want = c_oracle.equ_sweeps(A, X, B, 12)
    np.testing.assert_array_equal(proc.core.state(), want)
    np.testing.assert_allclose(err, c_oracle.equ_residual(A, want, B)[1], rtol=ERR_RTOL)
    # the scatter put every solved pixel where the mask is and left the rest of the target alone
    m_full, _ = np_oracle.canonical_mask(mask)
    assert np.array_equal(out[m_full == 0], tgt[m_full == 0])
    assert np.array_equal(out[m_full != 0], c_oracle.clip_u8(want)[1:])


@pytest.mark.parametrize("name,mode", [("rng24", "max"), ("ring_off", "avg"), ("holes_full", "src"), ("disk_sat", "max")])
def test_redblack_mode_matches_reference_openmp_scheme(golden, name, mode):
    """mode="redblack": the reference OpenMP EquSolver's deterministic red-black Gauss-Seidel
    (openmp/equ.cc:22-56, 107-118) -- ids from the device partition, then bit-exact sweeps."""
    import fpie_b200

    c = golden_case(golden, name)
    m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(c["mask"])
    crop = m_full[x0:x1, y0:y1]
    s = fpie_b200.EquSolver(256, mode="redblack")
    ids = s.partition(crop)
    want_ids, n_mid = np_oracle.partition_redblack(crop)
    np.testing.assert_array_equal(ids, want_ids)
    n, A, X, B, index = np_oracle.equ_system(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"], mode, ids=ids)
    s.reset(n, A, X, B)
    s.step(4)
    img, err = s.step(7)
    want = np_oracle.equ_sweeps_redblack(A, X, B, 11, n_mid)
    np.testing.assert_array_equal(s.state(), want)
    np.testing.assert_array_equal(img, np_oracle.clip_u8(want))
    np.testing.assert_allclose(err, np_oracle.equ_residual_f64(A, want, B), rtol=ERR_RTOL, atol=1e-3)
    ref = c_oracle.load_reference_core("core_openmp")
    if ref is not None:  # the compiled reference itself, when oracle/_ref travelled to this box
        r = ref.EquSolver(4)
        r.partition(np.ascontiguousarray(crop))
        r.reset(n, A, X, B)
        rimg, rerr = r.step(11)
        np.testing.assert_array_equal(img, rimg)
        np.testing.assert_allclose(err, rerr, rtol=ERR_RTOL, atol=1e-3)
    # processor level (device-side red-black labelling + build + paste)
    proc = fpie_b200.EquProcessor(mode, "b200", mode="redblack")
    assert proc.reset(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"]) == n
    A2, X2, B2 = proc.core.system()
    np.testing.assert_array_equal(A2, A)
    np.testing.assert_array_equal(B2, B)
    out, err2 = proc.step(11)
    canvas = c["tgt"].copy()
    canvas[index] = np_oracle.clip_u8(want)[1:]
    np.testing.assert_array_equal(out, canvas)


def test_redblack_large_and_errors():
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("star", 700, 900, seed=4)
    proc = fpie_b200.EquProcessor("max", "b200", mode="redblack")
    n = proc.reset(src, mask, tgt, (0, 0), (0, 0))
    A, X, B = proc.core.system()
    m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
    _, n_mid = np_oracle.partition_redblack(m_full[x0:x1, y0:y1])
    proc.step(30)
    np.testing.assert_array_equal(proc.core.state(), np_oracle.equ_sweeps_redblack(A, X, B, 30, n_mid))
    with pytest.raises(ValueError):
        fpie_b200.EquSolver(256, mode="sor")
    s = fpie_b200.EquSolver(256, mode="redblack")
    s.reset(n, A, X, B)  # no partition() on this solver: the colour split is unknown
    with pytest.raises(RuntimeError, match="partition"):
        s.step(1)


def test_unstructured_ids_use_the_generic_table(golden):
    """Any bijection of the masked pixels onto 1..K is a legal labelling (process.py:187-190).  A random
    relabelling breaks the left = i-1 / right = i+1 pattern, so the compact table must not be used --
    and the result must still be the relabelled Jacobi iterate, bit for bit."""
    A, X, B = (golden[f"holes_full/equ/max/{k}"] for k in ("A", "X0", "B"))
    n = A.shape[0]
    rng = np.random.default_rng(5)
    perm = np.concatenate([[0], 1 + rng.permutation(n - 1)])  # new id of old id i
    inv = np.argsort(perm)
    A2 = perm[A][inv].astype(np.int32)
    X2, B2 = X[inv], B[inv]
    s = _solver()
    s.reset(n, A2, X2, B2)
    s.step(21)
    want = np_oracle.equ_sweeps(A, X, B, 21)
    np.testing.assert_array_equal(s.state(), want[inv])
    # and the structured path on the original labelling gives the same values
    s.reset(n, A, X, B)
    s.step(21)
    np.testing.assert_array_equal(s.state(), want)


@pytest.mark.parametrize("kind,mode", [("ring", "avg"), ("star", "max"), ("holes", "src")])
def test_promotion_to_tiled_kernel_is_bit_exact(kind, mode):
    """The reference call sequence (partition -> reset -> step, process.py:187, 270, 275): when A is the
    4-neighbour structure of the mask this solver labelled, EquSolver runs the temporally blocked grid
    kernel; mode="gather" keeps the index-mapped kernels.  Both must give the same bits as the oracle."""
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem(kind, 420, 380, seed=3)
    m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
    crop = m_full[x0:x1, y0:y1]
    results = {}
    for solver_mode, want_path in (("jacobi", "tiled"), ("gather", "gather-compact")):
        s = fpie_b200.EquSolver(256, mode=solver_mode)
        ids = s.partition(crop)
        n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), mode, ids=ids)
        s.reset(n, A, X, B)
        assert s.info()["path"] == want_path
        s.step(17)
        img, err = s.step(26)
        results[solver_mode] = (s.state(), img, err)
    want = c_oracle.equ_sweeps(A, X, B, 43)
    for state, img, err in results.values():
        np.testing.assert_array_equal(state, want)
        np.testing.assert_array_equal(img, c_oracle.clip_u8(want))
        np.testing.assert_allclose(err, c_oracle.equ_residual(A, want, B)[1], rtol=ERR_RTOL)
    np.testing.assert_array_equal(results["jacobi"][2], results["gather"][2])  # same residual kernel, same bits


def test_promotion_is_refused_when_the_system_does_not_match_the_mask(golden):
    import fpie_b200

    c = golden_case(golden, "holes_full")
    m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(c["mask"])
    crop = m_full[x0:x1, y0:y1]
    n, A, X, B, _ = np_oracle.equ_system(c["src"], c["mask"], c["tgt"], c["off_src"], c["off_tgt"], "max")
    s = fpie_b200.EquSolver(256)
    # (1) labelled a different mask with the same number of unknowns: transposed crop
    s.partition(np.ascontiguousarray(crop.T))
    s.reset(n, A, X, B)
    assert s.info()["path"] != "tiled"
    s.step(9)
    np.testing.assert_array_equal(s.state(), np_oracle.equ_sweeps(A, X, B, 9))
    # (2) right mask, but one neighbour link cut (still a legal system): must stay on the gather path
    s.partition(crop)
    A2 = A.copy()
    row = np.flatnonzero(A2[:, 0] > 0)[5]
    A2[row, 0] = 0
    s.reset(n, A2, X, B)
    assert s.info()["path"] != "tiled"
    s.step(9)
    np.testing.assert_array_equal(s.state(), np_oracle.equ_sweeps(A2, X, B, 9))
    # (3) non-zero constant row: the grid embedding does not apply
    X3 = X.copy()
    X3[0] = 1.0
    s.reset(n, A, X3, B)
    assert s.info()["path"] != "tiled"
    s.step(5)
    np.testing.assert_array_equal(s.state(), np_oracle.equ_sweeps(A, X3, B, 5))
    # (4) the matching system is promoted, and a later partition() on the same solver demotes it safely
    s.reset(n, A, X, B)
    assert s.info()["path"] == "tiled"
    s.step(6)
    s.partition(np.ascontiguousarray(crop[:, ::-1]))
    s.step(3)
    np.testing.assert_array_equal(s.state(), np_oracle.equ_sweeps(A, X, B, 9))


@pytest.mark.parametrize("generic", [False, True])
def test_gather_long_runs_replay_a_cuda_graph(generic, monkeypatch):
    """BASELINE config 1 (1026^2 square, 1 048 576 unknowns, L2-resident) on the index-mapped gather path:
    long runs replay a captured graph of 64 sweeps chained by programmatic dependent launch.  Same bits as the C
    restatement, as the same run without graphs, and across repeated step / reset calls."""
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("square", 1026, 1026, seed=0)
    n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "max")
    if generic:  # a legal relabelling (ids reversed) defeats the compact table: the generic int4 kernel
        perm = np.concatenate([[0], np.arange(n - 1, 0, -1)]).astype(np.int32)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(n, dtype=np.int32)
        A, X, B = inv[A[perm]], X[perm], B[perm]
    s = fpie_b200.EquSolver(256, mode="gather")
    s.reset(n, A, X, B)
    assert s.info()["path"] == ("gather-int4" if generic else "gather-compact")
    s.step(150)  # warm launches + two graph replays + a tail
    img, err = s.step(211)
    want = c_oracle.equ_sweeps(A, X, B, 361)
    np.testing.assert_array_equal(s.state(), want)
    np.testing.assert_array_equal(img, c_oracle.clip_u8(want))
    s.reset(n, A, X, B)  # buffers may move: the graphs must be rebuilt, not replayed
    s.step(361)
    np.testing.assert_array_equal(s.state(), want)
    monkeypatch.setenv("FPIE_B200_NO_GRAPH", "1")
    plain = fpie_b200.EquSolver(256, mode="gather")
    plain.reset(n, A, X, B)
    plain.step(361)
    np.testing.assert_array_equal(plain.state(), want)


def test_gather_table_forms_are_bit_identical(monkeypatch):
    """The structured gather path streams 34 bytes per unknown and sweep (4-byte distance table + fp16 B) when the
    system allows it, 40 (fp32 B) when some B is not exactly a half, 44 (8-byte table) when a neighbour is further
    than 32767 ids away -- all with the reference's operands and add order: same bits."""
    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("ring", 700, 900, seed=4)
    n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "avg")
    want = c_oracle.equ_sweeps(A, X, B, 77)
    # by default only systems beyond the L2 (>= 2^21 unknowns) take the 4-byte table: small ones are faster with one
    # unknown per thread
    d = fpie_b200.EquSolver(256, mode="gather")
    d.reset(n, A, X, B)
    assert d.info()["table"] == "int2 + fp32 B"
    monkeypatch.setenv("FPIE_B200_DELTA16_MIN", "0")
    s = fpie_b200.EquSolver(256, mode="gather")
    s.reset(n, A, X, B)
    assert s.info()["path"] == "gather-compact" and s.info()["table"] == "delta16 + fp16 B"
    s.step(77)
    np.testing.assert_array_equal(s.state(), want)
    # the one-pass form of the same kernel (A/B switch; the default is persistent and software-pipelined)
    monkeypatch.setenv("FPIE_B200_D16_PIPE", "0")
    one = fpie_b200.EquSolver(256, mode="gather")
    one.reset(n, A, X, B)
    assert one.info()["table"] == "delta16 + fp16 B"
    one.step(77)
    np.testing.assert_array_equal(one.state(), want)
    monkeypatch.setenv("FPIE_B200_D16_PIPE", "1")  # one CTA per SM: every thread strides over many chunks
    few = fpie_b200.EquSolver(256, mode="gather")
    few.reset(n, A, X, B)
    few.step(77)
    np.testing.assert_array_equal(few.state(), want)
    monkeypatch.delenv("FPIE_B200_D16_PIPE")
    # B that is not representable in fp16 (a legal system: B is just numbers): the fp32 stream, same bits
    B2 = B.copy()
    B2[5, 1] += np.float32(0.001)
    B2[n // 2, 2] = np.float32(70000.25)
    s.reset(n, A, X, B2)
    assert s.info()["table"] == "delta16 + fp32 B"
    s.step(40)
    np.testing.assert_array_equal(s.state(), c_oracle.equ_sweeps(A, X, B2, 40))
    # the 8-byte table when asked for (and the distance check: a system whose up-neighbours are far away)
    monkeypatch.setenv("FPIE_B200_NO_DELTA16", "1")
    t = fpie_b200.EquSolver(256, mode="gather")
    t.reset(n, A, X, B)
    assert t.info()["table"] == "int2 + fp32 B"
    t.step(77)
    np.testing.assert_array_equal(t.state(), want)
    monkeypatch.delenv("FPIE_B200_NO_DELTA16")
    wide_src, wide_mask, wide_tgt = synth.make_problem("square", 6, 40000, seed=1)  # rows of 39998 unknowns
    nw, Aw, Xw, Bw, _ = np_oracle.equ_system(wide_src, wide_mask, wide_tgt, (0, 0), (0, 0), "max")
    w = fpie_b200.EquSolver(256, mode="gather")
    w.reset(nw, Aw, Xw, Bw)
    assert w.info()["table"] == "int2 + fp32 B"
    w.step(9)
    np.testing.assert_array_equal(w.state(), c_oracle.equ_sweeps(Aw, Xw, Bw, 9))

"""Host logic of the row-band sharded GridSolver on the CPU: band plan,
halo exchange over ``gloo`` (world_size 2 and 3, real processes), global err
all-reduce.  The per-rank compute is the numpy oracle (OracleBandCore), so what
is tested is exactly the orchestration that the GPU path reuses with
``fpie_b200.GridSolver`` + NCCL."""

import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp
from conftest import PKG_ROOT, ROOT

from oracle import np_oracle


def test_band_offsets_follow_the_reference_rule():
    from fpie_b200 import band

    assert band.band_offsets(10, 3) == [0, 4, 7, 10]  # mpi/grid.cc:27-31
    assert band.band_offsets(8, 8) == list(range(9))
    assert band.band_offsets(97, 4) == np_oracle.band_offsets(97, 4)


def test_plan_geometry():
    from fpie_b200 import band

    p = band.make_plan(100, 4, 0, 8)
    assert (p.band_lo, p.band_hi, p.slab_lo, p.slab_hi) == (0, 25, 0, 33) and p.up is None and p.down == 1
    p = band.make_plan(100, 4, 2, 8)
    assert (p.band_lo, p.band_hi, p.slab_lo, p.slab_hi) == (50, 75, 42, 83) and (p.up, p.down) == (1, 3)
    assert p.local_band == (8, 33) and p.slab_rows == 41
    p = band.make_plan(100, 4, 3, 8)
    assert p.slab_hi == 100 and p.down is None
    with pytest.raises(ValueError):
        band.make_plan(20, 4, 1, 8)  # 5-row bands cannot carry an 8-row halo
    with pytest.raises(ValueError):
        band.make_plan(20, 2, 0, 0)
    one = band.make_plan(7, 1, 0, 16)
    assert (one.slab_lo, one.slab_hi, one.up, one.down) == (0, 7, None, None)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, halo, steps, shape, out_dir, overlap=True, density=0.65):
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from band_helpers import OracleBandCore, random_grid

    from fpie_b200 import band

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mask, tgt, grad = random_grid(*shape, seed=5, density=density)
        solver = band.BandGridSolver(OracleBandCore(block_k=4), dist, halo=halo, overlap=overlap)
        solver.reset(mask.size, mask, tgt, grad)
        assert solver.exchange_overlaps == overlap  # split passes + exchange beside the interior, or plain
        solver.sync()
        errs = []
        for it in steps:
            img, err = solver.step(it)
            errs.append(err)
        p = solver.plan
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=solver.band_state(), img=img, err=np.array(errs),
                 lo=p.band_lo, hi=p.band_hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo,steps,overlap,density", [
    (2, 4, (37,), True, 1.0), (2, 8, (5, 20, 12), True, 1.0), (2, 16, (5, 20, 12), True, 1.0), (3, 7, (30, 7), True, 1.0),
    (2, 16, (5, 20, 12), False, 1.0),
    (3, 3, (11,), True, 0.65), (2, 16, (40,), True, 0.65)])
def test_gloo_bands_reproduce_global_jacobi(tmp_path, world, halo, steps, overlap, density):
    """Row bands over gloo, with the overlapped schedule (halo 4 = one pass per interval, 16 = four
    passes: first / middle / last, 8 = first / last, 7 and 3 = a short last pass) and with the plain
    one.  The oracle-backed core completes a scoped exchange only at ``wait_exchange`` (the latest the
    GPU could), and the fully masked grids with shallow halos are the sensitive cases: a stale halo row
    moves one row per sweep and fades 4x per row, so an exchange started one pass too early shows at
    any depth and one joined a pass too late at halo 8 (both checked by mutating the schedule)."""
    shape = (97, 83)
    mp.spawn(_worker, args=(world, _free_port(), halo, steps, shape, str(tmp_path), overlap, density), nprocs=world,
             join=True)
    from band_helpers import random_grid

    mask, tgt, grad = random_grid(*shape, seed=5, density=density)
    want = np_oracle.grid_sweeps(mask, tgt, grad, sum(steps))
    got = np.zeros_like(want)
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        got[int(z["lo"]) : int(z["hi"])] = z["state"]
        covered += int(z["hi"]) - int(z["lo"])
        np.testing.assert_array_equal(z["img"], np_oracle.clip_u8(want[int(z["lo"]) : int(z["hi"])]))
        # err is global and identical on every rank
        np.testing.assert_allclose(z["err"][-1], np_oracle.grid_residual_f64(mask, want, grad), rtol=1e-6)
    assert covered == shape[0]
    np.testing.assert_array_equal(got, want)  # bit-identical to single-domain Jacobi


def _proc_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from band_helpers import OracleBandCore

    from fpie_b200 import band, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        src, mask, tgt = synth.make_problem("circle", 120, 90, seed=6)
        big_tgt = np.random.default_rng(1).integers(0, 256, (150, 130, 3), dtype=np.uint8)
        proc = band.BandGridProcessor("avg", OracleBandCore(), dist, halo=6)
        n = proc.reset(src, np.repeat(mask[:, :, None], 3, 2), big_tgt, (0, 0), (20, 30))
        proc.sync()
        proc.step(10)
        res = proc.step(13)
        if rank == 0:
            out, err = res
            np.savez(os.path.join(out_dir, "proc.npz"), out=out, err=err, n=n)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def test_gloo_band_processor_matches_single_processor(tmp_path):
    """Image-level sharded run == the oracle's single-domain GridProcessor run (uint8 image, err, n)."""
    mp.spawn(_proc_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("circle", 120, 90, seed=6)
    big_tgt = np.random.default_rng(1).integers(0, 256, (150, 130, 3), dtype=np.uint8)
    want = np_oracle.GridOracle("avg")
    n = want.reset(src, mask, big_tgt, (0, 0), (20, 30))
    want.step(10)
    wout, werr = want.step(13)
    z = np.load(tmp_path / "proc.npz")
    assert int(z["n"]) == n
    np.testing.assert_array_equal(z["out"], wout)
    np.testing.assert_allclose(z["err"], werr, rtol=1e-5)


def _equ_proc_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from band_helpers import OracleBandCore

    from fpie_b200 import band, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        src, mask, tgt = synth.make_problem("holes", 110, 96, seed=9)
        big_tgt = np.random.default_rng(2).integers(0, 256, (140, 120, 3), dtype=np.uint8)
        proc = band.BandEquProcessor("max", OracleBandCore(), dist, halo=5)
        n = proc.reset(src, mask, big_tgt, (0, 0), (17, 11))
        proc.sync()
        proc.step(9)
        res = proc.step(14)
        state = proc.solver.band_state()
        p = proc.solver.plan
        np.savez(os.path.join(out_dir, f"equ{rank}.npz"), state=state, lo=p.band_lo, hi=p.band_hi, n=n,
                 out=res[0] if rank == 0 else 0, err=res[1] if rank == 0 else 0)
    finally:
        dist.destroy_process_group()


def test_gloo_band_equ_processor_matches_single_equ_processor(tmp_path):
    """Row-band sharded EquSolver arithmetic == the oracle's single-domain EquProcessor: same fp32
    unknowns (bit for bit), same uint8 image, same N, err within tolerance."""
    world = 3
    mp.spawn(_equ_proc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("holes", 110, 96, seed=9)
    big_tgt = np.random.default_rng(2).integers(0, 256, (140, 120, 3), dtype=np.uint8)
    want = np_oracle.EquOracle("max")
    n = want.reset(src, mask, big_tgt, (0, 0), (17, 11))
    want.step(9)
    wout, werr = want.step(14)
    m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
    crop = m_full[x0:x1, y0:y1]
    ids = np_oracle.partition_rowmajor(crop)
    z0 = np.load(tmp_path / "equ0.npz")
    assert int(z0["n"]) == n
    np.testing.assert_array_equal(z0["out"], wout)
    np.testing.assert_allclose(z0["err"], werr, rtol=1e-4)
    for r in range(world):
        z = np.load(tmp_path / f"equ{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        on = crop[lo:hi] > 0
        np.testing.assert_array_equal(z["state"][on], want.X[ids[lo:hi][on]])  # unknowns, bit for bit
        assert not z["state"][~on].any()  # everything else is the constant 0


def test_canonical_crop_matches_oracle():
    from fpie_b200 import band, synth

    for kind in ("circle", "star", "holes", "square"):
        mask = synth.make_mask(kind, 61, 77, seed=2)
        got_mask, box = band.canonical_crop(np.repeat(mask[:, :, None], 3, 2))
        m_full, wbox = np_oracle.canonical_mask(mask)
        assert box == wbox
        np.testing.assert_array_equal(got_mask > 0, m_full[wbox[0] : wbox[1], wbox[2] : wbox[3]] > 0)
    with pytest.raises(RuntimeError, match="empty"):
        band.canonical_crop(np.zeros((9, 9), np.uint8))

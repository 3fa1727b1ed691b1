"""Row-band sharded GridSolver on the GPU.

* one GPU: the bands live in threads of this process (ThreadDist), each with
  its own ``fpie_b200.GridSolver`` slab; halo rows move as device copies through
  the same ``rows_view`` tensors NCCL would use;
* >= 2 GPUs: real ``torch.distributed`` NCCL ranks, one process per GPU.

In both cases the stitched fp32 state must equal single-solver Jacobi bit for
bit, and the all-reduced err must equal the global residual."""

import os
import socket
import sys
import threading

import numpy as np
import pytest
from band_helpers import ThreadDist, random_grid
from conftest import PKG_ROOT, ROOT

from oracle import c_oracle, np_oracle

pytestmark = pytest.mark.gpu


def _run_threads(world, halo, make_and_reset, steps):
    import torch

    import fpie_b200
    from fpie_b200 import band

    dist = ThreadDist(world)
    out = [None] * world
    errors = []

    def work(rank):
        try:
            dist.bind(rank)
            torch.cuda.set_device(0)
            solver = band.BandGridSolver(band.CudaBandCore(fpie_b200.GridSolver(8, 8, device=0)), dist, halo=halo)
            make_and_reset(solver)
            solver.sync()
            for it in steps:
                img, err = solver.step(it)
            out[rank] = (solver.plan, solver.band_state(), img, err)
        except Exception as exc:  # surface in the main thread
            errors.append(exc)
            try:
                dist.bar.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out


@pytest.mark.parametrize("world,halo,steps", [(2, 8, (37,)), (3, 16, (5, 40)), (4, 24, (64,))])
def test_thread_bands_core_level(world, halo, steps):
    shape = (530, 301)
    mask, tgt, grad = random_grid(*shape, seed=8)
    out = _run_threads(world, halo, lambda s: s.reset(mask.size, mask, tgt, grad), steps)
    want = c_oracle.grid_sweeps(mask, tgt, grad, sum(steps))
    got = np.zeros_like(want)
    for plan, state, img, err in out:
        got[plan.band_lo : plan.band_hi] = state
        np.testing.assert_array_equal(img, c_oracle.clip_u8(want[plan.band_lo : plan.band_hi]))
        np.testing.assert_allclose(err, c_oracle.grid_residual(mask, want, grad)[1], rtol=1e-5)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("kind", ["square", "circle"])
def test_thread_bands_from_image_slabs(kind):
    """Each band uploads only its slab of the uint8 images (reset_slab)."""
    from fpie_b200 import band, synth

    n, m, world, halo, iters = 700, 640, 3, 16, 48
    src, mask, tgt = synth.make_problem(kind, n, m, seed=3)
    # global canonical crop (host side, cheap): the slabs are cut from it
    mc, tcrop, g, box = np_oracle.grid_system(src, mask, tgt, (0, 0), (0, 0), "max")
    x0, x1, y0, y1 = box
    csrc, cmask, ctgt = src[x0:x1, y0:y1], (mc * 255).astype(np.uint8), tgt[x0:x1, y0:y1]

    def reset(solver):
        p = band.make_plan(x1 - x0, solver.world, solver.rank, halo)
        sl = slice(p.slab_lo, p.slab_hi)
        solver.reset_slab(x1 - x0, csrc[sl], cmask[sl], ctgt[sl], "max")

    out = _run_threads(world, halo, reset, (iters,))
    want = c_oracle.grid_sweeps(mc, tcrop, g, iters)
    for plan, state, img, err in out:
        np.testing.assert_array_equal(state, want[plan.band_lo : plan.band_hi])
        np.testing.assert_allclose(err, c_oracle.grid_residual(mc, want, g)[1], rtol=1e-5)


def test_thread_band_processor_image_level():
    """BandGridProcessor: full images in on every rank, blended target out on rank 0."""
    import torch

    import fpie_b200
    from fpie_b200 import band, synth

    src, mask, tgt = synth.make_problem("star", 600, 500, seed=12)
    big_tgt = np.random.default_rng(4).integers(0, 256, (700, 640, 3), dtype=np.uint8)
    world = 3
    dist = ThreadDist(world)
    results, errors = [None] * world, []

    def work(rank):
        try:
            dist.bind(rank)
            torch.cuda.set_device(0)
            proc = band.BandGridProcessor("max", band.CudaBandCore(fpie_b200.GridSolver(8, 8, device=0)), dist, halo=16)
            n = proc.reset(src, mask, big_tgt, (0, 0), (40, 70))
            proc.sync()
            proc.step(20)
            results[rank] = (n, proc.step(31))
        except Exception as exc:
            errors.append(exc)
            try:
                dist.bar.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    want = np_oracle.GridOracle("max")
    n = want.reset(src, mask, big_tgt, (0, 0), (40, 70))
    want.t = c_oracle.grid_sweeps(want.mask, want.t, want.g, 51)
    wout, werr = want.step(0)
    assert results[0][0] == n and results[1][1] is None and results[2][1] is None
    out, err = results[0][1]
    np.testing.assert_array_equal(out, wout)
    np.testing.assert_allclose(err, werr, rtol=1e-4)


@pytest.mark.parametrize("kind,world,halo", [("star", 3, 16), ("holes", 2, 24), ("ring", 4, 8)])
def test_thread_band_equ_processor_matches_single_gpu_equ_solver(kind, world, halo):
    """BandEquProcessor (EquSolver arithmetic on row bands) == EquProcessor on one GPU: same unknowns
    bit for bit, same blended image, same N; err within tolerance."""
    import torch

    import fpie_b200
    from fpie_b200 import band, synth

    src, mask, tgt = synth.make_problem(kind, 610, 540, seed=5)
    big_tgt = np.random.default_rng(8).integers(0, 256, (700, 640, 3), dtype=np.uint8)
    single = fpie_b200.EquProcessor("max", "b200")
    n = single.reset(src, mask, big_tgt, (0, 0), (31, 52))
    single.step(30)
    wout, werr = single.step(45)
    wout = wout.copy()
    wstate = single.core.state()
    dist = ThreadDist(world)
    results, errors = [None] * world, []

    def work(rank):
        try:
            dist.bind(rank)
            torch.cuda.set_device(0)
            proc = band.BandEquProcessor("max", band.CudaBandCore(fpie_b200.GridSolver(8, 8, device=0)), dist, halo=halo)
            got_n = proc.reset(src, mask, big_tgt, (0, 0), (31, 52))
            proc.sync()
            proc.step(30)
            res = proc.step(45)
            results[rank] = (got_n, res, proc.solver.plan, proc.solver.band_state())
        except Exception as exc:
            errors.append(exc)
            try:
                dist.bar.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
    crop = m_full[x0:x1, y0:y1]
    ids = np_oracle.partition_rowmajor(crop)
    assert results[0][0] == n
    out, err = results[0][1]
    np.testing.assert_array_equal(out, wout)
    # random holes leave tiny domains that converge to rounding noise within 75 sweeps; the residual is
    # then a sum of last-bit residues, which the two expression orders (|4t-g-U-D-L-R| on the grid,
    # |B+sum(X[A])-4X| in the EquSolver) round differently
    np.testing.assert_allclose(err, werr, rtol=1e-4 if kind != "holes" else 0.1)
    for got_n, res, plan, state in results:
        on = crop[plan.band_lo : plan.band_hi] > 0
        np.testing.assert_array_equal(state[on], wstate[ids[plan.band_lo : plan.band_hi][on]])
    # and against the CPU oracle of the EquSolver
    want = np_oracle.EquOracle("max")
    want.reset(src, mask, big_tgt, (0, 0), (31, 52))
    want.X = c_oracle.equ_sweeps(want.A, want.X, want.B, 75)
    np.testing.assert_array_equal(wstate, want.X)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import fpie_b200
    from fpie_b200 import band

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        mask, tgt, grad = random_grid(900, 517, seed=2)
        solver = band.BandGridSolver(band.CudaBandCore(fpie_b200.GridSolver(8, 8, device=rank)), dist, halo=16)
        solver.reset(mask.size, mask, tgt, grad)
        solver.sync()
        solver.step(20)
        img, err = solver.step(45)
        p = solver.plan
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=solver.band_state(), err=err, lo=p.band_lo, hi=p.band_hi)
    finally:
        dist.destroy_process_group()


def test_nccl_bands(tmp_path):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mask, tgt, grad = random_grid(900, 517, seed=2)
    want = c_oracle.grid_sweeps(mask, tgt, grad, 65)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(z["state"], want[int(z["lo"]) : int(z["hi"])])
        np.testing.assert_allclose(z["err"], c_oracle.grid_residual(mask, want, grad)[1], rtol=1e-5)

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


TRANSPORTS = ["p2p", "nccl"]  # "nccl" on ThreadDist = the round-1 schedule with device copies standing in for NCCL


def _thread_core(transport):
    """One band's GridSolver inside a thread.  The p2p link waits on flag words AT STREAM LEVEL, so two bands
    of one process must not share a stream (the wait of one would block the other's sends): each gets its own."""
    import torch

    import fpie_b200

    if transport != "p2p":
        return fpie_b200.GridSolver(8, 8, device=0), None
    stream = torch.cuda.Stream(device=0)
    with torch.cuda.stream(stream):
        return fpie_b200.GridSolver(8, 8, device=0), stream


def _run_threads(world, halo, make_and_reset, steps, transport="p2p"):
    import torch

    from fpie_b200 import band

    dist = ThreadDist(world)
    out = [None] * world
    errors = []

    def work(rank):
        try:
            dist.bind(rank)
            torch.cuda.set_device(0)
            core, stream = _thread_core(transport)
            solver = band.make_band_solver(core, dist, halo=halo, transport=transport, same_process=True)
            make_and_reset(solver)
            solver.sync()
            for it in steps:
                img, err = solver.step(it)
            out[rank] = (solver.plan, solver.band_state(), img, err, getattr(solver, "exchanges_done", None))
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


@pytest.mark.parametrize("transport", TRANSPORTS)
@pytest.mark.parametrize("world,halo,steps", [(2, 8, (37,)), (3, 16, (5, 40)), (4, 24, (64,)), (2, 24, (100, 3, 48))])
def test_thread_bands_core_level(world, halo, steps, transport):
    shape = (530, 301)
    mask, tgt, grad = random_grid(*shape, seed=8)
    out = _run_threads(world, halo, lambda s: s.reset(mask.size, mask, tgt, grad), steps, transport)
    want = c_oracle.grid_sweeps(mask, tgt, grad, sum(steps))
    got = np.zeros_like(want)
    if transport == "p2p":  # one exchange per started interval of `halo` sweeps, in every step
        assert all(o[4] == sum(-(-it // halo) for it in steps) for o in out)
    for plan, state, img, err, _ in out:
        got[plan.band_lo : plan.band_hi] = state
        np.testing.assert_array_equal(img, c_oracle.clip_u8(want[plan.band_lo : plan.band_hi]))
        np.testing.assert_allclose(err, c_oracle.grid_residual(mask, want, grad)[1], rtol=1e-5)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("transport", TRANSPORTS)
@pytest.mark.parametrize("kind", ["square", "circle"])
def test_thread_bands_from_image_slabs(kind, transport):
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

    out = _run_threads(world, halo, reset, (iters,), transport)
    want = c_oracle.grid_sweeps(mc, tcrop, g, iters)
    for plan, state, img, err, _ in out:
        np.testing.assert_array_equal(state, want[plan.band_lo : plan.band_hi])
        np.testing.assert_allclose(err, c_oracle.grid_residual(mc, want, g)[1], rtol=1e-5)


@pytest.mark.parametrize("transport", TRANSPORTS)
def test_thread_band_processor_image_level(transport):
    """BandGridProcessor: full images in on every rank, blended target out on rank 0."""
    import torch

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
            core, stream = _thread_core(transport)
            proc = band.BandGridProcessor("max", band.CudaBandCore(core), dist, halo=16, transport=transport,
                                          same_process=True)
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


@pytest.mark.parametrize("transport", TRANSPORTS)
@pytest.mark.parametrize("kind,world,halo", [("star", 3, 16), ("holes", 2, 24), ("ring", 4, 8)])
def test_thread_band_equ_processor_matches_single_gpu_equ_solver(kind, world, halo, transport):
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
            core, stream = _thread_core(transport)
            proc = band.BandEquProcessor("max", band.CudaBandCore(core), dist, halo=halo, transport=transport,
                                         same_process=True)
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
    # (the band residual is evaluated in the EquSolver's own expression, |B + sum X[A] - 4X|, so it holds the
    # stated tolerance also where the residual is rounding noise: the tiny domains random holes leave)
    np.testing.assert_allclose(err, werr, rtol=1e-4)
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


def _band_worker(rank, world, port, out_dir, backend, transport, devices, shape, halo, steps):
    """One band in its own process.  ``backend`` is the control plane (nccl: one process per GPU; gloo: the
    processes may share a GPU -- the halo rows then cross process boundaries through CUDA IPC on one device,
    the same code path as between GPUs)."""
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import fpie_b200
    from fpie_b200 import band

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = devices[rank]
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        mask, tgt, grad = random_grid(*shape, seed=2)
        solver = band.make_band_solver(fpie_b200.GridSolver(8, 8, device=dev), dist, halo=halo, transport=transport)
        for round_ in range(2):  # the second reset keeps the link (same geometry) and its counters
            solver.reset(mask.size, mask, tgt, grad)
            solver.sync()
            for it in steps:
                img, err = solver.step(it)
        p = solver.plan
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=solver.band_state(), err=err, lo=p.band_lo, hi=p.band_hi,
                 img=img, exchanges=getattr(solver, "exchanges_done", -1))
    finally:
        dist.destroy_process_group()


def _check_band_files(tmp_path, world, shape, steps):
    mask, tgt, grad = random_grid(*shape, seed=2)
    want = c_oracle.grid_sweeps(mask, tgt, grad, sum(steps))
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(z["state"], want[int(z["lo"]) : int(z["hi"])])
        np.testing.assert_array_equal(z["img"], c_oracle.clip_u8(want[int(z["lo"]) : int(z["hi"])]))
        np.testing.assert_allclose(z["err"], c_oracle.grid_residual(mask, want, grad)[1], rtol=1e-5)
    return [int(np.load(tmp_path / f"rank{r}.npz")["exchanges"]) for r in range(world)]


@pytest.mark.parametrize("world,halo,steps", [(2, 16, (20, 45)), (3, 24, (130,))])
def test_ipc_bands_processes_sharing_one_gpu(tmp_path, world, halo, steps):
    """The SHIPPED multi-GPU transport on a one-GPU box: one process per band, all on cuda:0, gloo as the
    control plane, halo rows through cudaIpcOpenMemHandle-mapped receive boxes, flag words and stream-level
    waits -- exactly what runs between GPUs, minus NVLink."""
    import torch.multiprocessing as mp

    shape = (900, 517)
    mp.spawn(_band_worker, args=(world, _free_port(), str(tmp_path), "gloo", "p2p", [0] * world, shape, halo, steps),
             nprocs=world, join=True)
    exchanges = _check_band_files(tmp_path, world, shape, steps)
    assert all(e == 2 * sum(-(-it // halo) for it in steps) for e in exchanges)


@pytest.mark.parametrize("transport", TRANSPORTS)
def test_multi_gpu_bands(tmp_path, transport):
    """One process per GPU over NCCL (skipped on a one-GPU box; bench.py --gpus N checks the same path in
    every multi-GPU run and reports it as `parity`)."""
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    shape, steps = (900, 517), (20, 45)
    mp.spawn(_band_worker, args=(world, _free_port(), str(tmp_path), "nccl", transport, list(range(world)), shape, 16,
                                 steps), nprocs=world, join=True)
    _check_band_files(tmp_path, world, shape, steps)


def test_two_bands_of_a_16384_class_grid_equal_the_single_solve():
    """A grid of the size class the multi-GPU path exists for (16384 x 16384, 268 M pixels, square mask): two
    row bands linked by the p2p halo transport against ONE solver on the same images -- state, uint8 image
    and err.  (On a one-GPU box both bands and the single solve share the device; bench.py --gpus N runs the
    same check between GPUs in every multi-GPU run.)"""
    import sys

    import torch

    import fpie_b200
    from fpie_b200 import band

    sys.path.insert(0, ROOT)
    import bench

    n = m = 16384
    world, halo, sweeps = 2, 24, 96
    work = dict(mask="square", grad="max")
    dist = ThreadDist(world)
    results, errors = [None] * world, []

    def work_fn(rank):
        try:
            dist.bind(rank)
            torch.cuda.set_device(0)
            core, stream = _thread_core("p2p")
            solver = band.make_band_solver(core, dist, halo=halo, transport="p2p", same_process=True)
            plan = band.make_plan(n, world, rank, halo)
            src, mask, tgt, _ = bench.slab_images(work, plan.slab_lo, plan.slab_hi, n, m)
            solver.reset_slab(n, src, mask, tgt, "max")
            del src, mask, tgt
            solver.sync()
            img, err = solver.step(sweeps)
            results[rank] = (plan, np.array(img, copy=True), err, solver.band_state())
            core.close()
        except Exception as exc:
            errors.append(exc)
            try:
                dist.bar.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=work_fn, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    src, mask, tgt, unknowns = bench.slab_images(work, 0, n, n, m)
    one = fpie_b200.GridSolver(8, 8, device=0)
    one.reset_slab(src, mask, tgt, "max")
    del src, mask, tgt
    assert one.info()["unknowns"] == unknowns == (n - 2) * (m - 2)
    img1, err1 = one.step(sweeps)
    for plan, img, err, state in results:
        np.testing.assert_array_equal(img, img1[plan.band_lo : plan.band_hi])
        np.testing.assert_allclose(err, err1, rtol=1e-4)
    state1 = one.state()
    for plan, img, err, state in results:
        assert np.array_equal(state.view(np.uint32), state1[plan.band_lo : plan.band_hi].view(np.uint32))
    one.close()

"""The id-range sharded EquSolver on the GPU (fpie_b200/shard.py): the C-ABI data plane (row window,
gather_rows / scatter_rows on device buffers) and the sharded solve against single-domain Jacobi -- one process
per shard sharing cuda:0 with gloo as the process group on a one-GPU box, one process per GPU over NCCL where
there are several."""

import os
import socket
import sys

import numpy as np
import pytest
from conftest import PKG_ROOT, ROOT

from oracle import c_oracle, np_oracle

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_window_gather_scatter_through_the_c_abi():
    import torch

    import fpie_b200
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem("holes", 90, 120, seed=8)
    n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "avg")
    s = fpie_b200.EquSolver(256, mode="gather")
    s.reset(n, A, X, B)
    s.step(7)
    state = s.state()
    rng = np.random.default_rng(0)
    idx = rng.integers(1, n, 500).astype(np.int32)
    d_idx = torch.from_numpy(idx).cuda()
    out = torch.empty((idx.size, 3), dtype=torch.float32, device="cuda")
    s.gather_rows(d_idx.data_ptr(), idx.size, out.data_ptr())
    np.testing.assert_array_equal(out.cpu().numpy(), state[idx])
    # scatter: unique rows take the values handed in, everything else is untouched
    uniq = np.unique(idx)
    vals = rng.random((uniq.size, 3), dtype=np.float32) * 255
    d_uniq, d_vals = torch.from_numpy(uniq).cuda(), torch.from_numpy(vals).cuda()  # (kept alive across the call)
    s.scatter_rows(d_uniq.data_ptr(), uniq.size, d_vals.data_ptr())
    want = state.copy()
    want[uniq] = vals
    np.testing.assert_array_equal(s.state(), want)
    # the sweeps continue from the scattered state, like the oracle's
    s.step(3)
    np.testing.assert_array_equal(s.state(), c_oracle.equ_sweeps(A, want, B, 3))
    # residual window: rows [lo, hi) only; the uint8 rows of the same window
    cur = s.state()
    lo, hi = n // 3, n // 3 + 1000
    s.set_window(lo, hi)
    s.finish_async()
    img, err = s.fetch_rows(lo, hi)
    terms = np.abs((B + cur[A[:, 0]] + cur[A[:, 1]] + cur[A[:, 2]] + cur[A[:, 3]] - 4.0 * cur).astype(np.float64))
    np.testing.assert_allclose(err, terms[lo:hi].sum(0), rtol=1e-5)
    np.testing.assert_array_equal(img, c_oracle.clip_u8(cur)[lo:hi])
    s.set_window(0, n)
    _, err_all = s.step(0)
    np.testing.assert_allclose(err_all, terms.sum(0), rtol=1e-5)
    # errors: indices outside the system, row 0 as a scatter target, a window outside [0, N]
    bad = torch.tensor([1, n], dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="outside"):
        s.gather_rows(bad.data_ptr(), 2, out.data_ptr())
    zero = torch.tensor([0], dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="outside"):
        s.scatter_rows(zero.data_ptr(), 1, out.data_ptr())
    with pytest.raises(RuntimeError, match="set_window"):
        s.set_window(5, n + 1)
    promoted = fpie_b200.EquProcessor("max")
    promoted.reset(src, mask, tgt)
    with pytest.raises(RuntimeError, match="index-mapped"):
        promoted.core.scatter_rows(d_idx.data_ptr(), 1, out.data_ptr())


def _shard_worker(rank, world, port, out_dir, backend, devices, shape, depth, steps, labelling, env):
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(env)
    import torch
    import torch.distributed as dist
    from test_shard_cpu import make_system

    from fpie_b200 import shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = devices[rank]
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        n, A, X, B = make_system("ring", shape, 7, labelling)
        solver = shard.make_sharded_equ_solver(dist, depth=depth, device=dev)
        for _ in range(2):  # reset twice: the second one rebuilds plan and exchange lists
            solver.reset(n, A, X, B)
            solver.sync()
            for it in steps:
                img, err = solver.step(it)
        info = solver.core.solver.info()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=solver.state(), img=img, err=err,
                 exchanges=solver.exchanges, ghosts=solver.plan.ghosts, table=str(info["table"]), path=info["path"])
    finally:
        dist.destroy_process_group()


def _check_shard_files(tmp_path, world, shape, steps, labelling):
    from test_shard_cpu import make_system

    n, A, X, B = make_system("ring", shape, 7, labelling)
    want = c_oracle.equ_sweeps(A, X, B, sum(steps))
    werr = np_oracle.equ_residual_f64(A, want, B)
    out = []
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(z["state"], want)
        np.testing.assert_array_equal(z["img"][1:], c_oracle.clip_u8(want)[1:])
        np.testing.assert_allclose(z["err"], werr, rtol=1e-4)
        out.append((int(z["exchanges"]), int(z["ghosts"]), str(z["table"]), str(z["path"])))
    return out


@pytest.mark.parametrize("world,depth,steps,labelling,env", [
    (2, 8, (20, 45), "rowmajor", {}),
    (3, 16, (130,), "rowmajor", {"FPIE_B200_DELTA16_MIN": "0"}),
    (2, 5, (23,), "shuffled", {}),
    (3, 4, (21,), "redblack", {})])
def test_shards_in_processes_sharing_one_gpu(tmp_path, world, depth, steps, labelling, env):
    """One process per shard, all on cuda:0, gloo as the process group (messages staged through the host): the
    library's sweeps, row window, gather / scatter kernels and the exchange schedule, for row-major ids (compact
    tables, also the 4-byte distance table), a random permutation (int4 table, ghosts from every rank) and the
    OpenMP backend's red-black order."""
    import torch.multiprocessing as mp

    shape = (300, 260)
    mp.spawn(_shard_worker, args=(world, _free_port(), str(tmp_path), "gloo", [0] * world, shape, depth, steps,
                                  labelling, env), nprocs=world, join=True)
    seen = _check_shard_files(tmp_path, world, shape, steps, labelling)
    assert all(e == sum(steps) // depth and g > 0 for e, g, _, _ in seen)
    if labelling == "rowmajor":
        assert all(p == "gather-compact" for _, _, _, p in seen)
        if env:
            assert all(t.startswith("delta16") for _, _, t, _ in seen)
    if labelling == "shuffled":
        assert all(p == "gather-int4" for _, _, _, p in seen)


def test_multi_gpu_shards_over_nccl(tmp_path):
    """One process per GPU, ghost rows over NCCL send / recv (skipped on a one-GPU box)."""
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    shape, steps = (700, 900), (20, 45)
    for labelling in ("rowmajor", "shuffled"):
        d = tmp_path / labelling
        d.mkdir()
        mp.spawn(_shard_worker, args=(world, _free_port(), str(d), "nccl", list(range(world)), shape, 16, steps,
                                      labelling, {}), nprocs=world, join=True)
        _check_shard_files(d, world, shape, steps, labelling)

"""Host logic of the id-range sharded EquSolver (fpie_b200/shard.py) on the CPU: ghost layers by breadth-first
search over A, the pairwise exchange lists, exchange every ``depth`` sweeps over ``gloo`` (real processes),
global err and image.  The per-rank compute is the numpy oracle (OracleEquShardCore), so what is tested is the
orchestration the GPU path reuses with ``fpie_b200.EquSolver(mode="gather")`` + NCCL."""

import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp
from conftest import PKG_ROOT, ROOT

from oracle import np_oracle


def make_system(kind, shape, seed, labelling):
    """An Equ system as the Processor builds it (row-major ids) or relabelled: ``shuffled`` = a random
    permutation of the ids (the worst case: every rank needs ghosts from every other), ``redblack`` = the
    OpenMP backend's odd-before-even order (openmp/equ.cc:22-56)."""
    from fpie_b200 import synth

    src, mask, tgt = synth.make_problem(kind, *shape, seed=seed)
    n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "max")
    if labelling == "rowmajor":
        return n, A, X, B
    rng = np.random.default_rng(seed + 100)
    if labelling == "shuffled":
        perm = np.concatenate([[0], 1 + rng.permutation(n - 1)])  # new id of old id
    else:
        m_full, (x0, x1, y0, y1) = np_oracle.canonical_mask(mask)
        crop = m_full[x0:x1, y0:y1]
        ids_rb, _ = np_oracle.partition_redblack(crop)
        ids_rm = np_oracle.partition_rowmajor(crop)
        perm = np.zeros(n, np.int64)
        on = crop > 0
        perm[ids_rm[on]] = ids_rb[on]
    inv = np.argsort(perm)
    return n, perm[A[inv]].astype(np.int32), X[inv], B[inv]


def test_id_ranges_follow_the_reference_rule():
    from fpie_b200 import shard

    assert shard.id_ranges(11, 3) == [1, 5, 8, 11]  # mpi/equ.cc:55-59 over the 10 unknowns
    assert shard.id_ranges(1, 4) == [1, 1, 1, 1, 1]
    assert list(shard.owner_of(np.array([1, 4, 5, 10]), 11, 3)) == [0, 0, 1, 2]


@pytest.mark.parametrize("labelling", ["rowmajor", "shuffled"])
def test_ghost_layers_are_breadth_first_and_local_tables_close(labelling):
    from fpie_b200 import shard

    n, A, X, B = make_system("holes", (40, 37), 3, labelling)
    for world, depth in ((3, 1), (3, 4), (5, 9)):
        seen_owned = np.zeros(n, int)
        for rank in range(world):
            plan, rows, A_loc = shard.build_shard(A, rank, world, depth)
            seen_owned[plan.lo : plan.hi] += 1
            assert rows[0] == 0 and np.all(np.diff(rows) > 0)
            assert np.array_equal(rows[plan.own_lo : plan.own_hi], np.arange(plan.lo, plan.hi))
            # distance of every local row from the owned set, by an independent relaxation
            dist = np.full(n, 10**6)
            dist[plan.lo : plan.hi] = 0
            for _ in range(depth):
                for i in np.flatnonzero(dist < 10**6):
                    for j in A[i]:
                        if j > 0:
                            dist[j] = min(dist[j], dist[i] + 1)
            want_local = np.flatnonzero((dist <= depth) & (np.arange(n) > 0))
            assert np.array_equal(plan.local_ids, want_local)
            # local table: neighbours inside the local set keep their identity, others read row 0
            back = rows[A_loc]
            inside = np.isin(A[rows], rows)
            assert np.array_equal(back[inside], A[rows][inside]) and not A_loc[~inside].any()
            # rows at distance < depth have all their neighbours locally
            full = dist[rows] < depth
            assert inside[full].all()
        assert np.all(seen_owned[1:] == 1) and seen_owned[0] == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, depth, steps, labelling, out_dir):
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from band_helpers import OracleEquShardCore
    from test_shard_cpu import make_system

    from fpie_b200 import shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, A, X, B = make_system("holes", (48, 41), 7, labelling)
        solver = shard.ShardedEquSolver(OracleEquShardCore(), dist, depth=depth)
        with pytest.raises(RuntimeError):
            solver.step(1)
        solver.reset(n, A, X, B)
        solver.sync()
        errs = []
        for it in steps:
            img, err = solver.step(it)
            errs.append(err)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=solver.state(), img=img, err=np.array(errs),
                 exchanges=solver.exchanges, ghosts=solver.plan.ghosts, sent=solver.bytes_sent)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,depth,steps,labelling", [
    (2, 4, (37,), "rowmajor"), (3, 5, (5, 20, 12), "rowmajor"), (2, 1, (9,), "rowmajor"),
    (3, 6, (13, 11), "shuffled"), (2, 3, (10,), "redblack"), (4, 16, (40,), "rowmajor")])
def test_gloo_shards_reproduce_global_jacobi(tmp_path, world, depth, steps, labelling):
    """Id-range shards with ``depth`` ghost layers, exchanged every ``depth`` sweeps: fp32 state bit-identical
    to single-domain Jacobi for row-major, randomly permuted and red-black labellings; the whole uint8 image and
    the global err on every rank."""
    mp.spawn(_worker, args=(world, _free_port(), depth, steps, labelling, str(tmp_path)), nprocs=world, join=True)
    n, A, X, B = make_system("holes", (48, 41), 7, labelling)
    want = np_oracle.equ_sweeps(A, X, B, sum(steps))
    werr = np_oracle.equ_residual_f64(A, want, B)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(z["state"], want)
        np.testing.assert_array_equal(z["img"][1:], np_oracle.clip_u8(want)[1:])
        np.testing.assert_allclose(z["err"][-1], werr, rtol=1e-5)
        assert int(z["exchanges"]) == sum(steps) // depth
    if labelling == "rowmajor" and world == 3:
        z = np.load(tmp_path / "rank1.npz")
        # a middle rank's ghosts: `depth` layers on each side, each about one image row of unknowns
        assert 0 < int(z["ghosts"]) < 2 * depth * 2 * 41


def _tiny_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG_ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from band_helpers import OracleEquShardCore

    from fpie_b200 import shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 3 unknowns in a row on 5 ranks: two ranks own nothing (and hold only the constant row)
        A = np.array([[0, 0, 0, 0], [0, 0, 0, 2], [0, 0, 1, 3], [0, 0, 2, 0]], np.int32)
        X = np.array([[0, 0, 0], [8, 16, 24], [1, 2, 3], [40, 80, 120]], np.float32)
        B = np.array([[0, 0, 0], [4, 4, 4], [-2, 0, 2], [9, 9, 9]], np.float32)
        solver = shard.ShardedEquSolver(OracleEquShardCore(), dist, depth=2)
        solver.reset(4, A, X, B)
        img, err = solver.step(7)
        np.savez(os.path.join(out_dir, f"tiny{rank}.npz"), state=solver.state(), img=img, err=err, lo=solver.plan.lo,
                 hi=solver.plan.hi)
    finally:
        dist.destroy_process_group()


def test_more_ranks_than_unknowns_and_a_single_rank(tmp_path):
    """Empty id ranges (5 ranks, 3 unknowns) still take part in the collectives and return the whole result; one
    rank alone has no peers and never exchanges."""
    mp.spawn(_tiny_worker, args=(5, _free_port(), str(tmp_path)), nprocs=5, join=True)
    A = np.array([[0, 0, 0, 0], [0, 0, 0, 2], [0, 0, 1, 3], [0, 0, 2, 0]], np.int32)
    X = np.array([[0, 0, 0], [8, 16, 24], [1, 2, 3], [40, 80, 120]], np.float32)
    B = np.array([[0, 0, 0], [4, 4, 4], [-2, 0, 2], [9, 9, 9]], np.float32)
    want = np_oracle.equ_sweeps(A, X, B, 7)
    owned = 0
    for r in range(5):
        z = np.load(tmp_path / f"tiny{r}.npz")
        np.testing.assert_array_equal(z["state"], want)
        np.testing.assert_array_equal(z["img"][1:], np_oracle.clip_u8(want)[1:])
        np.testing.assert_allclose(z["err"], np_oracle.equ_residual_f64(A, want, B), rtol=1e-5)
        owned += int(z["hi"]) - int(z["lo"])
    assert owned == 3
    mp.spawn(_worker, args=(1, _free_port(), 4, (9,), "rowmajor", str(tmp_path)), nprocs=1, join=True)
    n, A, X, B = make_system("holes", (48, 41), 7, "rowmajor")
    z = np.load(tmp_path / "rank0.npz")
    np.testing.assert_array_equal(z["state"], np_oracle.equ_sweeps(A, X, B, 9))
    assert int(z["ghosts"]) == 0 and int(z["sent"]) == 0


def test_ghost_layer_scheme_is_exact_on_arbitrary_gather_graphs():
    """`build_shard` on RANDOM index tables -- any row may gather from any four rows, nothing grid-like about
    them -- with the exchange simulated in-process: owned rows after S sweeps equal global Jacobi bit for bit for
    every (ranks, depth), i.e. the breadth-first ghost layers follow the direction in which rows READ."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from fpie_b200 import shard

    @settings(max_examples=40, deadline=None)
    @given(st.integers(2, 70), st.integers(1, 6), st.integers(1, 5), st.integers(0, 17), st.integers(0, 2**31 - 1),
           st.sampled_from([0.0, 0.3, 0.8]))
    def check(n, world, depth, sweeps, seed, absent):
        rng = np.random.default_rng(seed)
        A = rng.integers(1, n, (n, 4)).astype(np.int32) if n > 1 else np.zeros((n, 4), np.int32)
        A[rng.random((n, 4)) < absent] = 0
        A[0] = 0
        X = rng.integers(0, 256, (n, 3)).astype(np.float32)
        B = (rng.integers(-2040, 2041, (n, 3)) / 2).astype(np.float32)
        X[0] = B[0] = 0
        want = np_oracle.equ_sweeps(A, X, B, sweeps)
        shards = []
        for r in range(world):
            plan, rows, A_loc = shard.build_shard(A, r, world, depth)
            shards.append([plan, rows, A_loc, X[rows].copy(), B[rows]])
        done = 0
        while done < sweeps:
            k = min(depth, sweeps - done)
            for sh in shards:
                sh[3] = np_oracle.equ_sweeps(sh[2], sh[3], sh[4], k)
            done += k
            if k == depth:  # refresh every ghost from its owner
                full = np.zeros_like(X)
                for plan, rows, _, xl, _ in shards:
                    full[plan.lo : plan.hi] = xl[plan.own_lo : plan.own_hi]
                for sh in shards:
                    plan, rows = sh[0], sh[1]
                    ghost = np.ones(rows.size, bool)
                    ghost[0] = False
                    ghost[plan.own_lo : plan.own_hi] = False
                    sh[3][ghost] = full[rows[ghost]]
        for plan, rows, _, xl, _ in shards:
            np.testing.assert_array_equal(xl[plan.own_lo : plan.own_hi], want[plan.lo : plan.hi])

    check()

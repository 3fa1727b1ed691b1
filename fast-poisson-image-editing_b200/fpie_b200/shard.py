"""Id-range sharded EquSolver: a general ``(A, X, B)`` system across devices, one process per GPU.

The reference's multi-worker EquSolver (fpie/core/mpi/equ.cc) gives worker ``p`` the id range
``offset[p] .. offset[p+1]`` with ``offset[i+1] = offset[i] + N/P + (i < N%P)`` (equ.cc:55-59), lets every
worker update its range in place for ``min_interval`` sweeps on stale copies of everybody else's unknowns, and
then moves ALL of ``X`` through rank 0 (equ.cc:136-146) -- neither Jacobi nor scalable.  Here the same id
ranges are kept, but a rank also holds ``depth`` LAYERS OF GHOST UNKNOWNS around its range (breadth-first
over ``A``: layer ``d`` = unknowns ``d`` gather steps away from an owned one), runs at most ``depth`` sweeps
on that local system and then refreshes the ghosts from their owners: after ``s <= depth`` sweeps only the
ghosts of layers ``> depth - s`` are stale, the owned range is exact, so the sharded result equals
single-device Jacobi bit for bit -- the general-graph form of the deep halo of ``band.py``,
valid for ANY labelling, row-major or not.  What moves per exchange are the ghost rows only, pairwise between
the ranks that share them (``batch_isend_irecv``: NCCL over NVLink between GPUs), packed / unpacked by the
solver's own kernels (``fpie_b200_equ_gather_rows`` / ``scatter_rows``).

``ShardedEquSolver`` is host logic over an injected per-rank core and process group -- the CPU test-suite
runs it with ``gloo`` over a numpy stand-in core, the GPU path with ``fpie_b200.EquSolver(mode="gather")``
behind ``CudaEquShardCore``.  Systems the Processor builds (row-major ids of an image mask) are better served
by ``BandEquProcessor`` (band.py: row bands on the temporally blocked kernel); this class is for the core
API's general case -- ``reset(N, A, X, B)`` with whatever ids the caller's ``partition`` produced.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .band import band_offsets


@dataclass
class ShardPlan:
    """What one rank holds of the global system."""

    rank: int
    world: int
    N: int  # global rows (row 0 = the constant)
    depth: int
    lo: int  # owned global ids [lo, hi)
    hi: int
    local_ids: np.ndarray  # global ids of local rows 1.. (sorted; owned ids are one contiguous run of it)
    own_lo: int  # local rows [own_lo, own_hi) are the owned ids
    own_hi: int
    layers: list = field(default_factory=list)  # ghost ids per layer (global, sorted)

    @property
    def ghosts(self) -> int:
        return int(self.local_ids.size - (self.hi - self.lo))


def id_ranges(N: int, world: int) -> list[int]:
    """Owned id ranges of the unknowns 1..N-1: rank ``p`` owns ``[1 + off[p], 1 + off[p+1])``
    (fpie/core/mpi/equ.cc:55-59, applied to the unknowns; row 0, the constant, is everybody's)."""
    return [1 + o for o in band_offsets(max(int(N) - 1, 0), world)]


def build_shard(A: np.ndarray, rank: int, world: int, depth: int):
    """``(plan, rows, A_local)``: ``rows`` = global rows of the local system (row 0 first), ``A_local`` the
    index table in local numbering.  Neighbours outside the local set read local row 0: they are needed only by
    the outermost ghost layer, which is stale after one sweep anyway."""
    A = np.ascontiguousarray(A, dtype=np.int32)
    N = A.shape[0]
    if depth < 1:
        raise ValueError("depth must be >= 1")
    if N < 1 or A.shape != (N, 4):
        raise ValueError("A must be int32 [N, 4]")
    if np.any(A[0] != 0):
        raise ValueError("row 0 must be the constant row (A[0] = 0): process.py:248-250")
    if A.min() < 0 or A.max() >= N:
        raise ValueError("A holds an index outside [0, N)")
    off = id_ranges(N, world)
    lo, hi = off[rank], off[rank + 1]
    mark = np.zeros(N, np.bool_)
    mark[lo:hi] = True
    mark[0] = True  # (never a ghost)
    layers = []
    frontier = np.arange(lo, hi, dtype=np.int64)
    for _ in range(depth):
        if frontier.size == 0:
            break
        nb = A[frontier].ravel()
        nb = nb[~mark[nb]]
        new = np.unique(nb).astype(np.int64)
        mark[new] = True
        if new.size:
            layers.append(new)
        frontier = new
    mark[0] = False
    local_ids = np.flatnonzero(mark).astype(np.int64)
    local_of = np.zeros(N, np.int32)
    local_of[local_ids] = np.arange(1, local_ids.size + 1, dtype=np.int32)
    rows = np.concatenate([np.zeros(1, np.int64), local_ids])
    A_local = local_of[A[rows]]
    own_lo = int(local_of[lo]) if hi > lo else 1
    plan = ShardPlan(rank, world, N, depth, lo, hi, local_ids, own_lo, own_lo + (hi - lo), layers)
    return plan, rows, A_local


def owner_of(ids: np.ndarray, N: int, world: int) -> np.ndarray:
    off = np.asarray(id_ranges(N, world))
    return np.searchsorted(off, ids, side="right") - 1


class ShardedEquSolver:
    """``partition / reset / sync / step`` of the reference EquSolver (fpie/core/mpi/solver.h role), sharded
    by id range.  Every rank is handed the whole system at ``reset`` (the reference broadcasts it from rank 0 in
    ``sync``, equ.cc:63-87) and every rank returns the whole uint8 result and the global ``err``."""

    def __init__(self, core, dist, group=None, depth: int = 16):
        self.core, self.dist, self.group = core, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.depth = int(depth)
        self.plan: ShardPlan | None = None
        self.since_exchange = 0
        self.exchanges = 0
        self.bytes_sent = 0

    # -- reference interface ----------------------------------------------------------------------------
    def partition(self, mask) -> np.ndarray:
        return self.core.partition(mask)

    def reset(self, N, A, X, B) -> None:
        import torch

        N = int(N)
        A = np.ascontiguousarray(A, dtype=np.int32)
        X = np.ascontiguousarray(X, dtype=np.float32)
        B = np.ascontiguousarray(B, dtype=np.float32)
        if A.shape != (N, 4) or X.shape != (N, 3) or B.shape != (N, 3):
            raise ValueError("expected A[N,4], X[N,3], B[N,3]")
        plan, rows, A_local = build_shard(A, self.rank, self.world, self.depth)
        self.plan = plan
        self.core.reset(rows.size, A_local, X[rows], B[rows])
        self.core.set_window(plan.own_lo, plan.own_hi)
        # who owns my ghosts -> what I receive; the peers' requests -> what I send
        ghosts = np.concatenate([plan.local_ids[: plan.own_lo - 1], plan.local_ids[plan.own_hi - 1 :]])
        owners = owner_of(ghosts, N, self.world)
        want = {int(p): ghosts[owners == p] for p in np.unique(owners)}
        asked = [None] * self.world
        self.dist.all_gather_object(asked, want, group=self.group)

        def rows_of(ids):
            return (np.searchsorted(plan.local_ids, ids) + 1).astype(np.int32)

        self._peers = sorted(set(want) | {p for p, req in enumerate(asked) if req and self.rank in req})
        send_idx, recv_idx, self._send_slices, self._recv_slices = [], [], {}, {}
        s_at = r_at = 0
        for p in self._peers:
            need = np.asarray((asked[p] or {}).get(self.rank, np.zeros(0, np.int64)), dtype=np.int64)
            if need.size:
                if need.min() < plan.lo or need.max() >= plan.hi:
                    raise RuntimeError(f"rank {p} asked rank {self.rank} for ids it does not own")
                send_idx.append(rows_of(need))
                self._send_slices[p] = (s_at, s_at + need.size)
                s_at += need.size
            got = want.get(p)
            if got is not None and got.size:
                recv_idx.append(rows_of(got))
                self._recv_slices[p] = (r_at, r_at + got.size)
                r_at += got.size
        self._n_send, self._n_recv = s_at, r_at
        self._send_idx = self.core.make_index(np.concatenate(send_idx) if send_idx else np.zeros(0, np.int32))
        self._recv_idx = self.core.make_index(np.concatenate(recv_idx) if recv_idx else np.zeros(0, np.int32))
        self.since_exchange = 0
        self.exchanges = 0
        self.bytes_sent = 0
        self._torch = torch
        # one checked round trip of the data plane (validates every index list on the device), then unchecked
        self._exchange()
        self.core.rows_checked(True)
        self.exchanges = 0
        self.bytes_sent = 0

    def sync(self) -> None:
        self.dist.barrier(group=self.group)

    def step(self, iteration: int):
        """``iteration`` more Jacobi sweeps; ``(uint8 [N, 3], err float32 [3])`` of the WHOLE system on every rank."""
        self._need_reset()
        self.sweeps(iteration)
        plan = self.plan
        self.core.finish_async()
        own, err = self.core.fetch_rows(plan.own_lo, plan.own_hi)
        total = self._all_reduce_sum(np.asarray(err, np.float64))
        img = self._all_gather_rows(np.ascontiguousarray(own, dtype=np.uint8), np.uint8)
        return img, total.astype(np.float32)

    # -- pieces -----------------------------------------------------------------------------------------
    def sweeps(self, iteration: int) -> None:
        self._need_reset()
        left = int(iteration)
        while left > 0:
            k = min(left, self.depth - self.since_exchange)
            self.core.sweeps_async(k)
            self.since_exchange += k
            left -= k
            # eagerly: the residual (and anything else that reads a ghost) then always finds layer 1 exact
            if self.since_exchange == self.depth:
                self._exchange()

    def state(self) -> np.ndarray:
        """fp32 state ``[N, 3]`` of the whole system, stitched from the ranks' owned rows (parity tests)."""
        self._need_reset()
        plan = self.plan
        own = np.ascontiguousarray(self.core.state()[plan.own_lo : plan.own_hi], dtype=np.float32)
        out = self._all_gather_rows(own, np.float32)
        out[0] = self.core.state()[0]
        return out

    def _need_reset(self) -> None:
        if self.plan is None:
            raise RuntimeError("ShardedEquSolver: step called before reset")

    def _comm_device(self):
        """Where tensors handed to the process group must live: the GPU for NCCL, the host for gloo."""
        backend = str(self.dist.get_backend(self.group)).lower()
        return self.core.device if "nccl" in backend else self._torch.device("cpu")

    def _exchange(self) -> None:
        torch, dist = self._torch, self.dist
        dev = self._comm_device()
        send = self.core.gather(self._send_idx, self._n_send).to(dev)
        recv = torch.empty((self._n_recv, 3), dtype=torch.float32, device=dev)
        ops = []
        for p in self._peers:
            if p in self._send_slices:
                a, b = self._send_slices[p]
                ops.append(dist.P2POp(dist.isend, send[a:b], self._global_rank(p), self.group))
            if p in self._recv_slices:
                a, b = self._recv_slices[p]
                ops.append(dist.P2POp(dist.irecv, recv[a:b], self._global_rank(p), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self.core.scatter(self._recv_idx, self._n_recv, recv.to(self.core.device))
        self.since_exchange = 0
        self.exchanges += 1
        self.bytes_sent += self._n_send * 12

    def _global_rank(self, p: int) -> int:
        return p if self.group is None else self.dist.get_global_rank(self.group, p)

    def _all_reduce_sum(self, values: np.ndarray) -> np.ndarray:
        t = self._torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(self._comm_device())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def _all_gather_rows(self, own: np.ndarray, dtype) -> np.ndarray:
        """Stitch ``[N, 3]`` from every rank's owned rows (row 0 left zero)."""
        torch = self._torch
        off = id_ranges(self.plan.N, self.world)
        most = max(off[p + 1] - off[p] for p in range(self.world))
        dev = self._comm_device()
        pad = np.zeros((most, 3), dtype)
        pad[: own.shape[0]] = own
        mine = torch.from_numpy(pad).to(dev)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine, group=self.group)
        out = np.zeros((self.plan.N, 3), dtype)
        for p in range(self.world):
            n = off[p + 1] - off[p]
            if n:
                out[off[p] : off[p + 1]] = parts[p][:n].cpu().numpy()
        return out


class CudaEquShardCore:
    """The per-rank core on a GPU: ``fpie_b200.EquSolver(mode="gather")``; index lists and message buffers are
    torch tensors on the solver's device (torch: buffers and the process group only -- packing, unpacking and the
    sweeps are the library's kernels, enqueued on the stream that was torch's current one at construction)."""

    def __init__(self, solver):
        import torch

        self._torch = torch
        self.solver = solver
        self.device = torch.device("cuda", solver.device)

    def partition(self, mask):
        return self.solver.partition(mask)

    def reset(self, N, A, X, B):
        self.solver.reset(N, A, X, B)

    def set_window(self, lo, hi):
        self.solver.set_window(lo, hi)

    def rows_checked(self, on):
        self.solver.rows_checked(on)

    def sweeps_async(self, k):
        self.solver.sweeps_async(k)

    def finish_async(self):
        self.solver.finish_async()

    def fetch_rows(self, lo, hi):
        return self.solver.fetch_rows(lo, hi)

    def state(self):
        return self.solver.state()

    def make_index(self, rows: np.ndarray):
        return self._torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int32)).to(self.device)

    def gather(self, idx, n: int):
        out = self._torch.empty((n, 3), dtype=self._torch.float32, device=self.device)
        if n:
            self.solver.gather_rows(idx.data_ptr(), n, out.data_ptr())
        return out

    def scatter(self, idx, n: int, rows) -> None:
        if n:
            rows = rows.contiguous()
            self.solver.scatter_rows(idx.data_ptr(), n, rows.data_ptr())
            # `rows` must outlive the kernel: same stream as torch's allocator, so releasing it here is safe


def make_sharded_equ_solver(dist, group=None, depth: int = 16, device: int | None = None, block_size: int = 256):
    """The GPU wiring: one ``EquSolver(mode="gather")`` per rank behind ``ShardedEquSolver``."""
    from .solver import EquSolver

    return ShardedEquSolver(CudaEquShardCore(EquSolver(block_size, device=device, mode="gather")), dist, group, depth)

"""Shared pieces of the row-band tests: a numpy stand-in core (backed by the
oracle -- test infrastructure) and an in-process thread transport."""

import contextlib
import queue
import threading
from collections import namedtuple

import numpy as np
import torch

from oracle import np_oracle


class OracleBandCore:
    """Core protocol of ``fpie_b200.band.BandGridSolver`` on the CPU: the slab is
    swept by the numpy oracle with its outer frame held fixed, exactly what the
    CUDA GridSolver does to a slab.  Two state buffers and split passes
    (``set_edge_rows`` / ``pass_async`` / ``flip``) are modelled like the CUDA
    core's, so the overlapped exchange schedule of ``BandGridSolver`` runs -- and
    is checked for its ordering -- on the CPU too: a pass over one part computes
    the sweeps on the whole slab from the CURRENT buffer and commits only that
    part's rows to the NEXT buffer; an exchange that ran too early or a part that
    wrote rows it does not own would change the result."""

    torch_device = "cpu"
    equ_form = False
    EDGE, INTERIOR = 0, 1

    def __init__(self, block_k: int = 4):
        self.block_k = int(block_k)

    def set_formulation(self, equ):
        self.equ_form = bool(equ)

    def reset(self, N, mask, tgt, grad):
        m = np.array(mask, np.int32, copy=True)
        m[0, :] = m[-1, :] = 0
        m[:, 0] = m[:, -1] = 0  # GridSolver treats the grid frame as unmasked
        self.mask = m
        self.grad = np.array(grad, np.float32, copy=True)
        first = np.ascontiguousarray(np.asarray(tgt, np.float32).transpose(2, 0, 1))
        self.buf = [first, first.copy()]
        self.cur = 0
        self.window = (0, m.shape[0])
        self.edge_rows = 0

    @property
    def planes(self):
        return self.buf[self.cur]

    def reset_slab(self, src, mask, tgt, gradient):
        """uint8 slab images -> the same grid problem the CUDA slab reset builds: the whole slab is the
        grid, mask thresholded at 128, gradient from the slab's own pixels, outer frame fixed."""
        m = (np.asarray(mask).reshape(mask.shape[0], mask.shape[1], -1).mean(-1) >= 128).astype(np.int32)
        m[0, :] = m[-1, :] = 0
        m[:, 0] = m[:, -1] = 0
        n, w = m.shape
        grad = np_oracle._pixel_gradient(gradient, src, tgt, (0, 0), (0, 0), (n, w))
        grad[m == 0] = 0
        if self.equ_form:
            # the EquSolver's system on the grid (process.py:227-266): B = grad + targets of the neighbours
            # outside the GLOBAL mask; X = target on globally masked pixels (also on the halo frame rows,
            # which hold the neighbour band's unknowns), 0 elsewhere
            raw = np.asarray(mask).reshape(n, w, -1).mean(-1) >= 128
            t = tgt.astype(np.float32)
            pad_raw = np.pad(raw, 1)
            pad_t = np.pad(t, ((1, 1), (1, 1), (0, 0)))
            for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                nb_raw = pad_raw[1 + dr : 1 + dr + n, 1 + dc : 1 + dc + w]
                nb_t = pad_t[1 + dr : 1 + dr + n, 1 + dc : 1 + dc + w]
                grad += np.where((m > 0) & ~nb_raw, 1.0, 0.0)[:, :, None].astype(np.float32) * nb_t
            self.reset(n * w, m, t * raw[:, :, None], grad)
            return
        self.reset(n * w, m, tgt.astype(np.float32), grad)

    def _aos(self):
        return self.planes.transpose(1, 2, 0)

    def _swept(self, k):
        return np_oracle.grid_sweeps(self.mask, self._aos(), self.grad, k).transpose(2, 0, 1)

    def sweeps_async(self, k):
        left = int(k)
        while left > 0:  # passes of block_k, ping-pong like the CUDA core
            ns = min(left, self.block_k)
            self.buf[self.cur ^ 1][...] = self._swept(ns)
            self.cur ^= 1
            left -= ns

    # -- split passes ----------------------------------------------------------
    def set_edge_rows(self, rows):
        self.edge_rows = int(rows)

    def pass_async(self, nsweeps, part):
        assert 1 <= nsweeps <= self.block_k and self.edge_rows > 0
        n = self.mask.shape[0]
        e = min(self.edge_rows, n)
        edge = np.zeros(n, bool)
        edge[:e] = edge[n - e :] = True
        rows = edge if part == self.EDGE else ~edge
        self.buf[self.cur ^ 1][:, rows] = self._swept(nsweeps)[:, rows]

    def flip(self):
        self.cur ^= 1

    def next_buffer(self):
        return self.cur ^ 1

    @contextlib.contextmanager
    def exchange_scope(self):
        """Model the asynchronous exchange at its LATEST legal completion: rows received inside the scope
        land in staging tensors and only reach the halo rows at ``wait_exchange`` -- so a schedule that
        joins the exchange too late (after tiles that read the halo) computes from stale rows here,
        just as it could on the GPU."""
        self._staging = []
        try:
            yield
        finally:
            self._staged, self._staging = self._staging, None

    def wait_exchange(self):
        for which, lo, hi, tensors in getattr(self, "_staged", None) or []:
            for p in range(3):
                self.buf[which][p, lo:hi] = tensors[p].numpy()
        self._staged = None

    def set_row_window(self, lo, hi):
        self.window = (lo, hi)

    def finish_async(self):
        pass

    def fetch(self):
        lo, hi = self.window
        t = np.ascontiguousarray(self._aos())
        m = self.mask.copy()
        m[:lo] = 0
        m[hi:] = 0
        return np_oracle.clip_u8(t), np_oracle.grid_residual_f64(m, t, self.grad)

    def state(self):
        return np.ascontiguousarray(self._aos())

    def rows_view(self, lo, hi, which=None):
        which = self.cur if which is None else which
        buf = self.buf[which]
        band_lo, band_hi = self.window
        if getattr(self, "_staging", None) is not None and (hi <= band_lo or lo >= band_hi):
            # halo rows requested inside an exchange scope = receive targets: stage them
            tensors = [torch.empty((hi - lo, buf.shape[2]), dtype=torch.float32) for _ in range(3)]
            self._staging.append((which, lo, hi, tensors))
            return tensors
        return [torch.from_numpy(buf[p, lo:hi]) for p in range(3)]


class PlainOracleBandCore(OracleBandCore):
    """The same core without split passes: ``BandGridSolver`` exchanges between passes."""

    pass_async = None  # ``hasattr`` stays true for None, so hide it properly:

    def __getattribute__(self, name):
        if name == "pass_async":
            raise AttributeError(name)
        return object.__getattribute__(self, name)


class ThreadDist:
    """Minimal ``torch.distributed`` look-alike for ranks living in threads of one
    process (lets a single GPU exercise the CUDA halo path)."""

    P2POp = namedtuple("P2POp", "op tensor peer group")
    isend, irecv = "isend", "irecv"

    def __init__(self, world):
        self.world = world
        self.local = threading.local()
        self.q = {(a, b): queue.Queue() for a in range(world) for b in range(world)}
        self.bar = threading.Barrier(world)
        self.slots = [None] * world

    def bind(self, rank):
        self.local.rank = rank

    def get_rank(self, group=None):
        return self.local.rank

    def get_world_size(self, group=None):
        return self.world

    def barrier(self, group=None):
        self.bar.wait()

    def get_backend(self, group=None):
        return "threads"

    def all_gather_py(self, obj):
        """Every rank's python object, in rank order (what ``all_gather_object`` returns)."""
        me = self.local.rank
        self.slots[me] = obj
        self.bar.wait()
        out = list(self.slots)
        self.bar.wait()
        return out

    def batch_isend_irecv(self, ops):
        me = self.local.rank
        for op in ops:
            if op.op == "isend":
                sent = op.tensor.clone()
                if sent.is_cuda:  # the receiver copies on ITS stream: the clone must have completed
                    torch.cuda.current_stream(sent.device).synchronize()
                self.q[(me, op.peer)].put(sent)
        for op in ops:
            if op.op == "irecv":
                op.tensor.copy_(self.q[(op.peer, me)].get(timeout=120))
        return []

    def gather(self, tensor, gather_list=None, dst=0, group=None):
        me = self.local.rank
        self.slots[me] = tensor.clone()
        self.bar.wait()
        if me == dst:
            for i, out in enumerate(gather_list):
                out.copy_(self.slots[i])
        self.bar.wait()

    def all_reduce(self, tensor, group=None):
        me = self.local.rank
        self.slots[me] = tensor.clone()
        self.bar.wait()
        total = sum(s.to(tensor.device) for s in self.slots)
        self.bar.wait()
        tensor.copy_(total)


def random_grid(n, m, seed, density=0.7):
    rng = np.random.default_rng(seed)
    mask = np.zeros((n, m), np.int32)
    mask[1:-1, 1:-1] = rng.random((n - 2, m - 2)) < density
    tgt = rng.integers(0, 256, (n, m, 3)).astype(np.float32)
    grad = (rng.integers(-2040, 2041, (n, m, 3)) / 2).astype(np.float32)
    grad[mask == 0] = 0
    return mask, tgt, grad


class OracleEquShardCore:
    """Core protocol of ``fpie_b200.shard.ShardedEquSolver`` on the CPU: the local system (owned ids +
    ghost layers) is swept by the numpy oracle, exactly what ``EquSolver(mode="gather")`` does to it; index
    lists and message buffers are host tensors."""

    device = torch.device("cpu")

    def partition(self, mask):
        return np_oracle.partition_rowmajor(np.asarray(mask))

    def reset(self, N, A, X, B):
        self.A = np.array(A, np.int32, copy=True)
        self.X = np.array(X, np.float32, copy=True)
        self.B = np.array(B, np.float32, copy=True)
        assert self.A.shape == (N, 4) and self.X.shape == (N, 3)
        self.window = (0, N)
        self.checked = False

    def set_window(self, lo, hi):
        self.window = (lo, hi)

    def rows_checked(self, on):
        self.checked = bool(on)

    def sweeps_async(self, k):
        self.X = np_oracle.equ_sweeps(self.A, self.X, self.B, int(k))

    def finish_async(self):
        pass

    def fetch_rows(self, lo, hi):
        a, b = self.window
        keep = np.zeros(self.A.shape[0], bool)
        keep[a:b] = True
        s = self.B + self.X[self.A[:, 0]] + self.X[self.A[:, 1]] + self.X[self.A[:, 2]] + self.X[self.A[:, 3]] - 4.0 * self.X
        err = np.abs(s[keep].astype(np.float64)).sum(0)
        return np_oracle.clip_u8(self.X[lo:hi]), err.astype(np.float32)

    def state(self):
        return self.X.copy()

    def make_index(self, rows):
        rows = np.asarray(rows, np.int64)
        assert rows.size == 0 or (rows.min() >= 1 and rows.max() < self.A.shape[0])
        return rows

    def gather(self, idx, n):
        return torch.from_numpy(self.X[idx].copy()).reshape(n, 3)

    def scatter(self, idx, n, rows):
        self.X[idx] = rows.numpy().reshape(n, 3)

"""Row-band sharded GridSolver: one process per GPU, halo rows over NCCL.

The reference's multi-worker GridSolver (fpie/core/mpi/grid.cc) cuts the grid
into row bands with ``offset[i+1] = offset[i] + N/P + (i < N%P)`` (grid.cc:27-31)
and swaps ONE halo row every ``S`` sweeps (grid.cc:118-135), so its bands run on
stale neighbours in between and the result is not Jacobi.  Here every band keeps
``halo`` rows of each neighbour, runs at most ``halo`` sweeps on its slab
(band + halos) and then refreshes the halo rows from the neighbours' band
edges: after ``s <= halo`` sweeps only the outer ``s`` halo rows are stale, the
band itself is exact, so the sharded result equals single-device Jacobi bit for
bit (the test-suite carries a numpy model of exactly this scheme).

``BandGridSolver`` is transport- and device-agnostic host logic: the per-rank
compute object ("core") and the process group are injected, which is how the
CPU test-suite runs it with ``gloo`` over a numpy stand-in core, and how the
GPU path runs it with ``nccl`` over ``fpie_b200.GridSolver``.
"""

from __future__ import annotations

from dataclasses import dataclass

import contextlib

import numpy as np


def band_offsets(n_rows: int, parts: int) -> list[int]:
    """Row offsets of ``parts`` bands (fpie/core/mpi/grid.cc:27-31)."""
    off = [0]
    for i in range(parts):
        off.append(off[-1] + n_rows // parts + (1 if i < n_rows % parts else 0))
    return off


@dataclass(frozen=True)
class BandPlan:
    """Geometry of one rank's slab inside the global grid (rows only)."""

    rank: int
    world: int
    n_rows: int  # rows of the global grid
    halo: int
    band_lo: int  # first global row owned by this rank
    band_hi: int  # one past the last owned row
    slab_lo: int  # first global row held (band_lo - halo, clipped)
    slab_hi: int

    @property
    def up(self) -> int | None:
        """Rank owning the rows above (None at the top or for an empty band)."""
        return self._neighbour(-1)

    @property
    def down(self) -> int | None:
        return self._neighbour(+1)

    def _neighbour(self, step: int) -> int | None:
        if self.band_hi == self.band_lo:
            return None
        off = band_offsets(self.n_rows, self.world)
        r = self.rank + step
        while 0 <= r < self.world:
            if off[r + 1] > off[r]:
                return r
            r += step
        return None

    @property
    def local_band(self) -> tuple[int, int]:
        """Band rows in slab-local coordinates."""
        return self.band_lo - self.slab_lo, self.band_hi - self.slab_lo

    @property
    def slab_rows(self) -> int:
        return self.slab_hi - self.slab_lo


def make_plan(n_rows: int, world: int, rank: int, halo: int) -> BandPlan:
    if halo < 1:
        raise ValueError("halo depth must be >= 1")
    off = band_offsets(n_rows, world)
    lo, hi = off[rank], off[rank + 1]
    # every non-empty band must be at least `halo` rows tall, or a neighbour's halo would
    # reach past it into a third band
    sizes = [off[i + 1] - off[i] for i in range(world) if off[i + 1] > off[i]]
    if len(sizes) > 1 and min(sizes) < halo:
        raise ValueError(f"bands of {min(sizes)} rows are shorter than the halo depth {halo}")
    return BandPlan(rank, world, n_rows, halo, lo, hi, max(lo - halo, 0), min(hi + halo, n_rows))


class BandGridSolver:
    """GridSolver interface (``reset / sync / step``) over row bands.

    ``core`` must provide ``reset(N, mask, tgt, grad)``, ``sweeps_async(k)``,
    ``finish_async()``, ``fetch() -> (uint8 slab image, err[3])``,
    ``set_row_window(lo, hi)``, ``state()`` and ``rows_view(lo, hi) -> list of
    torch tensors`` (views of the CURRENT state rows, one contiguous tensor per
    channel plane) -- ``fpie_b200.band.CudaBandCore`` adapts ``GridSolver``.
    ``dist`` is ``torch.distributed`` (or any object with the same
    ``batch_isend_irecv / P2POp / isend / irecv / all_reduce`` surface).
    The default halo of 24 rows is a multiple of both blocking depths the
    solver picks by itself (8 and 12 sweeps per pass), so an exchange interval
    never ends in a short pass.
    """

    def __init__(self, core, dist, group=None, halo: int = 24, overlap: bool = True):
        self.core, self.dist, self.group = core, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.halo = int(halo)
        self.plan: BandPlan | None = None
        self.overlap = bool(overlap)
        self._split = False  # passes are split into edge / interior tiles and the exchange overlaps the interior

    @property
    def exchange_overlaps(self) -> bool:
        """True when passes are split and the halo exchange runs beside their interior tiles."""
        return self._split

    # -- reference interface -------------------------------------------------
    def reset(self, N, mask, tgt, grad) -> None:
        """Every rank passes the same global arrays (what ``sync()`` leaves on each
        MPI rank in the reference, mpi/grid.cc:34-54) and keeps only its slab."""
        mask = np.asarray(mask)
        self.plan = make_plan(mask.shape[0], self.world, self.rank, self.halo)
        p = self.plan
        if p.band_hi == p.band_lo:
            self._empty = True
            self._split = False
            return
        self._empty = False
        sl = slice(p.slab_lo, p.slab_hi)
        self.core.reset(int(p.slab_rows * mask.shape[1]), np.ascontiguousarray(mask[sl]), tgt[sl], grad[sl])
        self.core.set_row_window(*p.local_band)
        self._plan_overlap()

    def reset_slab(self, n_rows: int, src_slab, mask_slab, tgt_slab, gradient: str) -> BandPlan:
        """Each rank passes only its own slab of the uint8 images (rows
        ``plan.slab_lo:plan.slab_hi`` of the global crop, see ``make_plan``)."""
        self.plan = make_plan(n_rows, self.world, self.rank, self.halo)
        p = self.plan
        self._empty = p.band_hi == p.band_lo
        if not self._empty:
            if src_slab.shape[0] != p.slab_rows:
                raise ValueError(f"rank {self.rank}: slab has {src_slab.shape[0]} rows, plan needs {p.slab_rows}")
            self.core.reset_slab(src_slab, mask_slab, tgt_slab, gradient)
            self.core.set_row_window(*p.local_band)
        self._plan_overlap()
        return p

    def _plan_overlap(self) -> None:
        """Split every pass into edge and interior tiles when the core can (CUDA) and there is a
        neighbour: the exchange then runs on a second stream beside the interior of a pass."""
        p = self.plan
        self._split = (self.overlap and not self._empty and (p.up is not None or p.down is not None)
                       and hasattr(self.core, "pass_async"))
        if self._split:
            # edge tiles = those that write the rows we send (2 * halo from the slab edge) or read the halo
            # rows we receive (their load region reaches block_k rows beyond what they store)
            self.core.set_edge_rows(2 * self.halo + self.core.block_k)

    def sync(self) -> None:
        self.dist.barrier(self.group)

    def sweeps(self, iteration: int) -> None:
        """``iteration`` Jacobi sweeps, halos refreshed every ``halo`` sweeps."""
        left = int(iteration)
        while left > 0:
            s = min(self.halo, left)
            if self._split:
                self._interval_overlapped(s)
            else:
                if not self._empty:
                    self.core.sweeps_async(s)
                self.exchange()
            left -= s
        self._wait_exchange()

    def _interval_overlapped(self, s: int) -> None:
        """``s <= halo`` sweeps as passes of ``block_k``.  The LAST pass runs its edge tiles first and
        hands the finished band-edge rows to the exchange while its interior tiles are still being
        swept; the FIRST pass of the next interval sweeps its interior before it needs the received
        halo rows.  Same arithmetic, same bits -- only the order of the tiles within a pass changes."""
        core = self.core
        k = core.block_k
        sizes = [min(k, s - i) for i in range(0, s, k)]
        for idx, ns in enumerate(sizes):
            first, last = idx == 0, idx == len(sizes) - 1
            if first and last:
                self._wait_exchange()
                core.pass_async(ns, core.EDGE)
                self._start_exchange()
                core.pass_async(ns, core.INTERIOR)
            elif first:
                core.pass_async(ns, core.INTERIOR)
                self._wait_exchange()
                core.pass_async(ns, core.EDGE)
            elif last:
                core.pass_async(ns, core.EDGE)
                self._start_exchange()
                core.pass_async(ns, core.INTERIOR)
            else:
                core.pass_async(ns, core.EDGE)
                core.pass_async(ns, core.INTERIOR)
            core.flip()

    def _start_exchange(self) -> None:
        """Start the halo exchange, ordered after the edge tiles of the pass just enqueued; it moves rows
        of the buffer that pass writes (current after ``flip``).  How "ordered after" and "beside the
        interior" are realised is the core's business (CUDA: an event and a second stream)."""
        with self.core.exchange_scope():
            self.exchange(which=self.core.next_buffer())

    def _wait_exchange(self) -> None:
        if self._split:
            self.core.wait_exchange()

    def step(self, iteration: int):
        """Returns ``(uint8 image of this rank's band [rows, m, 3], err[3])``; ``err``
        is the global residual (sum over bands, as mpi/grid.cc:90-99 sums on root)."""
        import torch

        self.sweeps(iteration)
        if self._empty:
            img, err = np.zeros((0, 0, 3), np.uint8), np.zeros(3, np.float32)
        else:
            self.core.finish_async()
            lo, hi = self.plan.local_band
            if hasattr(self.core, "fetch_rows"):  # download the band, not the neighbours' halo rows
                img, err = self.core.fetch_rows(lo, hi, self._band_buffer(hi - lo))
            else:
                slab_img, err = self.core.fetch()
                img = slab_img[lo:hi]
        total = torch.tensor(np.asarray(err, np.float64), device=self._reduce_device())
        self.dist.all_reduce(total, group=self.group)
        return img, total.cpu().numpy().astype(np.float32)

    def _band_buffer(self, rows: int):
        """Page-locked landing buffer for the band's uint8 image (None = let the core allocate a pageable one).
        Large bands only -- hundreds of megabytes per step at PCIe speed instead of pageable speed; the buffer is
        recycled, so ``step`` returns an array that stays valid until the next ``step`` of this solver."""
        shape = getattr(self.core, "shape", None)
        if shape is None or not hasattr(self.core, "pinned_empty"):
            return None
        want = (rows, shape[1], 3)
        if int(np.prod(want)) < (32 << 20):
            return None
        buf = getattr(self, "_band_img", None)
        if buf is None or buf.shape != want:
            buf = self._band_img = self.core.pinned_empty(want, np.uint8)
        return buf

    def band_state(self) -> np.ndarray:
        lo, hi = self.plan.local_band
        return self.core.state()[lo:hi]

    # -- halo exchange ---------------------------------------------------------
    def exchange(self, which=None) -> None:
        """Send the band's edge rows to the neighbours, receive their edge rows into the halo rows
        (``which``: state buffer to operate on, default the current one)."""
        p = self.plan
        view = self.core.rows_view if which is None else (lambda lo, hi: self.core.rows_view(lo, hi, which))
        if self._empty or (p.up is None and p.down is None):
            return
        dist = self.dist
        lo, hi = p.local_band
        ops = []
        if p.up is not None:
            h = lo  # halo rows actually held above the band (== self.halo away from the top edge)
            for send, recv in zip(view(lo, lo + h), view(0, h)):
                ops.append(dist.P2POp(dist.isend, send, p.up, self.group))
                ops.append(dist.P2POp(dist.irecv, recv, p.up, self.group))
        if p.down is not None:
            h = p.slab_rows - hi
            for send, recv in zip(view(hi - h, hi), view(hi, hi + h)):
                ops.append(dist.P2POp(dist.isend, send, p.down, self.group))
                ops.append(dist.P2POp(dist.irecv, recv, p.down, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def _reduce_device(self):
        return getattr(self.core, "torch_device", "cpu")


class P2PBandGridSolver(BandGridSolver):
    """``BandGridSolver`` whose halo exchange runs behind the C ABI (``csrc/halo.cu``): copy-engine peer
    copies over NVLink into the neighbour's receive box, flag words and stream-level waits instead of
    NCCL send/recv kernels, and the whole per-interval schedule (edge tiles, exchange, interior tiles)
    issued by one C call per ``sweeps``.  ``torch.distributed`` is the control plane only: the
    128-byte link descriptors at (re)connection, the barrier of ``sync`` and the 3-float ``err`` sum --
    so the group may be NCCL (one process per GPU) or gloo (CPU tensors; several processes sharing one
    GPU in the tests).  ``same_process=True``: the neighbours are solvers of this very process (threads),
    linked by raw device pointers instead of CUDA IPC handles.

    Needs every band to be non-empty (bands >= halo rows); ``make_band_solver`` falls back to the NCCL
    transport otherwise."""

    def __init__(self, core, dist, group=None, halo: int = 24, overlap: bool = True, same_process: bool = False):
        super().__init__(core, dist, group, halo, overlap)
        self.same_process = bool(same_process)
        self._split = False
        self.exchanges_done = 0

    @property
    def exchange_overlaps(self) -> bool:
        return self.plan is not None and (self.plan.up is not None or self.plan.down is not None)

    def describe(self) -> str:
        return ("copy-engine peer copies + stream memory operations behind the C ABI, overlapped with the interior "
                "tiles of the passes around the exchange")

    def _check_bands(self, n_rows: int) -> None:
        off = band_offsets(int(n_rows), self.world)
        if min(off[i + 1] - off[i] for i in range(self.world)) <= 0:  # (same verdict on every rank)
            raise ValueError("the p2p halo transport needs every rank to own rows (use transport='nccl')")

    def reset(self, N, mask, tgt, grad) -> None:
        self._check_bands(np.asarray(mask).shape[0])
        super().reset(N, mask, tgt, grad)

    def reset_slab(self, n_rows: int, src_slab, mask_slab, tgt_slab, gradient: str) -> BandPlan:
        self._check_bands(n_rows)
        return super().reset_slab(n_rows, src_slab, mask_slab, tgt_slab, gradient)

    def _plan_overlap(self) -> None:
        """(Re)connect the halo link after a reset.  Collective: every rank calls it."""
        p = self.plan
        solver = self.core.solver
        lo, hi = p.local_band
        changed = solver.halo_config(lo, hi)
        flag = self._control_tensor([1.0 if changed else 0.0])
        self.dist.all_reduce(flag, group=self.group)
        if float(flag.sum()) == 0.0:
            return  # same geometry as before on every rank: the link and its counters stay
        if not changed:
            solver.halo_config(lo, hi, force=True)
        mine = (solver.halo_export(solver.UP) if p.up is not None else None,
                solver.halo_export(solver.DOWN) if p.down is not None else None)
        blobs = self._all_gather_blobs(mine)
        if p.up is not None:  # my upper edge rows land in the upper neighbour's box for rows from below
            solver.halo_connect(solver.UP, blobs[p.up][1], self.same_process)
        if p.down is not None:
            solver.halo_connect(solver.DOWN, blobs[p.down][0], self.same_process)
        self.dist.barrier(self.group)

    def _control_tensor(self, values):
        import torch

        return torch.tensor(values, dtype=torch.float64, device=self._reduce_device())

    def _all_gather_blobs(self, mine):
        if hasattr(self.dist, "all_gather_object"):
            out = [None] * self.world
            self.dist.all_gather_object(out, mine, group=self.group)
            return out
        return self.dist.all_gather_py(mine)  # in-process stand-ins (tests)

    def _reduce_device(self):
        backend = getattr(self.dist, "get_backend", lambda *_: "nccl")(self.group)
        return "cpu" if backend == "gloo" else getattr(self.core, "torch_device", "cpu")

    def sweeps(self, iteration: int) -> None:
        solver = self.core.solver
        if self.same_process:
            # Several bands inside ONE process (tests): a stream that waits for a neighbour's rows stalls every
            # device-wide synchronisation of the process (cudaMalloc / cudaFree, legacy-stream work of torch),
            # and the neighbour's thread may be the one making that call -- a deadlock that cannot occur between
            # processes, whose contexts synchronise independently.  So the threads enqueue in lock-step and
            # leave no wait pending when they go back to host work.
            self.dist.barrier(self.group)
            solver.band_sweeps_async(int(iteration))
            solver.wait()
            self.dist.barrier(self.group)
        else:
            solver.band_sweeps_async(int(iteration))
        self.exchanges_done = solver.halo_exchanges()

    def exchange(self, which=None) -> None:  # (the base class's NCCL exchange is never used here)
        raise RuntimeError("P2PBandGridSolver exchanges inside band_sweeps_async")

    def close(self) -> None:
        pass


def make_band_solver(core_solver, dist, group=None, halo: int = 24, overlap: bool = True, transport: str = "p2p",
                     same_process: bool = False):
    """Row-band solver over ``fpie_b200.GridSolver`` ``core_solver`` with the halo transport named:
    ``"p2p"`` (default; the exchange behind the C ABI) or ``"nccl"`` (round 1: ``batch_isend_irecv`` on views
    of the solver's buffers)."""
    core = CudaBandCore(core_solver)
    if transport == "p2p":
        return P2PBandGridSolver(core, dist, group, halo, overlap, same_process)
    if transport != "nccl":
        raise ValueError("transport must be 'p2p' or 'nccl'")
    solver = BandGridSolver(core, dist, group, halo, overlap)
    solver.describe = lambda: ("NCCL send/recv, " + ("overlapped with the interior tiles of a pass (second stream)"
                                                     if solver.exchange_overlaps else "between passes"))
    solver.close = lambda: None
    return solver


def canonical_crop(mask: np.ndarray):
    """Host-side mask canonicalisation of the Processor (fpie/process.py:338-351), cheap uint8 work:
    threshold on the channel mean, clear the 1-pixel frame, bounding box +-1.
    Returns ``(mask_u8 [rows, cols] with 255 inside, (x0, x1, y0, y1))`` in mask coordinates."""
    m = np.asarray(mask)
    if m.ndim == 3:
        m = m.astype(np.uint16).sum(-1) >= 128 * m.shape[2]  # == mean(-1) >= 128, exactly
    else:
        m = m >= 128
    m = m.copy()
    m[0, :] = m[-1, :] = False
    m[:, 0] = m[:, -1] = False
    rows = np.flatnonzero(m.any(axis=1))
    cols = np.flatnonzero(m.any(axis=0))
    if rows.size == 0:
        raise RuntimeError("reset: the mask is empty after thresholding and clearing its 1-pixel frame")
    x0, x1, y0, y1 = int(rows[0]) - 1, int(rows[-1]) + 2, int(cols[0]) - 1, int(cols[-1]) + 2
    return (m[x0:x1, y0:y1].astype(np.uint8) * 255), (x0, x1, y0, y1)


class BandGridProcessor:
    """``GridProcessor`` interface over row bands (one process per GPU).

    The call sequence is the reference's MPI one (fpie/cli.py:34-57 with ``-b mpi``): every rank
    constructs the processor, ``reset`` is called with the images, ``sync()``, then ``step(n)`` in
    lock-step on every rank.  Unlike the reference (rank 0 resets, ``sync`` broadcasts the whole fp32
    problem, mpi/grid.cc:34-54) every rank takes the uint8 images and keeps only its slab; ``step``
    returns the full blended target on rank 0 and ``None`` elsewhere (process.py:388-395)."""

    def __init__(self, gradient: str = "max", core=None, dist=None, group=None, halo: int = 24,
                 transport: str = "nccl", same_process: bool = False):
        """``transport="p2p"``: the halo exchange behind the C ABI (``P2PBandGridSolver``; ``core`` must be a
        ``CudaBandCore``); ``"nccl"``: ``dist.batch_isend_irecv`` on views of the core's buffers (any core)."""
        self.gradient = gradient
        if transport == "p2p":
            self.solver = P2PBandGridSolver(core, dist, group, halo, same_process=same_process)
        elif transport == "nccl":
            self.solver = BandGridSolver(core, dist, group, halo)
        else:
            raise ValueError("transport must be 'p2p' or 'nccl'")
        self.dist, self.group = dist, group
        self.rank = self.solver.rank
        self.root = self.rank == 0

    def reset(self, src, mask, tgt, mask_on_src=(0, 0), mask_on_tgt=(0, 0)) -> int:
        src, tgt = np.asarray(src), np.asarray(tgt)
        crop_mask, (x0, x1, y0, y1) = canonical_crop(mask)
        n, m = crop_mask.shape
        so = (mask_on_src[0] + x0, mask_on_src[1] + y0)
        to = (mask_on_tgt[0] + x0, mask_on_tgt[1] + y0)
        for img, o, what in ((src, so, "source"), (tgt, to, "target")):
            if o[0] < 0 or o[1] < 0 or o[0] + n > img.shape[0] or o[1] + m > img.shape[1]:
                raise RuntimeError(f"reset: the mask bounding box falls outside the {what} image")
        plan = make_plan(n, self.solver.world, self.rank, self.solver.halo)
        sl = slice(plan.slab_lo, plan.slab_hi)
        self.solver.reset_slab(
            n,
            np.ascontiguousarray(src[so[0] : so[0] + n, so[1] : so[1] + m][sl]),
            np.ascontiguousarray(crop_mask[sl]),
            np.ascontiguousarray(tgt[to[0] : to[0] + n, to[1] : to[1] + m][sl]),
            self.gradient,
        )
        self.box = (to[0], to[0] + n, to[1], to[1] + m)
        self.tgt = np.array(tgt, dtype=np.uint8, copy=True) if self.root else None
        self._crop_mask = crop_mask
        return n * m

    def sync(self) -> None:
        self.solver.sync()

    def step(self, iteration: int):
        import torch

        band, err = self.solver.step(iteration)
        plan, world = self.solver.plan, self.solver.world
        off = band_offsets(plan.n_rows, world)
        m = self.box[3] - self.box[2]
        dev = self.solver._reduce_device()
        mine = torch.from_numpy(np.ascontiguousarray(band)).to(dev)
        # bands have different heights: gather into equal-size padded buffers
        tallest = max(off[i + 1] - off[i] for i in range(world))
        padded = torch.zeros((tallest, m, 3), dtype=torch.uint8, device=dev)
        padded[: mine.shape[0]] = mine
        parts = [torch.empty_like(padded) for _ in range(world)] if self.root else None
        self.dist.gather(padded, parts, dst=0, group=self.group)
        if not self.root:
            return None
        for i in range(world):
            rows = off[i + 1] - off[i]
            self._paste(off[i], off[i + 1], parts[i][:rows].cpu().numpy())
        return self.tgt, err

    def _paste(self, lo: int, hi: int, band_img: np.ndarray) -> None:
        """Rows [lo, hi) of the solved crop -> the full target (process.py:393)."""
        x0, _, y0, y1 = self.box
        self.tgt[x0 + lo : x0 + hi, y0:y1] = band_img


class BandEquProcessor(BandGridProcessor):
    """``EquProcessor`` interface over row bands.

    The EquSolver's gather is all-to-all for arbitrary ids (the reference's MPI EquSolver broadcasts
    the whole vector every few sweeps, mpi/equ.cc:136-146); on row-major ids, though, unknown i only
    reads unknowns of the rows above and below, and the system IS a grid problem: state ``X`` on the
    masked pixels and 0 elsewhere, gradient ``B`` (the form ``EquSolver`` promotes to on one GPU).
    Built directly from the uint8 slab on each rank, it shards by row bands exactly like the
    GridSolver and returns the EquSolver's numbers: same fp32 state, same uint8 image, ``reset``
    returns ``K + 1`` and ``step`` scatters only the K solved pixels (process.py:273-280)."""

    def __init__(self, gradient: str = "max", core=None, dist=None, group=None, halo: int = 24,
                 transport: str = "nccl", same_process: bool = False):
        super().__init__(gradient, core, dist, group, halo, transport, same_process)
        self.solver.core.set_formulation(True)

    def reset(self, src, mask, tgt, mask_on_src=(0, 0), mask_on_tgt=(0, 0)) -> int:
        super().reset(src, mask, tgt, mask_on_src, mask_on_tgt)
        return int(np.count_nonzero(self._crop_mask)) + 1  # max_id (process.py:190)

    def _paste(self, lo: int, hi: int, band_img: np.ndarray) -> None:
        x0, _, y0, y1 = self.box
        on = self._crop_mask[lo:hi] > 0
        self.tgt[x0 + lo : x0 + hi, y0:y1][on] = band_img[on]


class _DeviceRows:
    """``__cuda_array_interface__`` window onto solver-owned device memory."""

    def __init__(self, ptr: int, shape: tuple[int, ...]):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}


class CudaBandCore:
    """Adapts ``fpie_b200.GridSolver`` to the core protocol of ``BandGridSolver``."""

    def __init__(self, solver):
        import torch

        self.solver = solver
        self.torch_device = torch.device("cuda", solver.device)
        self._views = {}
        # the solver enqueues on the torch stream that was current when it was built
        h = solver.stream_handle
        self.compute_stream = (torch.cuda.ExternalStream(h, device=self.torch_device) if h
                               else torch.cuda.default_stream(self.torch_device))
        self.comm_stream = torch.cuda.Stream(self.torch_device)
        self._pending = None  # event of the exchange in flight

    def reset(self, N, mask, tgt, grad):
        self.solver.reset(N, mask, tgt, grad)
        self._views.clear()

    def reset_slab(self, src, mask, tgt, gradient):
        self.solver.reset_slab(src, mask, tgt, gradient)
        self._views.clear()

    def set_formulation(self, equ: bool):
        self.solver.set_formulation(equ)

    def sweeps_async(self, k):
        self.solver.sweeps_async(k)

    def finish_async(self):
        self.solver.finish_async()

    def fetch(self):
        return self.solver.fetch()

    def fetch_rows(self, lo, hi, img=None):
        return self.solver.fetch_rows(lo, hi, img)

    @property
    def shape(self):
        return self.solver.shape

    @staticmethod
    def pinned_empty(shape, dtype):
        from . import _lib

        return _lib.pinned_empty(shape, dtype)

    def state(self):
        return self.solver.state()

    def set_row_window(self, lo, hi):
        self.solver.set_row_window(lo, hi)

    def _planes(self, which: int):
        """torch view [3, rows_alloc_from_pad, pitch] of state buffer ``which`` (cached)."""
        import torch

        if which not in self._views:
            v = self.solver.band_view(which)
            n = self.solver.shape[0]
            planes = []
            for p in range(3):
                ptr = v["base"] + 4 * (p * v["plane"] + v["pad_rows"] * v["pitch"])
                planes.append(torch.as_tensor(_DeviceRows(ptr, (n, v["pitch"])), device=self.torch_device))
            self._views[which] = planes
        return self._views[which]

    def rows_view(self, lo: int, hi: int, which: int | None = None):
        return [plane[lo:hi] for plane in self._planes(self.solver.current_buffer() if which is None else which)]

    # -- split passes: halo exchange beside the interior of a pass --------------
    EDGE, INTERIOR = 0, 1

    @property
    def block_k(self) -> int:
        return self.solver.info()["block_k"]

    def set_edge_rows(self, rows: int):
        self.solver.set_edge_rows(rows)

    def pass_async(self, nsweeps: int, part: int):
        self.solver.pass_async(nsweeps, part)

    def flip(self):
        self.solver.flip()

    def next_buffer(self) -> int:
        return self.solver.current_buffer() ^ 1

    @contextlib.contextmanager
    def exchange_scope(self):
        """Work enqueued inside runs on the communication stream, after everything the solver's stream
        holds so far, and concurrently with what the solver enqueues next; ``wait_exchange`` joins."""
        import torch

        done = torch.cuda.Event()
        done.record(self.compute_stream)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(done)
            yield
            self._pending = torch.cuda.Event()
            self._pending.record(self.comm_stream)

    def wait_exchange(self) -> None:
        """The solver's stream waits (on the device) for the exchange in flight, if any."""
        if self._pending is not None:
            self.compute_stream.wait_event(self._pending)
            self._pending = None


def init_process_group_from_env(backend: str | None = None):
    """``torch.distributed`` set-up for ``torchrun`` launches (RANK / WORLD_SIZE /
    LOCAL_RANK / MASTER_ADDR / MASTER_PORT from the environment)."""
    import os

    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return dist

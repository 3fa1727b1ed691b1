#!/usr/bin/env python3
"""Benchmark of the Jacobi Poisson hot path (BASELINE.json metric: Gupd/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one ``step(iters)`` of the hot path on the workload: ``iters``
Jacobi sweeps + residual + uint8 conversion.  Workloads (BASELINE.json configs):

    cfg2  GridSolver, 4096x4096x3 circle mask, grad max, 5000 sweeps   (N=1 default)
    cfg3  EquSolver,  8192x8192x3 ring mask,   grad avg, 10000 sweeps
    cfg4  GridSolver, 32768x32768x3 square mask, row bands over N GPUs, 5000 sweeps (N>1 default)
    cfg1  EquSolver,  1026x1026x3 square mask, grad max, 5000 sweeps
    cfg5  GridSolver, 512 independent 256x256x3 patches (one mosaic), grad src, 5000 sweeps

Prints ONE JSON line (rank 0).  ``value`` is timed with CUDA events with the
inputs resident in HBM; ``e2e`` goes through the Processor API with host
buffers (H2D of the uint8 images and D2H of the uint8 result inside the timed
region); ``cpu_baseline`` is the reference's own OpenMP core (oracle/_ref) timed
on this box's host cores on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.join(ROOT, "fast-poisson-image-editing_b200")
for p in (ROOT, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

GRID_BYTES_PER_UPDATE = 36  # x read 12 + grad read 12 + x' write 12      
EQU_BYTES_PER_UPDATE = 52  # A 16 + B 12 + X read 12 + X' write 12        
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md

WORKLOADS = {
    "cfg1": dict(solver="equ", size=1026, mask="square", grad="max", iters=5000),
    "cfg2": dict(solver="grid", size=4096, mask="circle", grad="max", iters=5000),
    "cfg3": dict(solver="equ", size=8192, mask="ring", grad="avg", iters=10000),
    "cfg4": dict(solver="grid", size=32768, mask="square", grad="max", iters=5000),
    "cfg5": dict(solver="batch", size=256, batch=512, mask="square", grad="src", iters=5000),
}


def equ_kernel_name(info):
    """The gather kernel EquSolver.sweeps_async launches for this system (csrc/equ.cu): 4-byte distance table (+ fp16
    B stream) for row-major systems beyond the L2, 8-byte (up, down) table for smaller ones, int4 table otherwise."""
    if info.get("path") != "gather-compact":
        return "equ_sweep_kernel"
    if (info.get("table") or "").startswith("delta16"):  # persistent + pipelined unless switched off (A/B)
        return "equ_sweep_d16_kernel" if os.environ.get("FPIE_B200_D16_PIPE", "")[:1] == "0" else "equ_sweep_d16p_kernel"
    return "equ_sweep_lr_kernel"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, device: int, period: float = 0.1):
        self.device, self.period = device, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return self
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def pinned_copy(arr: np.ndarray) -> np.ndarray:
    """A copy of ``arr`` in page-locked host memory (numpy view of a torch pinned tensor)."""
    import torch

    t = torch.empty(arr.shape, dtype=getattr(torch, str(arr.dtype)), pin_memory=True)
    out = t.numpy()
    out[...] = arr
    _PINNED.append(t)
    return out


_PINNED = []


# ---------------------------------------------------------------------------
# CPU baseline (reference OpenMP core, or the C restatement when it is absent)
# ---------------------------------------------------------------------------
def ref_cuda_time(solver_kind, system, unknowns, iters):
    """The reference's own CUDA backend (oracle/_ref/core_cuda: the unmodified sources compiled for sm_100a) on
    this GPU, same system, its published tuning (`-z 256`, docs/benchmark.md:56; `--grid-x 2 --grid-y 128`,
    :117), wall clock of `step(iters)` as the reference CLI times it (cli.py:46-61).  A throughput comparator
    only: it updates in place without synchronisation (cuda/equ.cu:193-197, grid.cu:138-142), so its results
    are not Jacobi.  None when the module is absent or no GPU is visible."""
    from oracle import c_oracle

    try:
        import torch

        if not torch.cuda.is_available():
            return None
        core_cuda = c_oracle.load_reference_core("core_cuda")
        if core_cuda is None:
            return None
        # its constructors printf a device table (fpie/core/cuda/utils.cu:5-22) -- on file descriptor 1, where
        # this script owes exactly one JSON line
        sys.stdout.flush()
        saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        try:
            os.dup2(devnull, 1)
            ref = core_cuda.GridSolver(2, 128) if solver_kind == "grid" else core_cuda.EquSolver(256)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
        ref.reset(*system)
        ref.step(10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref.step(iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": unknowns * iters / dt / 1e9, "unit": "Gupd/s", "seconds": dt, "sweeps": iters,
                "what": f"reference core_cuda.{'GridSolver(2, 128)' if solver_kind == 'grid' else 'EquSolver(256)'} "
                        "(unmodified fpie/core/cuda built for sm_100a), same system, wall clock of step()"}
    except Exception as exc:  # a comparator must never take the bench line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def cpu_baseline(work, src, mask, tgt, budget_s: float = 12.0, steps: int = 1, with_ref_cuda: bool = False):
    """Time the reference's CPU implementation on a bounded sample of the workload.
    Returns (dict for the JSON line, list of per-step seconds, sweeps per step)."""
    from oracle import c_oracle, np_oracle

    cores = os.cpu_count() or 1
    core = c_oracle.load_reference_core("core_openmp")
    kind = "reference" if core is not None else "port"
    ref_cuda = None
    if work["solver"] == "grid":
        m, t, g, _ = np_oracle.grid_system(src, mask, tgt, (0, 0), (0, 0), work["grad"])
        unknowns = int(m.sum())
        if with_ref_cuda:
            ref_cuda = ref_cuda_time("grid", (m.size, m, t, g), unknowns, min(work["iters"], 2000))
        if core is not None:
            solver = core.GridSolver(2, 16, cores)  # published tuning, docs/benchmark.md:111
            solver.reset(m.size, m, t, g)
            run = solver.step
        else:
            state = {"t": t}

            def run(k):
                state["t"] = c_oracle.grid_sweeps(m, state["t"], g, k, threads=cores)
    else:
        n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), work["grad"])
        unknowns = n - 1
        if with_ref_cuda:
            ref_cuda = ref_cuda_time("equ", (n, A, X, B), unknowns, min(work["iters"], 2000))
        # the reference's OpenMP EquSolver is red-black Gauss-Seidel (openmp/equ.cc:107-118): same
        # memory traffic per sweep, so it is the throughput baseline; the Jacobi port checks results.
        if core is not None:
            solver = core.EquSolver(cores)
            solver.reset(n, A, X, B)
            run = solver.step
        else:
            state = {"x": X}

            def run(k):
                state["x"] = c_oracle.equ_sweeps(A, state["x"], B, k, threads=cores)

    t0 = time.perf_counter()
    run(2)
    per_sweep = (time.perf_counter() - t0) / 2
    sweeps = int(max(2, min(work["iters"], budget_s / max(per_sweep, 1e-9) / max(steps, 1))))
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run(sweeps)
        times.append(time.perf_counter() - t0)
    best = min(times)
    value = unknowns * sweeps / best / 1e9
    info = {
        "value": value,
        "unit": "Gupd/s",
        "cores": cores,
        "kind": kind,
        "sample": f"{sweeps} of {work['iters']} sweeps on the full {work['size']}^2 {work['mask']} workload "
        f"({'core_openmp from oracle/_ref' if kind == 'reference' else 'oracle/jacobi_oracle.c'}, {cores} threads)",
    }
    if with_ref_cuda:
        info["_ref_cuda"] = ref_cuda
    return info, times, sweeps, unknowns


# ---------------------------------------------------------------------------
def run_reference(args, work, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fpie_b200 import synth

    size = work["size"]
    if name == "cfg4":
        size = 8192  # the reference overflows int32 offsets at 3*N*M >= 2^31 (base_solver.h:114-117); 8192^2 keeps the host-side numpy setup to about a minute, throughput per pixel is size-independent
    src, mask, tgt = synth.make_problem(work["mask"], size, size, seed=0)
    w = dict(work, size=size)
    total = args.steps + args.warmup
    info, times, sweeps, unknowns = cpu_baseline(w, src, mask, tgt, budget_s=90.0, steps=total)
    timed = times[args.warmup :]
    sec = sum(timed) / len(timed)
    value = unknowns * sweeps / sec / 1e9
    info["value"] = value
    line = {
        "impl": "reference",
        "metric": "jacobi_gupd_per_s",
        "value": value,
        "unit": "Gupd/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": sec * 1e3,
        "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{name}: {w['solver']} {size}x{size}x3 {w['mask']} mask, grad {w['grad']}",
                   "sweeps_per_step": sweeps, "unknowns": unknowns},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": "Gupd/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
def run_single(args, work, name):
    import torch

    import fpie_b200
    from fpie_b200 import synth

    dev = 0
    torch.cuda.set_device(dev)
    size, iters = (args.size or work["size"]), (args.iters or work["iters"])
    is_batch = work["solver"] == "batch"
    is_grid = work["solver"] in ("grid", "batch")
    if is_batch:
        rng = np.random.default_rng(0)
        nb = work["batch"]
        src = rng.integers(0, 256, (nb, size, size, 3), dtype=np.uint8)
        tgt = rng.integers(0, 256, (nb, size, size, 3), dtype=np.uint8)
        mask = np.full((nb, size, size), 255, np.uint8)
        proc = fpie_b200.BatchGridProcessor(work["grad"], "b200", device=dev, block_k=args.block_k)
        reset_args = (src, mask, tgt)
    else:
        src, mask, tgt = synth.make_problem(work["mask"], size, size, seed=0)
        Proc = fpie_b200.GridProcessor if is_grid else fpie_b200.EquProcessor
        # EquSolver: the configs name the index-mapped gather path; "jacobi" would let the solver promote
        # this (row-major, self-labelled) system to the tiled grid kernel -- reported separately below
        kw = dict(block_k=args.block_k) if is_grid else dict(mode=args.equ_mode)
        proc = Proc(work["grad"], "b200", device=dev, **kw)
        reset_args = (src, mask, tgt, (0, 0), (0, 0))
    Proc = type(proc)
    t0 = time.perf_counter()
    nvars = proc.reset(*reset_args)
    reset_s = time.perf_counter() - t0
    core = proc.core
    info0 = core.info()
    unknowns = info0["unknowns"]

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def device_step():
        ev[0].record()
        core.sweeps_async(iters)
        ev[1].record()
        core.finish_async()
        ev[2].record()

    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    sweep_ms, total_ms = [], []
    launches0 = core.info()["launches"]
    with ClockSampler(dev) as clocks:
        for _ in range(args.steps):
            torch.cuda.synchronize()
            device_step()
            torch.cuda.synchronize()
            sweep_ms.append(ev[0].elapsed_time(ev[1]))
            total_ms.append(ev[0].elapsed_time(ev[2]))
    launches = core.info()["launches"] - launches0
    ms_per_step = float(np.mean(total_ms))
    value = unknowns * iters / (ms_per_step * 1e-3) / 1e9

    # dominant kernel: the sweep kernel; algorithmic bytes per launch / mean launch duration
    per_update = GRID_BYTES_PER_UPDATE if is_grid else EQU_BYTES_PER_UPDATE
    k = info0.get("block_k", 1) if is_grid else 1
    kernel_cfg = ({"tile": list(info0["tile"]), "rows_per_thread": info0["rows_per_thread"], "warps": info0["warps"],
                   "ctas_per_sm": info0["ctas_per_sm"], "variant": info0["variant"]} if is_grid else None)
    sweep_launches = (iters + k - 1) // k
    launch_s = float(np.mean(sweep_ms)) * 1e-3 / sweep_launches
    sweeps_per_launch = iters / sweep_launches
    achieved = per_update * unknowns * sweeps_per_launch / launch_s / 1e9
    peak, peak_src = measured_peak()
    kernel = "grid_sweepk_pipe_kernel" if is_grid else equ_kernel_name(core.info())
    patch = core.patch_info() if is_grid else None
    if patch and patch["launches"] > 0:
        # the persistent small-image kernel ran: ONE launch holds all sweeps of a step (csrc/patch.cuh)
        kernel, sweep_launches, sweeps_per_launch = "grid_patch_kernel", 1, iters
        launch_s = float(np.mean(sweep_ms)) * 1e-3
        achieved = per_update * unknowns * sweeps_per_launch / launch_s / 1e9
        kernel_cfg = {"kernel": "grid_patch_kernel", "rows_per_thread": patch["rows_per_thread"],
                      "cols_per_thread": patch["cols_per_thread"], "ctas_per_cluster": patch["cluster"],
                      "clusters": patch["clusters"], "warps": 8}
    roofline = {
        "bound": "hbm",
        "kernel": kernel,
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        # (the ncu capture is of config 2 for the grid kernel and of config 3 for the gather kernel)
        "traffic": profiled_traffic(kernel)
        if (name == ("cfg2" if is_grid else "cfg3") and not args.size) else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": per_update * unknowns * sweeps_per_launch,
        "launch_us": launch_s * 1e6,
        "note": "effective (algorithmic) bandwidth; with k sweeps fused per launch real DRAM traffic is ~1/k of it",
    }
    if roofline["traffic"]:
        # the profiled DRAM bytes of one launch of exactly this workload over the live launch time: how busy HBM is
        roofline["dram_achieved"] = roofline["traffic"] / launch_s / 1e9
        roofline["dram_frac"] = roofline["dram_achieved"] / peak

    if is_grid:
        # second view of the same kernel: with k sweeps per pass the binding resource is instruction issue /
        # the fp32 pipe, not DRAM.  Useful work = 4 FFMA (8 flop) per update and channel; peak = SMs x 128
        # lanes x 2 flop x the SM clock seen during the timed region.
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        roofline["compute"] = {
            "bound": "fp32-fma",
            "achieved": value * 3 * 8 / 1e3,
            "peak": sm_count * 128 * 2 * clocks.summary()["sm_mhz"] / 1e6 if clocks.summary().get("sm_mhz") else None,
            "unit": "TFLOP/s",
            "note": "useful flops only (the halo recomputation of temporal blocking, x1.26 at k = 8, is not counted)",
        }
        if roofline["compute"]["peak"]:
            roofline["compute"]["frac"] = roofline["compute"]["achieved"] / roofline["compute"]["peak"]

    # end to end through the Processor API with pinned host buffers
    psrc, pmask, ptgt = pinned_copy(src), pinned_copy(mask), pinned_copy(tgt)
    e2e_s, e2e_reset_s = [], []
    e2e_warm = 2  # (the Processor's page-locked canvas pool reaches its steady state after two resets)
    for i in range(e2e_warm + max(2, min(args.steps, 3))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proc.reset(psrc, pmask, ptgt, *reset_args[3:])
        t1 = time.perf_counter()  # (reset returns after its own device work: no extra synchronisation added)
        out, err = proc.step(iters)
        torch.cuda.synchronize()
        if i >= e2e_warm:
            e2e_s.append(time.perf_counter() - t0)
            e2e_reset_s.append(t1 - t0)
    e2e_val = unknowns * iters / float(np.mean(e2e_s)) / 1e9
    crop_bytes = int(np.prod(core.batch_shape if is_batch else core.shape if is_grid else core.crop_shape)) * 3
    e2e = {
        "value": e2e_val,
        "unit": "Gupd/s",
        "h2d_bytes_per_step": int(src.nbytes + mask.nbytes + tgt.nbytes),
        "d2h_bytes_per_step": crop_bytes + 12,
        "ms_per_step": float(np.mean(e2e_s)) * 1e3,
        "reset_ms": float(np.mean(e2e_reset_s)) * 1e3,
        "api": f"fpie_b200.{Proc.__name__}.reset(src, mask, tgt) + step({iters}) on pinned host uint8 images",
    }

    base = ref_cuda = None
    if not args.no_cpu_baseline:
        if is_batch:  # one patch after the other on the host, as the reference GUI does; sample = first patches
            w1 = dict(work, solver="grid")
            base, _, _, _ = cpu_baseline(w1, src[0], mask[0], tgt[0], budget_s=args.cpu_budget, with_ref_cuda=True)
            base["sample"] = "patch 0 of the batch: " + base["sample"]
        else:
            base, _, _, _ = cpu_baseline(work, src, mask, tgt, budget_s=args.cpu_budget, with_ref_cuda=True)
        ref_cuda = base.pop("_ref_cuda", None)

    line = {
        "metric": "jacobi_gupd_per_s",
        "value": value,
        "unit": "Gupd/s",
        "n_gpus": 1,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": f"{name}: {work['solver']} solver, {(str(work['batch']) + ' patches of ') if is_batch else ''}"
            f"{size}x{size}x3 {work['mask']} mask, grad {work['grad']}, {iters} sweeps per step",
            "unknowns": unknowns,
            "n_vars": nvars,
            "sweeps_per_step": iters,
            "block_k": k,
            "kernel": kernel_cfg,
            "gpx_per_s": (nvars * iters / (ms_per_step * 1e-3) / 1e9) if is_grid and not is_batch else None,
            "tiles": [info0.get("active_tiles"), info0.get("total_tiles")] if is_grid else None,
            "l2": (f"working set {unknowns * per_update / 1e6:.0f} MB per sweep exceeds the 126 MB L2; no flush needed"
                   if unknowns * per_update > 2 * 126e6 else
                   "working set fits the 126 MB L2 (L2-resident configuration; reported separately from the HBM runs)"),
            "reset_s": reset_s,
            "solver_path": core.info().get("path") if not is_grid else "tiled",
            "equ_table": core.info().get("table") if not is_grid else None,
        },
        "roofline": roofline,
        "cpu_baseline": base,
        # the reference's existing CUDA backend on the same GPU, same run (north star: "... and the reference's
        # existing CUDA backend"); for the batch workload: one patch, as the reference GUI would solve it
        "ref_cuda": ref_cuda,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    print(json.dumps(line))


CHUNK_ROWS = 1024


def global_rows(kind: str, lo: int, hi: int, m: int, seed: int = 0) -> np.ndarray:
    """Rows [lo, hi) of the ONE synthetic image `kind` ("src" / "tgt") of a row-band sharded problem.  Every
    chunk of CHUNK_ROWS rows is drawn from its own generator seeded by (seed, kind, chunk index), so any rank
    reproduces any row range of the same global image: neighbouring slabs agree on the rows they share."""
    out = np.empty((hi - lo, m, 3), np.uint8)
    tag = {"src": 1, "tgt": 2}[kind]
    for c in range(lo // CHUNK_ROWS, (hi + CHUNK_ROWS - 1) // CHUNK_ROWS):
        r0 = c * CHUNK_ROWS
        chunk = np.random.default_rng([seed, tag, c]).integers(0, 256, size=(CHUNK_ROWS, m, 3), dtype=np.uint8)
        a, b = max(lo, r0), min(hi, r0 + CHUNK_ROWS)
        out[a - lo : b - lo] = chunk[a - r0 : b - r0]
    return out


def slab_images(work, lo, hi, n, m):
    """uint8 rows [lo, hi) of the global n x m blend (src, canonical crop mask, tgt) and its unknown count."""
    from fpie_b200 import synth

    src = global_rows("src", lo, hi, m)
    tgt = global_rows("tgt", lo, hi, m)
    if work["mask"] == "square":
        mask = np.full((hi - lo, m), 255, np.uint8)
        mask[:, 0] = mask[:, -1] = 0
        if lo == 0:
            mask[0] = 0
        if hi == n:
            mask[-1] = 0
        unknowns = (n - 2) * (m - 2)
    else:
        full = synth.make_mask(work["mask"], n, m)
        full[0] = full[-1] = 0
        full[:, 0] = full[:, -1] = 0
        mask = np.ascontiguousarray(full[lo:hi])
        unknowns = int((full > 127).sum())
    return src, mask, tgt, unknowns


def band_parity_check(work, world, rank, dev, dist, halo, block_k, overlap, transport):
    """The shipped multi-GPU path against one GPU, in the same run: a reduced problem of the same kind
    (world bands of 1024 x 8192 pixels, 96 sweeps = 4 halo exchanges) is solved by the row-band solver over the
    same transport as the timed run, and -- independently, on every rank's own GPU -- as one unsharded grid.
    Each rank compares its band of the fp32 state bit for bit; err must agree to 1e-4 relative."""
    import torch

    import fpie_b200
    from fpie_b200 import band

    m, rows_per_band, sweeps = 8192, 1024, 96
    n = rows_per_band * world
    plan = band.make_plan(n, world, rank, halo)
    src, mask, tgt, unknowns = slab_images(work, plan.slab_lo, plan.slab_hi, n, m)
    core = fpie_b200.GridSolver(8, 8, device=dev, block_k=block_k)
    solver = band.make_band_solver(core, dist, halo=halo, overlap=overlap, transport=transport)
    solver.reset_slab(n, src, mask, tgt, work["grad"])
    band_img, err = solver.step(sweeps)
    band_state = solver.band_state()
    exchanges = solver.exchanges_done if hasattr(solver, "exchanges_done") else None
    solver.close()
    core.close()
    # the same problem on this rank's GPU alone
    fsrc, fmask, ftgt, _ = slab_images(work, 0, n, n, m)
    one = fpie_b200.GridSolver(8, 8, device=dev, block_k=block_k)
    one.reset_slab(fsrc, fmask, ftgt, work["grad"])
    img1, err1 = one.step(sweeps)
    state1 = one.state()[plan.band_lo : plan.band_hi]
    one.close()
    ok_state = bool(np.array_equal(band_state.view(np.uint32), state1.view(np.uint32)))
    ok_img = bool(np.array_equal(band_img, img1[plan.band_lo : plan.band_hi]))
    err_rel = float(np.max(np.abs(err.astype(np.float64) - err1) / np.maximum(np.abs(err1), 1e-30)))
    flags = torch.tensor([int(ok_state), int(ok_img)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    worst = torch.tensor([err_rel], device="cuda", dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return {
        "state_bit_exact": bool(flags[0].item()),
        "u8_equal": bool(flags[1].item()),
        "err_rel": float(worst.item()),
        "err_tol": 1e-4,
        "problem": f"{n}x{m} {work['mask']} mask, {world} bands, {sweeps} sweeps, halo {halo}, vs one unsharded GPU "
                   f"solve of the same images on every rank",
        "exchanges": exchanges,
        "unknowns": unknowns,
    }


def single_gpu_reference(work, n, m, iters, dev, block_k, steps=1):
    """The multi-GPU workload on ONE GPU, measured in the same run (rank 0, the other ranks wait): what
    strong-scaling efficiency is computed against."""
    import torch

    import fpie_b200

    src, mask, tgt, unknowns = slab_images(work, 0, n, n, m)
    core = fpie_b200.GridSolver(8, 8, device=dev, block_k=block_k)
    core.reset_slab(src, mask, tgt, work["grad"])
    del src, mask, tgt
    info = core.info()
    k = info["block_k"]
    core.sweeps_async(20 * k)
    core.finish_async()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(steps):
        e0.record()
        core.sweeps_async(iters)
        core.finish_async()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    core.close()
    t = float(np.mean(ms))
    return {"n_gpus": 1, "value": unknowns * iters / (t * 1e-3) / 1e9, "unit": "Gupd/s", "ms_per_step": t,
            "steps": steps, "block_k": k, "variant": info["variant"],
            "source": "measured in this run on rank 0's GPU (same images, same sweeps, no sharding)"}


def run_band(args, work, name):
    """N > 1: the grid is cut into row bands, one process per GPU, deep halos between neighbours."""
    import torch

    import fpie_b200
    from fpie_b200 import band

    dist = band.init_process_group_from_env("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    n = m = args.size or work["size"]
    iters = args.iters or work["iters"]
    halo = args.halo
    plan = band.make_plan(n, world, rank, halo)
    src, mask, tgt, unknowns = slab_images(work, plan.slab_lo, plan.slab_hi, n, m)
    core = fpie_b200.GridSolver(8, 8, device=dev, block_k=args.block_k)
    solver = band.make_band_solver(core, dist, halo=halo, overlap=not args.no_overlap, transport=args.transport)
    t0 = time.perf_counter()
    solver.reset_slab(n, src, mask, tgt, work["grad"])
    torch.cuda.synchronize()
    reset_s = time.perf_counter() - t0

    def device_step():
        solver.sweeps(iters)
        core.finish_async()

    for _ in range(args.warmup):
        device_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = core.info()["launches"]
    times = []
    with ClockSampler(dev) as clocks:
        for _ in range(args.steps):
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            device_step()
            e1.record()
            torch.cuda.synchronize()
            dist.barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
            times.append(float(t.item()))
    launches = torch.tensor([core.info()["launches"] - launches0], device="cuda")
    dist.all_reduce(launches)
    if args.trace_intervals > 0 and args.transport == "p2p":
        # an untimed extra step with CUDA events around every phase of the first intervals (csrc/halo.cu tags)
        torch.cuda.synchronize()
        dist.barrier()
        core.halo_trace_begin(args.trace_intervals)
        device_step()
        trace = core.halo_trace()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"band_trace_rank{rank}.json"), "w") as f:
            json.dump({"rank": rank, "world": world, "halo": halo, "block_k": core.info()["block_k"],
                       "tags": {"1": "before edge tiles (single-pass interval)", "2": "after edge tiles of the last pass",
                                "3": "after interior tiles of the last pass", "4": "before interior tiles of the first pass",
                                "5": "after interior tiles of the first pass", "6": "after edge tiles of the first pass",
                                "10": "halo stream: peer copies start", "11": "halo stream: peer copies + flags issued done",
                                "20": "solver stream: wait for the neighbours' rows starts",
                                "21": "solver stream: rows received and copied into the halo"},
                       "events_ms": trace}, f)
        dist.barrier()
    ms_per_step = float(np.mean(times))
    value = unknowns * iters / (ms_per_step * 1e-3) / 1e9

    # end to end: every rank uploads its uint8 slab from pinned memory, solves, downloads its band
    psrc, pmask, ptgt = pinned_copy(src), pinned_copy(mask), pinned_copy(tgt)
    del src, mask, tgt
    e2e_t, e2e_reset = [], []
    for i in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        solver.reset_slab(n, psrc, pmask, ptgt, work["grad"])
        t1 = time.perf_counter()
        img, err = solver.step(iters)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0, t1 - t0], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i:
            e2e_t.append(float(dt[0].item()))
            e2e_reset.append(float(dt[1].item()))
    e2e_s = float(np.mean(e2e_t))
    h2d = torch.tensor([psrc.nbytes + pmask.nbytes + ptgt.nbytes], device="cuda", dtype=torch.float64)
    d2h = torch.tensor([(plan.band_hi - plan.band_lo) * m * 3 + 12], device="cuda", dtype=torch.float64)
    dist.all_reduce(h2d)
    dist.all_reduce(d2h)
    info = core.info()
    exchange_desc = solver.describe() if hasattr(solver, "describe") else "NCCL send/recv"
    solver.close()
    core.close()
    del psrc, pmask, ptgt, img
    _PINNED.clear()
    torch.cuda.empty_cache()

    # correctness of exactly this path, in this run (all ranks)
    parity = None
    if not args.no_parity:
        parity = band_parity_check(work, world, rank, dev, dist, halo, args.block_k, not args.no_overlap, args.transport)
    # the same workload on one GPU, in this run (rank 0; the others wait at the barrier)
    ref1 = None
    if rank == 0 and not args.no_scaling_reference:
        ref1 = single_gpu_reference(work, n, m, iters, dev, args.block_k)
    dist.barrier()
    base = ref_cuda = None
    if rank == 0 and not args.no_cpu_baseline:
        # the reference's OpenMP core overflows int32 offsets at 3*N*M >= 2^31 (base_solver.h:114-117): the host
        # baseline runs the same kind of problem at 8192^2 (per-pixel throughput is size-independent there)
        w8 = dict(work, size=8192)
        s8, m8, t8, _ = slab_images(w8, 0, 8192, 8192, 8192)
        base, _, _, _ = cpu_baseline(w8, s8, m8, t8, budget_s=args.cpu_budget, with_ref_cuda=True)
        base["sample"] += " -- 8192^2 instead of 32768^2: the reference core cannot index the full problem"
        ref_cuda = base.pop("_ref_cuda", None)

    k = info["block_k"]
    peak, peak_src = measured_peak()
    # per-GPU roofline of the sweep kernel on this rank's slab (rank 0 reports)
    band_unknowns = unknowns / world
    achieved = GRID_BYTES_PER_UPDATE * band_unknowns * iters / (ms_per_step * 1e-3) / 1e9
    if rank == 0:
        line = {
            "metric": "jacobi_gupd_per_s",
            "value": value,
            "unit": "Gupd/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong",
            # strong scaling of ONE fixed problem; its single-GPU figure is not what `--gpus 1` runs (config 2)
            "scaling_reference": ref1,
            "scaling_efficiency_vs_reference": (value / (world * ref1["value"])) if ref1 else None,
            "parity": parity,
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": f"{name}: grid solver, {n}x{m}x3 {work['mask']} mask, grad {work['grad']}, {iters} sweeps per "
                f"step, {world} row bands, halo {halo} rows exchanged every {halo} sweeps ({exchange_desc})",
                "unknowns": unknowns,
                "sweeps_per_step": iters,
                "block_k": k,
                "halo": halo,
                "transport": args.transport,
                "exchange": exchange_desc,
                "band_rows": plan.band_hi - plan.band_lo,
                "l2": "per-GPU slab far exceeds the 126 MB L2; no flush needed",
                "reset_s": reset_s,
                "images": "one global image pair, generated per row chunk from (seed, chunk) so that neighbouring slabs "
                          "hold identical rows: the sharded solve is the single-GPU solve",
            },
            "roofline": {
                "bound": "hbm", "kernel": "grid_sweepk_pipe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": profiled_traffic("grid_sweepk_pipe_kernel_band_k12"),
                "peak_source": peak_src,
                "note": "per-GPU effective bandwidth (36 B x unknowns of one band x sweeps / step time, halo "
                        "exchange and epilogue included); traffic = DRAM bytes of one 12-sweep pass over one of 8 bands "
                        "(ncu, one GPU, profiles/)",
            },
            "cpu_baseline": base,
            "ref_cuda": ref_cuda,  # (one GPU, 8192^2: the reference has no multi-GPU path)
            "e2e": {"value": unknowns * iters / e2e_s / 1e9, "unit": "Gupd/s", "h2d_bytes_per_step": int(h2d.item()),
                    "d2h_bytes_per_step": int(d2h.item()), "ms_per_step": e2e_s * 1e3,
                    "reset_ms": float(np.mean(e2e_reset)) * 1e3,
                    "h2d_gbs_per_gpu": float(h2d.item()) / world / max(float(np.mean(e2e_reset)), 1e-9) / 1e9,
                    "api": "BandGridSolver.reset_slab(uint8 slabs) + step() per rank, pinned host buffers; the band "
                           "rows only come back"},
            "gpu_launches": int(launches.item()),
            "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=0, help="override sweeps per step")
    ap.add_argument("--block-k", type=int, default=0)
    ap.add_argument("--no-overlap", action="store_true",
                    help="row bands: exchange halos between passes instead of beside the interior of a pass")
    ap.add_argument("--halo", type=int, default=24, help="halo depth (rows) of the row-band sharding")
    ap.add_argument("--size", type=int, default=0, help="override the image side")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--equ-mode", default="gather", choices=["gather", "jacobi", "redblack"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="row bands: halo rows by copy-engine peer copies behind the C ABI (p2p) or NCCL send/recv")
    ap.add_argument("--trace-intervals", type=int, default=0,
                    help="row bands: record a phase timeline of the first N exchange intervals of one extra step into "
                         "gpurun_out/band_trace_rank<r>.json (CUDA events; the extra step is not timed)")
    ap.add_argument("--no-parity", action="store_true", help="row bands: skip the in-run sharded-vs-single-GPU check")
    ap.add_argument("--no-scaling-reference", action="store_true",
                    help="row bands: skip the in-run single-GPU measurement of the same workload")
    args = ap.parse_args()
    name = args.workload or ("cfg2" if args.gpus == 1 else "cfg4")
    work = WORKLOADS[name]
    if args.impl == "reference":
        return run_reference(args, work, name)
    if args.gpus == 1 and "RANK" not in os.environ:
        return run_single(args, work, name)
    if work["solver"] != "grid":
        raise SystemExit("row-band sharding exists for the grid solver only (EquSolver: replicas, see DESIGN.md)")
    return run_band(args, work, name)


if __name__ == "__main__":
    main()

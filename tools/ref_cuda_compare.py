#!/usr/bin/env python3
"""Time the reference's own CUDA backend (oracle/_ref/core_cuda, compiled from the unmodified
reference for sm_100a) on this GPU, next to fpie_b200, on the same synthetic inputs.
The reference backend is a throughput comparator only: it updates in place without
synchronisation (cuda/equ.cu:193-197, grid.cu:138-142), so its results are not Jacobi."""
import argparse, contextlib, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]
import numpy as np
import torch
import fpie_b200
from fpie_b200 import synth
from oracle import c_oracle, np_oracle

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--mask", default="circle")
ap.add_argument("--iters", type=int, default=1000)
ap.add_argument("--solver", default="grid")
args = ap.parse_args()
core_cuda = c_oracle.load_reference_core("core_cuda")
assert core_cuda is not None, "oracle/_ref/core_cuda missing (make -C oracle ref-cuda)"
src, mask, tgt = synth.make_problem(args.mask, args.size, args.size, seed=0)
out = dict(size=args.size, mask=args.mask, solver=args.solver, iters=args.iters)
if args.solver == "grid":
    m, t, g, _ = np_oracle.grid_system(src, mask, tgt, (0, 0), (0, 0), "max")
    unknowns = int(m.sum())
    with contextlib.redirect_stdout(io.StringIO()):
        ref = core_cuda.GridSolver(2, 128)  # published tuning, docs/benchmark.md:117
    ref.reset(m.size, m, t, g)
    mine = fpie_b200.GridSolver(8, 8)
    mine.reset(m.size, m, t, g)
else:
    n, A, X, B, _ = np_oracle.equ_system(src, mask, tgt, (0, 0), (0, 0), "avg")
    unknowns = n - 1
    with contextlib.redirect_stdout(io.StringIO()):
        ref = core_cuda.EquSolver(256)  # docs/benchmark.md:56
    ref.reset(n, A, X, B)
    mine = fpie_b200.EquSolver(256)
    mine.reset(n, A, X, B)
for name, s in (("reference_cuda", ref), ("fpie_b200", mine)):
    s.step(10)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.step(args.iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name] = dict(seconds=round(dt, 4), gupd_per_s=round(unknowns * args.iters / dt / 1e9, 2))
out["unknowns"] = unknowns
out["speedup"] = round(out["fpie_b200"]["gupd_per_s"] / out["reference_cuda"]["gupd_per_s"], 2)
print(json.dumps(out))

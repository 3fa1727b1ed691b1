#!/usr/bin/env python3
"""Where does end-to-end time go for the EquSolver path?  EquProcessor.reset / step pieces."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]
import numpy as np, torch
import fpie_b200
from fpie_b200 import synth
from bench import pinned_copy

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1026
kind = sys.argv[2] if len(sys.argv) > 2 else "square"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
src, mask, tgt = synth.make_problem(kind, size, size, seed=0)
psrc, pmask, ptgt = pinned_copy(src), pinned_copy(mask), pinned_copy(tgt)
def T(label, fn, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"  {label:40s} {min(ts)*1e3:8.2f} ms (min of {n})", flush=True); return r
for mode in ("gather", "jacobi", "redblack"):
    print(f"mode {mode}: {size}^2 {kind}")
    proc = fpie_b200.EquProcessor("max", "b200", mode=mode)
    T("proc.reset (pinned)", lambda: proc.reset(psrc, pmask, ptgt, (0, 0), (0, 0)))
    T("core.reset_from_images (pinned)", lambda: proc.core.reset_from_images(psrc, pmask, ptgt, (0, 0), (0, 0), "max"))
    T("proc.step(0)", lambda: proc.step(0))
    T("proc.step(1)", lambda: proc.step(1))
    T(f"proc.step({iters})", lambda: proc.step(iters), n=2)
    T(f"core sweeps({iters})+wait", lambda: (proc.core.sweeps_async(iters), proc.core.wait()), n=2)
    print("  info", proc.core.info())

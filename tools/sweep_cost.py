#!/usr/bin/env python3
"""Separate per-pass overhead from per-sweep cost: time one pass of the
temporally blocked kernel with nsweeps = 1..k on a fixed tile geometry."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]
import numpy as np, torch
import fpie_b200
from fpie_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--mask", default="circle")
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--k", type=int, default=8)
args = ap.parse_args()
src, mask, tgt = synth.make_problem(args.mask, args.size, args.size, seed=0)
core = fpie_b200.GridSolver(8, 8, block_k=args.k, variant=args.variant)
core.reset_from_images(src, mask, tgt, (0, 0), (0, 0), "max")
info = core.info()
res = {}
for ns in range(1, args.k + 1):
    for _ in range(5):
        core.sweeps_async(ns)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        core.sweeps_async(ns)
    e1.record()
    torch.cuda.synchronize()
    res[ns] = e0.elapsed_time(e1) * 1e3 / 50
xs = np.array(list(res)); ys = np.array([res[k] for k in xs])
b, a = np.polyfit(xs, ys, 1)
print(json.dumps(dict(mask=args.mask, variant=args.variant, k=args.k, tiles=info["active_tiles"], us_by_nsweeps={int(k): round(v, 1) for k, v in res.items()},
                      overhead_us=round(a, 1), per_sweep_us=round(b, 2), gupd_at_k=round(info["unknowns"] * args.k / res[args.k] / 1e3, 1))))

#!/usr/bin/env python3
"""Sweep GridSolver kernel variants x blocking depth on one workload and print
Gupd/s for each (CUDA events, warm, inputs resident).  Run on the GPU box:

    python tools/tune_grid.py --size 4096 --mask circle --iters 2000 --variants 0,4,5 --ks 4,8,12,16
"""

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import fpie_b200  # noqa: E402
from fpie_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--mask", default="circle")
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--variants", default="0,4,2,3,5,6")
    ap.add_argument("--ks", default="4,6,8,10,12,16")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    src, mask, tgt = synth.make_problem(args.mask, args.size, args.size, seed=0)
    rows = []
    for v in [int(x) for x in args.variants.split(",")]:
        for k in [int(x) for x in args.ks.split(",")]:
            try:
                core = fpie_b200.GridSolver(8, 8, block_k=k, variant=v)
            except RuntimeError as e:
                print(f"variant {v} k {k}: {e}")
                continue
            core.reset_from_images(src, mask, tgt, (0, 0), (0, 0), "max")
            info = core.info()
            k = k or info["block_k"]  # k = 0: whatever the solver chose for this grid
            iters = args.iters // k * k
            core.sweeps_async(iters)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                core.sweeps_async(iters)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            gupd = info["unknowns"] * iters / (best * 1e-3) / 1e9
            row = dict(variant=v, k=k, gupd=round(gupd, 1), us_per_launch=round(best * 1e3 / (iters // k), 1),
                       active_tiles=info["active_tiles"], total_tiles=info["total_tiles"])
            rows.append(row)
            print(json.dumps(row), flush=True)
            core.close()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Persistent small-image kernel vs the mosaic path: Gupd/s for batches of full-square 256^2 patches.
    python tools/patch_bench.py [--sizes 256] [--batches 1,12,512] [--iters 5000]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]
import numpy as np, torch
import fpie_b200

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--batches", default="1,12,512")
ap.add_argument("--iters", type=int, default=5000)
ap.add_argument("--modes", default="4,8,off")
ap.add_argument("--out", default=None)
args = ap.parse_args()
rows = []
for b in [int(v) for v in args.batches.split(",")]:
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, (b, args.size, args.size, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, (b, args.size, args.size, 3), dtype=np.uint8)
    mask = np.full((b, args.size, args.size), 255, np.uint8)
    for mode in args.modes.split(","):
        os.environ.pop("FPIE_B200_PATCH", None); os.environ.pop("FPIE_B200_PATCH_ROWS", None)
        if mode == "off": os.environ["FPIE_B200_PATCH"] = "0"
        elif mode.startswith("force"):  # also for fewer items than the policy wants (single images)
            os.environ["FPIE_B200_PATCH"] = "2"
            if mode[5:]: os.environ["FPIE_B200_PATCH_ROWS"] = mode[5:]
        elif mode != "auto": os.environ["FPIE_B200_PATCH_ROWS"] = mode
        proc = fpie_b200.BatchGridProcessor("src", "b200")
        proc.reset(src, mask, tgt)
        core = proc.core
        unknowns = core.info()["unknowns"]
        core.sweeps_async(args.iters); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); core.sweeps_async(args.iters); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row = dict(batch=b, size=args.size, mode=mode, ms=round(best, 3), gupd=round(unknowns * args.iters / best / 1e6, 1),
                   us_per_sweep=round(best * 1e3 / args.iters, 3), patch=core.patch_info())
        rows.append(row); print(json.dumps(row), flush=True)
        core.close()
if args.out:
    json.dump(rows, open(args.out, "w"), indent=1)

#!/usr/bin/env python3
"""One small image through GridProcessor: persistent patch kernel (forced) vs the tiled kernel, per mask kind.
    python tools/single_image_bench.py [--sizes 160,200,256] [--kinds circle,star,square] [--iters 5000]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]
import torch
import fpie_b200
from fpie_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="160,200,256")
ap.add_argument("--kinds", default="circle,star,square")
ap.add_argument("--iters", type=int, default=5000)
args = ap.parse_args()
for size in [int(v) for v in args.sizes.split(",")]:
    for kind in args.kinds.split(","):
        src, mask, tgt = synth.make_problem(kind, size, size, seed=0)
        row = dict(size=size, kind=kind)
        for mode, env in (("patch", "2"), ("tiled", "0"), ("auto", None)):
            os.environ.pop("FPIE_B200_PATCH", None)
            if env is not None:
                os.environ["FPIE_B200_PATCH"] = env
            proc = fpie_b200.GridProcessor("max", "b200")
            proc.reset(src, mask, tgt, (0, 0), (0, 0))
            core = proc.core
            unknowns = core.info()["unknowns"]
            core.sweeps_async(args.iters); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); core.sweeps_async(args.iters); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            row[mode] = round(unknowns * args.iters / best / 1e6, 1)
            row[mode + "_us_per_sweep"] = round(best * 1e3 / args.iters, 3)
            if mode == "auto":
                row["auto_uses_patch"] = core.patch_info()["launches"] > 0
            core.close()
        print(json.dumps(row), flush=True)

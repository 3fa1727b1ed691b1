#!/usr/bin/env python3
"""Where does end-to-end time go?  GridProcessor.reset / step pieces on cfg2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]
import numpy as np, torch
import fpie_b200
from fpie_b200 import synth
sys.path.insert(0, ROOT)
from bench import pinned_copy

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
src, mask, tgt = synth.make_problem("circle", size, size, seed=0)
psrc, pmask, ptgt = pinned_copy(src), pinned_copy(mask), pinned_copy(tgt)
proc = fpie_b200.GridProcessor("max", "b200")
def T(label, fn, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"{label:40s} {min(ts)*1e3:8.2f} ms (min of {n})"); return r
T("proc.reset (pageable)", lambda: proc.reset(src, mask, tgt, (0, 0), (0, 0)))
T("proc.reset (pinned)", lambda: proc.reset(psrc, pmask, ptgt, (0, 0), (0, 0)))
T("core.reset_from_images (pinned)", lambda: proc.core.reset_from_images(psrc, pmask, ptgt, (0, 0), (0, 0), "max"))
T("np.array(tgt, copy=True)", lambda: np.array(ptgt, copy=True))
T("proc.step(0)", lambda: proc.step(0))
T("core.step(0)", lambda: proc.core.step(0))
T("core.finish_async + wait", lambda: (proc.core.finish_async(), proc.core.wait()))
img = np.empty(proc.core.shape + (3,), np.uint8)
T("core.fetch pageable", lambda: proc.core.fetch(img))
pimg = pinned_copy(img)
T("core.fetch pinned", lambda: proc.core.fetch(pimg))
T("core.sweeps(5000)+wait", lambda: (proc.core.sweeps_async(5000), proc.core.wait()), n=2)
T("proc.step(5000)", lambda: proc.step(5000), n=2)

#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED Python reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box): ``python tests/golden/make_golden.py``.  It imports
``fpie.process.{EquProcessor,GridProcessor}`` with ``backend="numpy"`` from
``/root/reference`` and records, for a handful of small cases, the inputs, the
core-level system the Processor hands to the solver, the fp32 solver state, the
uint8 image and ``err`` after each ``step`` call.  The result is committed as
``tests/golden/fpie_numpy_golden.npz`` and is what pins ``oracle/`` (and,
through it, the CUDA path) to the reference.
"""

import hashlib
import os
import sys

import numpy as np

REF = os.environ.get("FPIE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from fpie.process import EquProcessor, GridProcessor  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def disk(h, w, cy, cx, r):
    yy, xx = np.mgrid[0:h, 0:w]
    return ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.uint8) * 255


def ring(h, w, r_in, r_out):
    yy, xx = np.mgrid[0:h, 0:w]
    d2 = (yy - h / 2) ** 2 + (xx - w / 2) ** 2
    return ((d2 <= r_out**2) & (d2 >= r_in**2)).astype(np.uint8) * 255


def cases():
    out = {}
    # 1. the reference's own 6x6 smoke fixture (tests/test_smoke.py:19-24)
    src = np.zeros((6, 6, 3), np.uint8)
    mask = np.zeros((6, 6), np.uint8)
    mask[2:4, 2:4] = 255
    tgt = np.ones((6, 6, 3), np.uint8) * 10
    out["smoke6"] = dict(src=src, mask=mask, tgt=tgt, off_src=(0, 0), off_tgt=(0, 0), steps=(2,))

    # 2. the reference's own parity fixture (tests/test_smoke.py:48-66)
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, size=(24, 24, 3), dtype=np.uint8)
    mask = np.zeros((24, 24), np.uint8)
    mask[2:-2, 2:-2] = (rng.random((20, 20)) > 0.35).astype(np.uint8) * 255
    out["rng24"] = dict(src=src, mask=mask, tgt=tgt, off_src=(0, 0), off_tgt=(0, 0), steps=(5,))

    # 3. different shapes, 3-channel mask, non-zero offsets, repeated step calls
    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, size=(48, 57, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, size=(70, 81, 3), dtype=np.uint8)
    m = ring(33, 41, 6.0, 14.5)
    mask = np.repeat(m[:, :, None], 3, axis=2)
    mask[:, :, 1] = np.where(mask[:, :, 1] > 0, 200, 30)  # mean(-1) decides (process.py:209-211)
    out["ring_off"] = dict(src=src, mask=mask, tgt=tgt, off_src=(5, 9), off_tgt=(21, 30), steps=(7, 6))

    # 4. mask touching its own border (frame gets cleared), odd sizes, holes
    rng = np.random.default_rng(11)
    src = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    tgt = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    mask = (rng.random((37, 53)) > 0.2).astype(np.uint8) * 255
    out["holes_full"] = dict(src=src, mask=mask, tgt=tgt, off_src=(0, 0), off_tgt=(0, 0), steps=(3, 20))

    # 5. a disk with saturating values (clip to 0 / 255 is exercised)
    rng = np.random.default_rng(13)
    src = (rng.integers(0, 2, size=(45, 45, 3)) * 255).astype(np.uint8)
    tgt = (rng.integers(0, 2, size=(60, 64, 3)) * 255).astype(np.uint8)
    mask = disk(45, 45, 22, 22, 19)
    out["disk_sat"] = dict(src=src, mask=mask, tgt=tgt, off_src=(0, 0), off_tgt=(8, 11), steps=(4, 9))
    return out


def run(kind, case, mode):
    Proc = EquProcessor if kind == "equ" else GridProcessor
    proc = Proc(gradient=mode, backend="numpy")
    captured = {}
    core = proc.core
    real_reset = core.reset

    def spy(n, a, b, c):
        # snapshot what the Processor hands to the core (numpy core aliases it)
        captured["n"] = int(n)
        captured["args"] = (np.array(a, copy=True), np.array(b, copy=True), np.array(c, copy=True))
        return real_reset(n, a, b, c)

    core.reset = spy
    n = proc.reset(case["src"], case["mask"], case["tgt"].copy(), tuple(case["off_src"]), tuple(case["off_tgt"]))
    rec = {"n": np.int64(n)}
    if kind == "equ":
        rec["A"], rec["X0"], rec["B"] = captured["args"]
    else:
        rec["mask_crop"], rec["tgt_crop"], rec["grad"] = captured["args"]
    for si, it in enumerate(case["steps"]):
        img, err = proc.step(it)
        rec[f"img{si}"] = np.array(img, copy=True)
        rec[f"err{si}"] = np.asarray(err, np.float32).copy()
        state = core.X if kind == "equ" else core.tgt
        rec[f"state{si}"] = np.array(state, np.float32, copy=True)
    return rec


def main():
    blob = {}
    for name, case in cases().items():
        for key in ("src", "mask", "tgt"):
            blob[f"{name}/{key}"] = case[key]
        blob[f"{name}/off_src"] = np.array(case["off_src"], np.int64)
        blob[f"{name}/off_tgt"] = np.array(case["off_tgt"], np.int64)
        blob[f"{name}/steps"] = np.array(case["steps"], np.int64)
        for kind in ("equ", "grid"):
            for mode in ("max", "src", "avg"):
                rec = run(kind, case, mode)
                for k, v in rec.items():
                    blob[f"{name}/{kind}/{mode}/{k}"] = v
                last = len(case["steps"]) - 1
                img = rec[f"img{last}"]
                print(
                    f"{name:10s} {kind:4s} {mode:3s} n={int(rec['n']):5d} err={rec[f'err{last}']} "
                    f"sha1={hashlib.sha1(img.tobytes()).hexdigest()[:16]} sum={int(img.sum())}"
                )
    path = os.path.join(HERE, "fpie_numpy_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

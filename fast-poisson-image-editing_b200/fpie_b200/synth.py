"""Seeded synthetic blending problems.

Shapes follow the reference's benchmark generators (tests/data.py:19-33:
an all-255 square and a filled circle) plus ring / star / random-holes masks
for the irregular EquSolver configurations.  ``src`` is drawn before ``tgt``
from ``default_rng(seed)``, the convention of tests/test_smoke.py:51-53.
"""

from __future__ import annotations

import numpy as np

MASK_KINDS = ("square", "circle", "ring", "star", "holes")


def make_mask(kind: str, h: int, w: int, seed: int = 0) -> np.ndarray:
    """uint8 ``[h, w]`` mask, 255 inside."""
    if kind == "square":
        return np.full((h, w), 255, np.uint8)
    yy = np.arange(h, dtype=np.float32)[:, None] - (h // 2)
    xx = np.arange(w, dtype=np.float32)[None, :] - (w // 2)
    if kind == "circle":
        r = min(h, w) // 2
        return ((yy * yy + xx * xx) <= float(r) * r).astype(np.uint8) * 255
    if kind == "ring":
        s = float(min(h, w))
        d2 = yy * yy + xx * xx
        return ((d2 >= (0.25 * s) ** 2) & (d2 <= (0.48 * s) ** 2)).astype(np.uint8) * 255
    if kind == "star":
        s = float(min(h, w))
        ang = np.arctan2(yy, xx)
        rad = np.sqrt(yy * yy + xx * xx)
        # 5-point star: boundary radius oscillates between inner and outer
        phase = np.abs(((ang * 5 / (2 * np.pi)) % 1.0) - 0.5) * 2.0  # 0 at a tip, 1 between tips
        edge = 0.48 * s - (0.48 - 0.19) * s * phase
        return (rad <= edge).astype(np.uint8) * 255
    if kind == "holes":
        rng = np.random.default_rng(seed + 1000003)
        return (rng.random((h, w)) > 0.35).astype(np.uint8) * 255
    raise ValueError(f"unknown mask kind {kind!r}")


def make_images(h: int, w: int, seed: int = 0, chunk_rows: int = 4096):
    """Random uint8 ``src`` then ``tgt`` of shape ``[h, w, 3]`` (src drawn first)."""
    rng = np.random.default_rng(seed)
    if h * w <= 64 << 20:
        src = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        tgt = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        return src, tgt
    # very large images: fill in row chunks to bound temporaries
    src = np.empty((h, w, 3), np.uint8)
    tgt = np.empty((h, w, 3), np.uint8)
    for img in (src, tgt):
        for r in range(0, h, chunk_rows):
            img[r : r + chunk_rows] = rng.integers(0, 256, size=(min(chunk_rows, h - r), w, 3), dtype=np.uint8)
    return src, tgt


def make_problem(kind: str, h: int, w: int, seed: int = 0):
    """``(src, mask, tgt)`` for one synthetic blend, all the same size."""
    src, tgt = make_images(h, w, seed)
    return src, make_mask(kind, h, w, seed), tgt

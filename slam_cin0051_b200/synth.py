"""Deterministic synthetic grayscale sequences of the shapes BASELINE.json names.

A textured canvas with a jittered grid of small dark / bright squares (corner-rich), cropped with a
per-frame translation so consecutive frames have true correspondences.  numpy PCG64, seed-stable.
"""
from __future__ import annotations

import numpy as np


def make_canvas(rows: int, cols: int, pitch_px: int = 14, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    H, W = rows + 128, cols + 128
    canvas = (110 + rng.integers(-6, 7, size=(H, W))).astype(np.uint8)
    ys = np.arange(pitch_px // 2, H - 8, pitch_px)
    xs = np.arange(pitch_px // 2, W - 8, pitch_px)
    for y in ys:
        jy = rng.integers(-1, 2, size=len(xs))
        jx = rng.integers(-1, 2, size=len(xs))
        side = rng.integers(3, 6, size=len(xs))
        dark = rng.integers(0, 2, size=len(xs)).astype(bool)
        vd = rng.integers(0, 50, size=len(xs))
        vb = rng.integers(190, 256, size=len(xs))
        for k, x in enumerate(xs):
            yy, xx = max(y + jy[k], 0), max(x + jx[k], 0)
            canvas[yy:yy + side[k], xx:xx + side[k]] = vd[k] if dark[k] else vb[k]
    return canvas


def make_sequence(rows: int, cols: int, n_frames: int, pitch_px: int = 14, seed: int = 0) -> np.ndarray:
    """(n_frames, rows, cols) uint8; frame f is the canvas cropped at (64 + 2f mod 32, 64 + f mod 16)."""
    canvas = make_canvas(rows, cols, pitch_px, seed)
    out = np.empty((n_frames, rows, cols), np.uint8)
    for f in range(n_frames):
        oy, ox = 64 + (f % 16), 64 + ((2 * f) % 32)
        out[f] = canvas[oy:oy + rows, ox:ox + cols]
    return out

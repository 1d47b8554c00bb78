"""Deterministic synthetic rectified stereo pairs (SURVEY.md section 8d "Synthetic inputs").

Pure numpy `default_rng(seed)` (PCG64): blurred-uniform texture, smooth sinusoidal disparity
field plus a raised slab, integer gather for the right view, +-2 noise.
"""
from __future__ import annotations

import zlib
import numpy as np


def make_pair(width: int, height: int, num_disp: int, min_disp: int = 0, seed: int = 1000):
    W, H, D = width, height, num_disp
    rng = np.random.default_rng(seed)
    pad = D + abs(min_disp) + 8
    T = rng.integers(0, 256, (H + 2, W + 2 * pad + 2)).astype(np.int32)
    acc = np.zeros((H, W + 2 * pad), np.int32)
    for dy in range(3):
        for dx in range(3):
            acc += T[dy:dy + H, dx:dx + W + 2 * pad]
    T = ((acc + 4) // 9).astype(np.uint8)
    x = np.arange(W, dtype=np.float64)[None, :]
    y = np.arange(H, dtype=np.float64)[:, None]
    d = min_disp + np.rint(0.10 * D + 0.55 * D * (0.5 + 0.5 * np.sin(2 * np.pi * x / (W / 3.0)) * np.cos(2 * np.pi * y / (H / 2.0))))
    d = d.astype(np.int64)
    d[H // 3:H // 2, W // 3:2 * W // 3] += int(0.2 * D)
    left = np.ascontiguousarray(T[:, pad:pad + W])
    cols = pad + np.arange(W, dtype=np.int64)[None, :] + d
    right = np.take_along_axis(T, cols, axis=1).astype(np.int32)
    right = right + rng.integers(-2, 3, right.shape)
    right = np.clip(right, 0, 255).astype(np.uint8)
    return left, np.ascontiguousarray(right)


def crc32(a: np.ndarray) -> str:
    return "%08x" % (zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF)


def sample_camera(width: int, height: int, seed: int = 0, strength: float = 1.0):
    """A plausible calibrated camera (K, D, R, P) for the rectification tests and bench: focal ~ width, mild plumb_bob distortion, small rotation."""
    rng = np.random.default_rng(seed)
    f = width * (0.9 + 0.2 * rng.random())
    K = np.array([[f, 0, width / 2 + rng.uniform(-8, 8)], [0, f * (1 + rng.uniform(-0.002, 0.002)), height / 2 + rng.uniform(-8, 8)], [0, 0, 1]])
    D = strength * np.array([rng.uniform(-0.25, 0.05), rng.uniform(-0.05, 0.15), rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3),
                             rng.uniform(-0.05, 0.05)])
    rv = strength * rng.uniform(-0.01, 0.01, 3)
    th = float(np.linalg.norm(rv))
    kx = rv / th if th > 0 else np.zeros(3)
    Kx = np.array([[0, -kx[2], kx[1]], [kx[2], 0, -kx[0]], [-kx[1], kx[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
    fn = f * (1 + rng.uniform(-0.01, 0.01))
    P = np.array([[fn, 0, width / 2, -fn * 0.3 * (seed & 1)], [0, fn, height / 2, 0], [0, 0, 1, 0]])
    return K, D, R, P

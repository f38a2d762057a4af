"""Seeded synthetic KITTI-shaped inputs: stereo pairs, label masks, poses.

There is no dataset access, so tests and bench.py use these.  numpy only (no cv2, no torch) so the
same bytes are produced in the build container and on the GPU box.
"""
from __future__ import annotations

import numpy as np

from .params import SEGNET12_BGR, CITYSCAPES19_BGR


def _box_blur(a: np.ndarray, r: int, passes: int = 3) -> np.ndarray:
    """Separable box blur applied `passes` times (~Gaussian), edge-replicated."""
    out = a.astype(np.float32)
    k = 2 * r + 1
    for _ in range(passes):
        for axis in (0, 1):
            pad = [(0, 0), (0, 0)]
            pad[axis] = (r + 1, r)
            c = np.cumsum(np.pad(out, pad, mode="edge"), axis=axis, dtype=np.float64)
            if axis == 0:
                out = ((c[k:, :] - c[:-k, :]) / k).astype(np.float32)
            else:
                out = ((c[:, k:] - c[:, :-k]) / k).astype(np.float32)
    return out


def gt_disparity(H: int, W: int, D: int, rng: np.random.Generator) -> np.ndarray:
    """Piecewise-smooth ground truth: road-plane ramp plus fronto-parallel boxes, d in [4, D-8]."""
    dmax = D - 8
    y = np.arange(H, dtype=np.float32)[:, None]
    horizon = 0.42 * H
    ramp = 4.0 + np.clip(y - horizon, 0, None) * ((0.55 * dmax - 4.0) / max(H - horizon, 1.0))
    d = np.broadcast_to(ramp, (H, W)).copy()
    d[: int(horizon)] = 4.0 + 3.0 * rng.random()
    for _ in range(int(rng.integers(4, 9))):
        bw, bh = int(rng.integers(W // 20, W // 5)), int(rng.integers(H // 8, H // 2))
        x0, y0 = int(rng.integers(0, W - bw)), int(rng.integers(0, H - bh))
        val = float(rng.uniform(6.0, dmax))
        region = d[y0 : y0 + bh, x0 : x0 + bw]
        d[y0 : y0 + bh, x0 : x0 + bw] = np.maximum(region, val)
    return np.clip(d, 4.0, dmax)


def stereo_pair(H: int, W: int, D: int, seed: int):
    """Left = band-limited texture; right = left warped by the ground-truth disparity + noise(+-3)."""
    rng = np.random.default_rng(seed)
    tex = _box_blur(rng.random((H, W + D), dtype=np.float32), 2, 2)
    tex = (tex - tex.min()) / max(float(tex.max() - tex.min()), 1e-6)
    wide = np.clip(tex * 255.0, 0, 255)
    left = wide[:, D:]
    d = gt_disparity(H, W, D, rng)
    # right(x) = left(x + d(x)) sampled with linear interpolation in the wide canvas
    xs = np.arange(W, dtype=np.float32)[None, :] + D - 0.0
    # A point seen at x_l in the left image appears at x_r = x_l - d; approximate the inverse warp with d(x_r)
    src = xs + d - D  # left-image column that right pixel x shows
    src = np.clip(src, -D, W - 1.001)
    x0 = np.floor(src).astype(np.int64)
    f = (src - x0).astype(np.float32)
    rows = np.arange(H)[:, None]
    right = wide[rows, x0 + D] * (1 - f) + wide[rows, x0 + D + 1] * f
    right = right + rng.integers(-3, 4, size=(H, W))
    return (
        np.clip(np.rint(left), 0, 255).astype(np.uint8),
        np.clip(np.rint(right), 0, 255).astype(np.uint8),
        d,
    )


def label_mask(H: int, W: int, num_labels: int, seed: int, cell: int = 48):
    """Piecewise-constant class regions from a coarse random grid, NN-upsampled.  Returns (ids u8, BGR u8x3)."""
    rng = np.random.default_rng(seed + 7919)
    gh, gw = -(-H // cell), -(-W // cell)
    grid = rng.integers(0, num_labels, size=(gh, gw), dtype=np.int64)
    ids = np.repeat(np.repeat(grid, cell, axis=0), cell, axis=1)[:H, :W].astype(np.uint8)
    pal = np.asarray(SEGNET12_BGR if num_labels <= 12 else CITYSCAPES19_BGR, np.uint8)
    return ids, pal[ids]


def poses(n: int, seed: int = 0) -> np.ndarray:
    """KITTI-like trajectory: forward 0.8-1.4 m/frame along +z, slow yaw (<= 2 deg/frame).
    Returns [n][4][4] float64 camera->world, row-major (the layout of readGTPose.h:34-78 rows + [0 0 0 1])."""
    rng = np.random.default_rng(seed + 104729)
    T = np.eye(4)
    out = np.empty((n, 4, 4), np.float64)
    yaw_rate = 0.0
    for i in range(n):
        out[i] = T
        yaw_rate = np.clip(yaw_rate + rng.normal(0, 0.15), -2.0, 2.0)
        a = np.deg2rad(yaw_rate)
        step = np.eye(4)
        step[0, 0], step[0, 2], step[2, 0], step[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
        step[2, 3] = rng.uniform(0.8, 1.4)
        step[1, 3] = rng.normal(0, 0.01)
        T = T @ step
    return out


def sequence(n: int, H: int = 376, W: int = 1241, D: int = 128, num_labels: int = 12, seed: int = 0, distinct: int | None = None):
    """n frames: dict(left, right [n][H][W] u8; semantic, rgb [n][H][W][3] u8; label [n][H][W] u8; pose [n][4][4] f64).
    `distinct` < n re-uses images cyclically (poses stay distinct) to bound host generation time."""
    distinct = n if distinct is None else min(distinct, n)
    L = np.empty((distinct, H, W), np.uint8)
    R = np.empty((distinct, H, W), np.uint8)
    S = np.empty((distinct, H, W, 3), np.uint8)
    Lab = np.empty((distinct, H, W), np.uint8)
    for i in range(distinct):
        L[i], R[i], _ = stereo_pair(H, W, D, seed * 100003 + i)
        Lab[i], S[i] = label_mask(H, W, num_labels, seed * 100003 + i)
    idx = np.arange(n) % distinct
    rgb = np.repeat(L[..., None], 3, axis=-1)
    return {
        "left": L[idx], "right": R[idx], "semantic": S[idx], "rgb": rgb[idx], "label": Lab[idx],
        "pose": poses(n, seed),
    }

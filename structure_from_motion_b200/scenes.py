"""Synthetic two-view scenes for the bench and the parity tests.

Random 3-D points in front of two known cameras, Gaussian pixel noise, and a stated
fraction of outliers (image-2 coordinates replaced by uniform pixels).  numpy only.
"""
from __future__ import annotations

import numpy as np


def euler_xyz_intrinsic(deg_x: float, deg_y: float, deg_z: float) -> np.ndarray:
    """Rotation matrix of intrinsic X-Y-Z Euler angles (degrees): R = Rx @ Ry @ Rz."""
    ax, ay, az = np.radians([deg_x, deg_y, deg_z])
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cx, -sx], [0.0, sx, cx]])
    ry = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    rz = np.array([[cz, -sz, 0.0], [sz, cz, 0.0], [0.0, 0.0, 1.0]])
    return rx @ ry @ rz


def make_scene(n: int, outlier_frac: float, seed: int, noise_px: float = 0.5,
               f: float = 800.0, w: int = 1280, h: int = 960):
    """Return (K, x1[n,2], x2[n,2], R, t, outlier_idx) in pixel coordinates.

    cam1 = [I|0]; cam2 = [R|t] with R = euler XYZ (2, -8, 1) deg, t = (-1, 0.05, 0.1).
    """
    rng = np.random.default_rng(seed)
    K = np.array([[f, 0.0, w / 2], [0.0, f, h / 2], [0.0, 0.0, 1.0]])
    X = np.column_stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(4, 8, n)])
    R = euler_xyz_intrinsic(2.0, -8.0, 1.0)
    t = np.array([-1.0, 0.05, 0.1])
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:]
    X2 = (R @ X.T).T + t
    x2 = (K @ X2.T).T
    x2 = x2[:, :2] / x2[:, 2:]
    x1 += rng.normal(0, noise_px, x1.shape)
    x2 += rng.normal(0, noise_px, x2.shape)
    n_out = int(round(outlier_frac * n))
    idx = rng.choice(n, n_out, replace=False)
    x2[idx] = np.column_stack([rng.uniform(0, w, n_out), rng.uniform(0, h, n_out)])
    return K, np.ascontiguousarray(x1), np.ascontiguousarray(x2), R, t, idx

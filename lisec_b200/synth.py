"""Synthetic Lyft-shaped lidar sweeps. There is no dataset here: the reference's combine_lidar_data
(model_training.py:73-98) reads three lidar .bin files per sample and returns one (n,3) array in the ego frame;
these generators return arrays of that contract, float32-valued (the .bin files are float32)."""
from __future__ import annotations

import numpy as np


def adversarial_tail() -> np.ndarray:
    """Points on and around every edge of the key and range tests (model_training.py:103-107, 118-120)."""
    eps32 = np.float32(1.1920929e-07)
    xs = [-50.0, -49.5, -49.75, -0.5, -0.25, -0.0, 0.0, 0.25, 0.5, 49.5, 49.75, 50.0, 1e9, -1e9,
          float(np.nextafter(np.float32(50.0), np.float32(0.0))), float(np.nextafter(np.float32(-49.5), np.float32(-60.0)))]
    zs = [-0.25, -0.0, 0.0, 0.25, float(np.nextafter(np.float32(0.25), np.float32(0.0))), 0.5, 1.75,
          float(np.nextafter(np.float32(2.0), np.float32(0.0))), 2.0, 2.25]
    pts = []
    for x in xs:
        for y in xs:
            for z in (0.25, 1.0, zs[7]):
                pts.append((x, y, z))
    for z in zs:
        for x in (-49.5, 0.0, 12.5, 49.75):
            pts.append((x, x / 2, z))
    # exact multiples of the voxel sizes, and their float32 neighbours
    for k in range(-8, 9):
        v = np.float32(k * 0.25)
        for d in (-eps32 * 8, 0.0, eps32 * 8):
            pts.append((float(v + np.float32(d)), float(v / 2 + np.float32(d)), 1.0 + float(np.float32(d))))
    pts = np.asarray(pts, dtype=np.float32)
    # duplicates: the same point many times (more than T = 35 in one voxel, exactly 35 in another)
    dup_over = np.tile(np.asarray([[3.3, 4.4, 1.1]], dtype=np.float32), (50, 1))
    dup_exact = np.tile(np.asarray([[-7.3, 2.6, 0.6]], dtype=np.float32), (35, 1))
    near = np.asarray([[3.3, 4.4, 1.1]], dtype=np.float32) + np.linspace(0, 0.05, 40, dtype=np.float32)[:, None]
    return np.concatenate([pts, dup_over, dup_exact, near]).astype(np.float32)


def lyft_like_sweep(n_points: int = 100_000, seed: int = 0, theta: float = 9.0, tail: bool = True) -> np.ndarray:
    """One sweep: range ~ Gamma(k=2, theta m), azimuth uniform, height ~ N(0.6, 0.7), three sensor origins.
    About 65 % of the points pass the range test and they occupy about 50 k voxels at n_points = 100 k."""
    rng = np.random.default_rng(seed)
    adv = adversarial_tail() if tail else np.zeros((0, 3), np.float32)
    n = max(n_points - len(adv), 0)
    r = rng.gamma(2.0, theta, size=n)
    az = rng.uniform(0.0, 2 * np.pi, size=n)
    z = rng.normal(0.6, 0.7, size=n)
    sensor = rng.integers(0, 3, size=n)
    origin = np.asarray([[1.2, 0.0], [2.0, -0.6], [2.0, 0.6]])[sensor]
    xy = origin + np.stack([r * np.cos(az), r * np.sin(az)], axis=1)
    pts = np.concatenate([xy, z[:, None]], axis=1).astype(np.float32)
    pts = np.concatenate([pts, adv])[:n_points]
    perm = rng.permutation(len(pts))  # the tail is interleaved, not appended
    return np.ascontiguousarray(pts[perm])


def sweep_batch(n_sweeps: int, n_points: int = 100_000, seed0: int = 0, theta: float = 9.0):
    """Concatenated sweeps (seeds seed0..) and their offsets: the (points, sweep_offsets) pair of the C ABI."""
    sweeps = [lyft_like_sweep(n_points, seed0 + s, theta) for s in range(n_sweeps)]
    offsets = np.zeros(n_sweeps + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(s) for s in sweeps])
    return np.ascontiguousarray(np.concatenate(sweeps)), offsets


def saturated_cloud(n_points: int = 1_000_000, n_sweeps: int = 10, seed0: int = 0, theta: float = 3.5) -> np.ndarray:
    """BASELINE config 4: several sweeps merged into ONE cloud with a tight range law, so that a large share of the
    near-field voxels exceed T = 35 points."""
    per = n_points // n_sweeps
    return np.ascontiguousarray(
        np.concatenate([lyft_like_sweep(per, seed0 + s, theta, tail=(s == 0)) for s in range(n_sweeps)])
    )

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


class SyntheticLyftTables:
    """The two tables of lyft_dataset_sdk.LyftDataset that combine_lidar_data reads through `.get(table, token)`
    (model_training.py:81-84)."""

    def __init__(self):
        self.tables = {"sample_data": {}, "calibrated_sensor": {}}

    def get(self, table: str, token: str) -> dict:
        return self.tables[table][token]


def synthetic_lyft_sample(data_dir: str, n_points: int = 3000, seed: int = 0, sensors=("LIDAR_TOP", "LIDAR_FRONT_RIGHT",
                                                                                      "LIDAR_FRONT_LEFT")):
    """Writes one sample's sensor files the way the Lyft dataset lays them out — `lidar/<token>.bin`, float32 records
    (x, y, z, intensity, ring) in the SENSOR frame — and returns (sample, tables) for combine_lidar_data. The sensor
    poses are Lyft-like: the roof lidar nearly level, the two bumper lidars yawed outwards and slightly rolled."""
    import os

    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(data_dir, "lidar"), exist_ok=True)
    tables = SyntheticLyftTables()
    poses = {
        "LIDAR_TOP": ([0.99995, 0.0021, -0.0047, 0.0083], [1.2018, 0.0034, 1.8312]),
        "LIDAR_FRONT_RIGHT": ([0.9239, 0.0105, -0.0052, -0.3824], [2.0421, -0.6103, 0.6251]),
        "LIDAR_FRONT_LEFT": ([0.9236, -0.0098, -0.0049, 0.3831], [2.0397, 0.6124, 0.6247]),
    }
    sample = {"data": {}}
    per = [n_points // len(sensors)] * len(sensors)
    per[0] += n_points - sum(per)
    for k, (name, n) in enumerate(zip(sensors, per)):
        r = rng.gamma(2.0, 9.0, size=n)
        az = rng.uniform(0, 2 * np.pi, size=n)
        rec = np.stack([r * np.cos(az), r * np.sin(az), rng.normal(-0.9, 0.7, size=n), rng.uniform(0, 255, size=n),
                        rng.integers(0, 64, size=n).astype(np.float64)], axis=1).astype(np.float32)
        token = "%s_%d" % (name.lower(), seed)
        rec.tofile(os.path.join(data_dir, "lidar", token + ".bin"))
        quat, trans = poses[name]
        quat = (np.asarray(quat) + rng.normal(0, 1e-3, size=4)).tolist()  # not exactly unit: _normalise has work to do
        tables.tables["sample_data"][token] = {"filename": "lidar/%s.bin" % token, "calibrated_sensor_token": "cs_" + token}
        tables.tables["calibrated_sensor"]["cs_" + token] = {"rotation": quat, "translation": list(trans)}
        sample["data"][name] = token
    return sample, tables


def synthetic_rpn_output(seed: int = 0, out_x: int = 100, out_y: int = 200, n_anchors: int = 2, n_objects: int = 40):
    """(labelsClass (out_x,out_y,n_anchors), labelsRegress (out_x,out_y,7*n_anchors)) float32, shaped like what
    model.predict returns for one sample (Predict.py:38): a low, noisy score floor with `n_objects` bumps, small
    regressions. Scores are distinct (the reference's argsort leaves tie order unspecified)."""
    rng = np.random.default_rng(seed)
    cls = rng.uniform(0.0, 0.3, size=(out_x, out_y, n_anchors))
    for _ in range(n_objects):
        a, b, i = rng.integers(0, out_x), rng.integers(0, out_y), rng.integers(0, n_anchors)
        aa, bb = np.meshgrid(np.arange(out_x), np.arange(out_y), indexing="ij")
        cls[:, :, i] += rng.uniform(0.4, 0.7) * np.exp(-((aa - a) ** 2 / 8.0 + (bb - b) ** 2 / 30.0))
    cls = cls.astype(np.float32)
    flat = cls.reshape(-1)
    order = np.argsort(flat, kind="stable")
    flat[order] = np.sort(flat) + np.arange(flat.size, dtype=np.float32) * np.float32(1e-7)  # break ties, keep the order
    assert len(np.unique(flat)) == flat.size
    reg = rng.normal(0.0, 0.25, size=(out_x, out_y, 7 * n_anchors))
    reg[:, :, 6::7] = rng.normal(0.0, 0.6, size=(out_x, out_y, n_anchors))
    return cls, reg.astype(np.float32)

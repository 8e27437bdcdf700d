"""CPU oracle for the lidar ingest step.  *** TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT ***

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Restated, with the reference lines each function follows (bot15498/Lisec):
  rotate_points(points, rotation, inverse)        model_training.py:65-69
  combine_lidar_data(sample, dataDir, level5Data) model_training.py:73-98
  Quaternion                                      pyquaternion (a dependency absent from /root/reference and from this
                                                  image; the reference pins no version — `from pyquaternion import
                                                  Quaternion`, model_training.py:17). Restated from the published
                                                  algorithm of pyquaternion 0.9.x: Quaternion(array) stores float64
                                                  (w,x,y,z); .inverse = conjugate / sum of squares; .rotation_matrix
                                                  normalises unless |1 - q.q| < 1e-14, then returns the lower-right 3x3
                                                  of Q(q) . Qbar(q)^T.
  rotate_points_fma_chain                         the same product spelled out as the rounding sequence numpy's BLAS
                                                  executes for it: acc = R[i,0]*x; acc = fma(R[i,1], y, acc);
                                                  acc = fma(R[i,2], z, acc), fma emulated EXACTLY with fractions.Fraction
                                                  (slow: a few thousand points). tests/test_oracle.py checks it equals
                                                  np.dot on this host bit for bit.

literal_*: the reference's OWN lines model_training.py:65-98, read from /root/reference at run time and executed
unmodified in a namespace that supplies np, os and the Quaternion restatement above (build container only; used by
tests/golden/make_golden_ingest.py to mint tests/golden/ingest.npz).

Pinning status: the arithmetic of :65-69 and :93-96 is PINNED by that literal run (numpy is the only thing it touches
once the matrix exists); the quaternion -> matrix step is a restatement of a third-party library that is not here
(PARITY UNPINNED for that step alone; its result is checked against the textbook closed form to 1e-15).
"""
from __future__ import annotations

import os
from fractions import Fraction

import numpy as np

REFERENCE = "/root/reference/model_training.py"
SENSOR_TYPES = ["LIDAR_TOP", "LIDAR_FRONT_RIGHT", "LIDAR_FRONT_LEFT"]


class Quaternion:
    """The slice of pyquaternion.Quaternion the reference touches (model_training.py:66-69)."""

    def __init__(self, rotation):
        if isinstance(rotation, Quaternion):
            self.q = rotation.q.copy()
        else:
            self.q = np.asarray(rotation, dtype=float).reshape(4).copy()

    def _sum_of_squares(self):
        return np.dot(self.q, self.q)

    @property
    def conjugate(self):
        return Quaternion(np.array([self.q[0], -self.q[1], -self.q[2], -self.q[3]]))

    @property
    def inverse(self):
        ss = self._sum_of_squares()
        if ss > 0:
            return Quaternion(self.conjugate.q / ss)
        raise ZeroDivisionError("a zero quaternion (0 + 0i + 0j + 0k) cannot be inverted")

    def _normalise(self):
        if not abs(1.0 - self._sum_of_squares()) < 1e-14:
            n = np.sqrt(self._sum_of_squares())
            if n > 0:
                self.q = self.q / n

    def _q_matrix(self):
        q = self.q
        return np.array([[q[0], -q[1], -q[2], -q[3]], [q[1], q[0], -q[3], q[2]], [q[2], q[3], q[0], -q[1]],
                         [q[3], -q[2], q[1], q[0]]])

    def _q_bar_matrix(self):
        q = self.q
        return np.array([[q[0], -q[1], -q[2], -q[3]], [q[1], q[0], q[3], -q[2]], [q[2], -q[3], q[0], q[1]],
                         [q[3], q[2], -q[1], q[0]]])

    @property
    def rotation_matrix(self):
        self._normalise()
        product_matrix = np.dot(self._q_matrix(), self._q_bar_matrix().conj().transpose())
        return product_matrix[1:][:, 1:]


# model_training.py:65-69
def rotate_points(points, rotation, inverse=False):
    quaternion = Quaternion(rotation)
    if inverse:
        quaternion = quaternion.inverse
    return np.dot(quaternion.rotation_matrix, points.T).T


# model_training.py:73-98 (the Windows path rewrite of :86 is kept as the first spelling tried)
def combine_lidar_data(sample, dataDir, level5Data):
    actual = [s for s in SENSOR_TYPES if s in sample["data"]]
    frames = [level5Data.get("sample_data", sample["data"][x]) for x in actual]
    all_points = []
    for frame in frames:
        sensor = level5Data.get("calibrated_sensor", frame["calibrated_sensor_token"])
        path = os.path.join(dataDir, frame["filename"].replace("/", "\\"))
        if not os.path.exists(path):
            path = os.path.join(dataDir, frame["filename"])
        raw = np.fromfile(path, dtype=np.float32).reshape(-1, 5)[:, :3]
        points = rotate_points(raw, sensor["rotation"])
        points = points + np.array(sensor["translation"])
        all_points.append(points)
    return np.concatenate(all_points)


def transform_segments(records, segment_offsets, rotations, translations):
    """The batched form the GPU entry point takes: records float32 [n,5]; per segment a quaternion and a translation."""
    out = np.empty((len(records), 3), dtype=np.float64)
    for s in range(len(segment_offsets) - 1):
        a, b = int(segment_offsets[s]), int(segment_offsets[s + 1])
        out[a:b] = rotate_points(records[a:b, :3], rotations[s]) + np.array(translations[s])
    return out


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def rotate_points_fma_chain(points, matrix, translation=None):
    """Exact restatement of the rounding sequence (see the module docstring); non-finite inputs are not supported."""
    pts = np.asarray(points)
    out = np.empty((len(pts), 3), dtype=np.float64)
    m = np.asarray(matrix, dtype=np.float64)
    for j in range(len(pts)):
        x, y, z = float(pts[j, 0]), float(pts[j, 1]), float(pts[j, 2])
        for i in range(3):
            acc = float(m[i, 0]) * x
            acc = _fma(float(m[i, 1]), y, acc)
            acc = _fma(float(m[i, 2]), z, acc)
            out[j, i] = acc if translation is None else acc + float(translation[i])
    return out


# ---- literal mode: the reference's own lines, build container only -------------------------------------------------
def literal_available() -> bool:
    return os.path.exists(REFERENCE)


def literal_functions(path_rewrite: bool = True):
    """(rotate_points, combine_lidar_data) compiled from model_training.py:65-98 as they stand. On Linux the '\\\\'
    rewrite of :86 names a file that does not exist, so `os` is a shim whose path.join maps '\\\\' back to '/' — a knob
    on the function's dependency, not on its source."""
    import types

    with open(REFERENCE) as f:
        lines = f.readlines()
    src = "".join(lines[64:98])
    os_shim = types.ModuleType("os_shim")
    path_shim = types.ModuleType("os_shim.path")
    path_shim.join = (lambda *a: os.path.join(*[x.replace("\\", "/") for x in a])) if path_rewrite else os.path.join
    os_shim.path = path_shim
    ns = {"np": np, "os": os_shim, "Quaternion": Quaternion}
    exec(compile(src, REFERENCE + ":65-98", "exec"), ns)
    return ns["rotate_points"], ns["combine_lidar_data"]

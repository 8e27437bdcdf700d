"""Lidar ingest on the GPU — the step before the path.

    reference (model_training.py)                                   here
    --------------------------------------------------------------- ------------------------------------------------
    rotate_points(points, rotation, inverse=False)        :65-69    rotate_points(...)           -> numpy float64 (n,3)
    combine_lidar_data(sample, dataDir, level5Data)       :73-98    combine_lidar_data(...)      -> numpy float64 (n,3)
                                                                    combine_lidar_data_device(...) -> cuda float64 (n,3)
                                                                    LidarIngest.transform(...)   (batched, stays on the GPU)

The arithmetic (float32 records -> float64 rotate + translate) runs in `ingest_kernel` behind lisec_ingest_lidar
(lisec_b200/csrc/ingest.cu); the host only reads the files and turns each sensor's quaternion into a 3x3 matrix the way
pyquaternion does (the reference's dependency, version unpinned; algorithm of pyquaternion 0.9.x `Quaternion.rotation_matrix`:
normalise unless |1 - q.q| < 1e-14, then the lower-right 3x3 block of Q(q) . Qbar(q)^T). There is no CPU fallback for the
point arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N

SENSOR_TYPES = ("LIDAR_TOP", "LIDAR_FRONT_RIGHT", "LIDAR_FRONT_LEFT")  # model_training.py:74
RECORD_FLOATS = 5  # x, y, z, intensity, ring index (model_training.py:90)


def quaternion_rotation_matrix(rotation: Sequence[float], inverse: bool = False) -> np.ndarray:
    """Quaternion(rotation)[.inverse].rotation_matrix as pyquaternion computes it (float64, (w, x, y, z) order)."""
    q = np.asarray(rotation, dtype=np.float64).reshape(4).copy()
    if inverse:  # Quaternion.inverse: conjugate / sum of squares
        ss = np.dot(q, q)
        if ss <= 0:
            raise ZeroDivisionError("a zero quaternion (0 + 0i + 0j + 0k) cannot be inverted")
        q = np.array([q[0], -q[1], -q[2], -q[3]]) / ss
    ss = np.dot(q, q)
    if not abs(1.0 - ss) < 1e-14:  # Quaternion._normalise
        n = np.sqrt(ss)
        if n > 0:
            q = q / n
    w, x, y, z = q
    qm = np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])
    qb = np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])
    return np.dot(qm, qb.conj().transpose())[1:][:, 1:]


class LidarIngest:
    """Batched front of the front end: raw sensor records -> float64 (n,3) points on the GPU, one kernel per 24 files."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("lisec_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = N.load()
        self.device = torch.device("cuda", device)
        self.last_launch_count = 0

    @staticmethod
    def make_poses(rotations: Sequence, translations: Sequence):
        """The lisec_sensor_pose array of the C ABI: per segment a quaternion (4,) or a 3x3 matrix, and a translation.
        Sensor calibrations do not change between sweeps of a scene: build once, pass as `poses=` afterwards."""
        nseg = len(rotations)
        if len(translations) != nseg:
            raise ValueError("rotations / translations do not describe the same segments")
        poses = (N.lisec_sensor_pose * max(nseg, 1))()
        for s in range(nseg):
            r = np.asarray(rotations[s], dtype=np.float64)
            m = quaternion_rotation_matrix(r) if r.size == 4 else r.reshape(3, 3)
            poses[s].rotation[:] = [float(v) for v in m.reshape(9)]
            poses[s].translation[:] = [float(v) for v in np.asarray(translations[s], dtype=np.float64).reshape(3)]
        poses._n = nseg
        return poses

    def transform(self, records, segment_offsets: Sequence[int], rotations: Sequence = None,
                  translations: Sequence = None, out: Optional[torch.Tensor] = None,
                  record_floats: int = RECORD_FLOATS, poses=None) -> torch.Tensor:
        """records: float32 [n, record_floats] (numpy, or a torch tensor on the host or already on the device);
        segment_offsets: n_segments + 1 point offsets; rotations / translations as in make_poses(), or a ready `poses`.
        Returns float64 [n,3] on the device."""
        off = segment_offsets if (isinstance(segment_offsets, np.ndarray) and segment_offsets.dtype == np.int64 and
                                  segment_offsets.flags.c_contiguous) else \
            np.ascontiguousarray(np.asarray(segment_offsets, dtype=np.int64))
        nseg = len(off) - 1
        if poses is None:
            poses = self.make_poses(rotations, translations)
        if off.ndim != 1 or nseg < 0 or poses._n != nseg:
            raise ValueError("segment_offsets / rotations / translations do not describe the same segments")
        if isinstance(records, np.ndarray):
            if records.dtype != np.float32:
                raise ValueError("records must be float32 (np.fromfile(..., dtype=np.float32), model_training.py:87)")
            records = torch.from_numpy(np.ascontiguousarray(records))
        rec = records.to(self.device, non_blocking=True).contiguous().view(-1)
        if rec.dtype != torch.float32:
            raise ValueError("records must be float32")
        n = int(off[-1]) if nseg >= 0 and len(off) else 0
        if rec.numel() != n * record_floats:
            raise ValueError("%d record floats but segment_offsets[-1] * %d = %d" % (rec.numel(), record_floats,
                                                                                    n * record_floats))
        if out is None:
            out = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        elif out.dtype != torch.float64 or out.device != self.device or out.numel() < 3 * n or not out.is_contiguous():
            raise ValueError("out must be a contiguous cuda float64 tensor of at least n*3 elements")
        launches = C.c_int32(0)
        with torch.cuda.device(self.device):
            st = self._lib.lisec_ingest_lidar(
                C.c_void_p(rec.data_ptr() if n else 0), record_floats, off.ctypes.data_as(C.POINTER(C.c_int64)), poses,
                nseg, C.c_void_p(out.data_ptr() if n else 0),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream), C.byref(launches))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_ingest_last_error().decode("utf-8", "replace"))
        self.last_launch_count = launches.value
        self._keep = rec
        return out[:n] if out.shape[0] != n else out


_INGEST: dict = {}


def _ingest(device: int = 0) -> LidarIngest:
    if device not in _INGEST:
        _INGEST[device] = LidarIngest(device)
    return _INGEST[device]


def rotate_points(points, rotation, inverse: bool = False) -> np.ndarray:
    """Drop-in for rotate_points (model_training.py:65-69): np.dot(rotation_matrix, points.T).T, float64 (n,3).
    float32 points take the kernel's exact float32 -> float64 widening; float64 points are not the ingest path
    (the reference only ever passes the float32 records) and are rejected rather than silently rounded."""
    p = np.ascontiguousarray(points)
    if p.dtype != np.float32:
        raise TypeError("rotate_points on the GPU takes the float32 records of the sensor files (model_training.py:87-93)")
    m = quaternion_rotation_matrix(rotation, inverse)
    # translation -0.0: x + (-0.0) == x bit for bit, signed zeros included
    out = _ingest().transform(p, [0, len(p)], [m], [np.full(3, -0.0)], record_floats=p.shape[1])
    return out.cpu().numpy()


def _sensor_files(sample, dataDir, level5Data) -> List[Tuple[str, Sequence[float], Sequence[float]]]:
    """(path, rotation quaternion, translation) per present sensor, in the reference's order (model_training.py:74-84)."""
    res = []
    for sensor_type in SENSOR_TYPES:
        if sensor_type not in sample["data"]:
            continue  # "not all samples having all lidar data" (:76-79)
        frame = level5Data.get("sample_data", sample["data"][sensor_type])
        sensor = level5Data.get("calibrated_sensor", frame["calibrated_sensor_token"])
        # the reference was written on Windows and rewrites '/' to '\\' (:86); take that spelling when it exists
        literal = os.path.join(dataDir, frame["filename"].replace("/", "\\"))
        path = literal if os.path.exists(literal) else os.path.join(dataDir, frame["filename"])
        res.append((path, sensor["rotation"], sensor["translation"]))
    return res


def combine_lidar_data_device(sample, dataDir, level5Data, device: int = 0) -> torch.Tensor:
    """combine_lidar_data with the result left on the GPU (float64 (n,3)), ready for Frontend.forward()."""
    files = _sensor_files(sample, dataDir, level5Data)
    raws = [np.fromfile(path, dtype=np.float32).reshape(-1, RECORD_FLOATS) for path, _, _ in files]
    off = np.cumsum([0] + [len(r) for r in raws])
    rec = np.concatenate(raws) if raws else np.zeros((0, RECORD_FLOATS), dtype=np.float32)
    return _ingest(device).transform(rec, off, [f[1] for f in files], [f[2] for f in files])


def combine_lidar_data(sample, dataDir, level5Data) -> np.ndarray:
    """Drop-in for combine_lidar_data (model_training.py:73-98): numpy float64 (n,3), all sensors concatenated."""
    return combine_lidar_data_device(sample, dataDir, level5Data).cpu().numpy()

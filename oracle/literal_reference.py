"""Runs the reference's OWN source for the voxelizer, unmodified.  *** TEST INFRASTRUCTURE — build container only ***

model_training.py cannot be imported here (matplotlib, tensorflow, lyft_dataset_sdk, pyquaternion and shapely are
absent), but lines 103-152 — get_voxel and VFE_preprocessing — need only numpy, math.floor and the SparseTensor
*constructor*. This module reads exactly those lines from /root/reference at run time (nothing is copied into the
repo), compiles them, and executes them in a namespace that supplies `np`, `floor` and a three-field SparseTensor.

Two knobs, both on the function's DEPENDENCIES, never on its source:
  seed          np.random.seed(seed) before the call freezes np.random.choice at :132 (the reference is unseeded).
  sampler       "numpy_rng": the real np.random.choice.  "first_T": `np` is a shim whose random.choice returns the
                first `size` entries of the list it was given — the product's deterministic sampling contract.
The shim also records every list handed to np.random.choice, i.e. clusteredPoints in dict order.

/root/reference does not exist on the GPU box: only tests/golden/make_golden.py and the container-only tests use this.
"""
from __future__ import annotations

import os
import types
from math import floor

import numpy as np

REFERENCE = "/root/reference/model_training.py"
FIRST_LINE, LAST_LINE = 103, 152


def available() -> bool:
    return os.path.exists(REFERENCE)


class _SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = indices, values, dense_shape
        self.shape = tuple(dense_shape)


def _namespace(sampler: str, log: list):
    real_choice = np.random.choice

    def choice(a, size=None, replace=True, p=None):
        log.append(list(a))
        if sampler == "first_T":
            return np.asarray(list(a)[:size], dtype=np.int64)
        return real_choice(a, size=size, replace=replace, p=p)

    shim = types.ModuleType("np_shim")
    for k in dir(np):
        if not k.startswith("__"):
            try:
                setattr(shim, k, getattr(np, k))
            except Exception:
                pass
    rnd = types.ModuleType("np_shim.random")
    for k in dir(np.random):
        if not k.startswith("__"):
            setattr(rnd, k, getattr(np.random, k))
    rnd.choice = choice
    shim.random = rnd
    return {"np": shim, "floor": floor, "SparseTensor": _SparseTensor}


def load_functions(sampler: str = "numpy_rng"):
    with open(REFERENCE) as f:
        lines = f.readlines()
    src = "".join(lines[FIRST_LINE - 1:LAST_LINE])
    log: list = []
    ns = _namespace(sampler, log)
    exec(compile(src, REFERENCE + ":%d-%d" % (FIRST_LINE, LAST_LINE), "exec"), ns)
    return ns["get_voxel"], ns["VFE_preprocessing"], log


def run(points, xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8,
        sampler="numpy_rng", seed=None):
    """Returns (SparseTensor-like, clusteredPoints lists in dict order). points are passed as float64, the dtype
    combine_lidar_data produces (model_training.py:93-94)."""
    _, vfe_pre, log = load_functions(sampler)
    if seed is not None:
        np.random.seed(seed)
    st = vfe_pre(np.asarray(points, dtype=np.float64), xSize, ySize, zSize, sampleSize, maxVoxelX, maxVoxelY,
                 maxVoxelZ)
    return st, log

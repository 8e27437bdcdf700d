"""Mint the golden vectors of tests/golden/ by running the reference's own source (model_training.py:103-152,
unmodified, via oracle/literal_reference.py). Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Outputs (committed):
  tiny_rng.npz        small cloud, real np.random.choice with np.random.seed(3): full COO indices + values
  tiny_first.npz      same cloud, first_T sampler: full COO indices + values + clusteredPoints
  adversarial.npz     the edge-case cloud of lisec_b200.synth.adversarial_tail(): first_T COO + clusteredPoints
  sweep100k.json      config-1 sweep (100 k points, seed 0): sizes and sha256 digests of the first_T outputs
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from lisec_b200 import synth  # noqa: E402
from oracle import literal_reference as lit  # noqa: E402

ARGS = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)


def ragged(lists):
    flat = np.asarray([i for l in lists for i in l], dtype=np.int32)
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(l) for l in lists])
    return flat, off


def tiny_cloud():
    rng = np.random.default_rng(11)
    a = rng.uniform([-3, -2, 0.3], [3, 2, 1.9], size=(260, 3))
    b = np.asarray([[0.6, 0.3, 0.8]]) + rng.uniform(0, 0.2, size=(60, 3)) * [1, 0.5, 0.5]  # one voxel region, > T points
    c = rng.uniform(-80, 80, size=(40, 3))  # mostly out of range
    return np.concatenate([a, b, c]).astype(np.float32)[rng.permutation(360)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert lit.available(), "needs /root/reference"
    pts = tiny_cloud()
    st, groups = lit.run(pts, sampler="numpy_rng", seed=3, **ARGS)
    flat, off = ragged(groups)
    np.savez_compressed(os.path.join(HERE, "tiny_rng.npz"), points=pts, seed=3,
                        indices=np.asarray(st.indices, dtype=np.int16), values=np.asarray(st.values, dtype=np.float64),
                        dense_shape=np.asarray(st.dense_shape), groups_flat=flat, groups_off=off)
    st, groups = lit.run(pts, sampler="first_T", **ARGS)
    flat, off = ragged(groups)
    np.savez_compressed(os.path.join(HERE, "tiny_first.npz"), points=pts,
                        indices=np.asarray(st.indices, dtype=np.int16), values=np.asarray(st.values, dtype=np.float64),
                        dense_shape=np.asarray(st.dense_shape), groups_flat=flat, groups_off=off)
    print("tiny: %d voxels, %d nnz" % (len(groups), len(st.values)))

    adv = synth.adversarial_tail()
    adv = adv[np.isfinite(adv).all(1)]
    st, groups = lit.run(adv, sampler="first_T", **ARGS)
    flat, off = ragged(groups)
    np.savez_compressed(os.path.join(HERE, "adversarial.npz"), points=adv,
                        indices=np.asarray(st.indices, dtype=np.int16), values=np.asarray(st.values, dtype=np.float64),
                        dense_shape=np.asarray(st.dense_shape), groups_flat=flat, groups_off=off)
    print("adversarial: %d points, %d voxels, %d nnz" % (len(adv), len(groups), len(st.values)))

    sweep = synth.lyft_like_sweep(100_000, seed=0)
    t0 = time.time()
    st, groups = lit.run(sweep, sampler="first_T", **ARGS)
    dt = time.time() - t0
    ind = np.asarray(st.indices, dtype=np.int32)
    val = np.asarray(st.values, dtype=np.float64)
    flat, off = ragged(groups)
    meta = {
        "generator": "lisec_b200.synth.lyft_like_sweep(100000, seed=0)",
        "points_sha256": sha(sweep),
        "n_points": int(len(sweep)),
        "n_voxels": int(len(groups)),
        "nnz": int(len(val)),
        "n_in_range": int(len(flat)),
        "indices_int32_sha256": sha(ind),
        "values_float64_sha256": sha(val),
        "values_float32_sha256": sha(val.astype(np.float32)),
        "groups_flat_int32_sha256": sha(flat),
        "groups_off_int64_sha256": sha(off),
        "literal_seconds_1core": round(dt, 2),
    }
    with open(os.path.join(HERE, "sweep100k.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()

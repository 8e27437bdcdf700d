"""VFE weight packs keyed by Keras layer names, in the creation order of the reference's createModel
(model_training.py:229-235). The shipped blob SampleModel/15SampleEpoch0.h5 is absent from the
reference checkout and h5py is not in this image, so packs travel as .npz with the same keys an .h5 would carry:

    dense/kernel (6,16)                batch_normalization/{gamma,beta,moving_mean,moving_variance} (16,)
    dense_1/kernel (32,32)             batch_normalization_1/{...} (32,)
    dense_2/kernel (64,64)             batch_normalization_2/{...} (64,)
"""
from __future__ import annotations

import numpy as np

VFE_DENSE = ("dense", "dense_1", "dense_2")
VFE_BN = ("batch_normalization", "batch_normalization_1", "batch_normalization_2")
VFE_SHAPES = ((6, 16), (32, 32), (64, 64))
BN_FIELDS = ("gamma", "beta", "moving_mean", "moving_variance")


def vfe_keys():
    keys = []
    for d, b in zip(VFE_DENSE, VFE_BN):
        keys.append(d + "/kernel")
        keys.extend(b + "/" + f for f in BN_FIELDS)
    return keys


def synthetic_vfe_pack(seed: int = 0) -> dict:
    """Seeded stand-in for the missing .h5: Glorot-uniform kernels (the Keras Dense default) and non-trivial BN
    statistics, so that c_empty != 0 and the pad rows really do take part in the max-pools."""
    rng = np.random.default_rng(seed)
    pack = {}
    for d, b, (cin, cout) in zip(VFE_DENSE, VFE_BN, VFE_SHAPES):
        lim = np.sqrt(6.0 / (cin + cout))
        pack[d + "/kernel"] = rng.uniform(-lim, lim, size=(cin, cout)).astype(np.float32)
        pack[b + "/gamma"] = rng.uniform(0.5, 1.5, size=cout).astype(np.float32)
        pack[b + "/beta"] = rng.uniform(-0.3, 0.3, size=cout).astype(np.float32)
        pack[b + "/moving_mean"] = rng.uniform(-0.5, 0.5, size=cout).astype(np.float32)
        pack[b + "/moving_variance"] = rng.uniform(0.3, 2.0, size=cout).astype(np.float32)
    return pack


def validate_vfe_pack(pack: dict) -> dict:
    out = {}
    for d, b, shape in zip(VFE_DENSE, VFE_BN, VFE_SHAPES):
        k = np.ascontiguousarray(pack[d + "/kernel"], dtype=np.float32)
        if k.shape != shape:
            raise ValueError("%s/kernel has shape %s, expected %s" % (d, k.shape, shape))
        out[d + "/kernel"] = k
        for f in BN_FIELDS:
            v = np.ascontiguousarray(pack[b + "/" + f], dtype=np.float32)
            if v.shape != (shape[1],):
                raise ValueError("%s/%s has shape %s, expected (%d,)" % (b, f, v.shape, shape[1]))
            out[b + "/" + f] = v
    return out


def save_npz(path: str, pack: dict) -> None:
    np.savez(path, **{k.replace("/", "."): v for k, v in pack.items()})


def load_npz(path: str) -> dict:
    with np.load(path) as z:
        return {k.replace(".", "/"): z[k] for k in z.files}

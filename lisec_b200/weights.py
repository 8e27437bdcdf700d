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


# ---- the dense network behind the voxel grid (model_training.py:236-256), Keras names in creation order ----------
#   conv3d[_k] kernel (3,3,3,64,64) + bias (64) | batch_normalization_{3,4,5} | dense_{3,4,5} kernel (64,64)
#   conv2d[_k] kernel (3,3,Cin,Cout) + bias     | batch_normalization_{6..21}
#   conv2d_transpose[_k] kernel (k,k,Cout,Cin) + bias (Keras keeps transposed kernels output-channel-first)
#   ClassificationLayer kernel (1,1,768,2) + bias, RegressionLayer kernel (1,1,768,14) + bias
def _suffix(base: str, i: int) -> str:
    return base if i == 0 else "%s_%d" % (base, i)


def conv3d_blocks():
    """[(conv3d name, bn name, dense name, stride (z,x,y), pad (z,x,y))] — addConv3DLayer calls at :236-238."""
    geo = [((2, 1, 1), (1, 1, 1)), ((1, 1, 1), (0, 1, 1)), ((2, 1, 1), (1, 1, 1))]
    return [(_suffix("conv3d", i), "batch_normalization_%d" % (3 + i), "dense_%d" % (3 + i), s, p)
            for i, (s, p) in enumerate(geo)]


def rpn_blocks():
    """[(list of (conv2d name, bn name, cin, cout, stride), (transpose name, k, stride))] — :245-251."""
    blocks, ci, bi = [], 0, 6
    for cin, cout, q, (k, s) in ((64, 128, 3, (3, 1)), (128, 128, 5, (2, 2)), (128, 256, 5, (4, 4))):
        convs = []
        for j in range(q + 1):
            convs.append((_suffix("conv2d", ci), "batch_normalization_%d" % bi, cin if j == 0 else cout, cout,
                          2 if j == 0 else 1))
            ci += 1
            bi += 1
        blocks.append((convs, (_suffix("conv2d_transpose", len(blocks)), k, s, cout)))
    return blocks


def network_shapes() -> dict:
    shapes = {}
    for c, b, d, _, _ in conv3d_blocks():
        shapes[c + "/kernel"] = (3, 3, 3, 64, 64)
        shapes[c + "/bias"] = (64,)
        for f in BN_FIELDS:
            shapes[b + "/" + f] = (64,)
        shapes[d + "/kernel"] = (64, 64)
    for convs, (tname, k, s, cin) in rpn_blocks():
        for c, b, ci, co, _ in convs:
            shapes[c + "/kernel"] = (3, 3, ci, co)
            shapes[c + "/bias"] = (co,)
            for f in BN_FIELDS:
                shapes[b + "/" + f] = (co,)
        shapes[tname + "/kernel"] = (k, k, 256, cin)
        shapes[tname + "/bias"] = (256,)
    shapes["ClassificationLayer/kernel"] = (1, 1, 768, 2)
    shapes["ClassificationLayer/bias"] = (2,)
    shapes["RegressionLayer/kernel"] = (1, 1, 768, 14)
    shapes["RegressionLayer/bias"] = (14,)
    return shapes


def synthetic_network_pack(seed: int = 0) -> dict:
    """Seeded stand-in for the dense half of the missing .h5: Glorot-uniform kernels (Keras default), small non-zero
    biases and non-trivial BN statistics. Merged with synthetic_vfe_pack(seed) it is a whole createModel()."""
    rng = np.random.default_rng(1000 + seed)
    pack = {}
    for name, shape in network_shapes().items():
        leaf = name.split("/")[1]
        if leaf == "kernel":
            rf = int(np.prod(shape[:-2]))
            # Keras Glorot: fan_in = rf * shape[-2], fan_out = rf * shape[-1] (same formula for transposed kernels)
            lim = np.sqrt(6.0 / (rf * (shape[-2] + shape[-1])))
            pack[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif leaf == "bias":
            pack[name] = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        elif leaf == "gamma":
            pack[name] = rng.uniform(0.8, 1.6, size=shape).astype(np.float32)
        elif leaf == "beta":
            pack[name] = rng.uniform(-0.2, 0.3, size=shape).astype(np.float32)
        elif leaf == "moving_mean":
            pack[name] = rng.uniform(-0.3, 0.3, size=shape).astype(np.float32)
        else:
            pack[name] = rng.uniform(0.3, 1.5, size=shape).astype(np.float32)
    return pack


def keras_default_init_pack(seed: int = 0) -> dict:
    """What createModel() (model_training.py:222-257) holds before fit(): Keras's default initialisers — Glorot-uniform
    kernels, zero biases, gamma = 1, beta = 0, moving_mean = 0, moving_variance = 1 (seeded here; Keras seeds from the clock)."""
    rng = np.random.default_rng(seed)
    shapes = {}
    for d, b, sh in zip(VFE_DENSE, VFE_BN, VFE_SHAPES):
        shapes[d + "/kernel"] = sh
        for f in BN_FIELDS:
            shapes[b + "/" + f] = (sh[1],)
    shapes.update(network_shapes())
    pack = {}
    for name, shape in shapes.items():
        leaf = name.split("/")[1]
        if leaf == "kernel":
            rf = int(np.prod(shape[:-2]))
            lim = np.sqrt(6.0 / (rf * (shape[-2] + shape[-1])))
            pack[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif leaf in ("gamma", "moving_variance"):
            pack[name] = np.ones(shape, np.float32)
        else:
            pack[name] = np.zeros(shape, np.float32)
    return pack


def synthetic_model_pack(seed: int = 0) -> dict:
    pack = synthetic_vfe_pack(seed)
    pack.update(synthetic_network_pack(seed))
    return pack


def validate_network_pack(pack: dict) -> dict:
    out = {}
    for name, shape in network_shapes().items():
        if name not in pack:
            raise KeyError("weight pack has no %s" % name)
        v = np.ascontiguousarray(pack[name], dtype=np.float32)
        if v.shape != shape:
            raise ValueError("%s has shape %s, expected %s" % (name, v.shape, shape))
        out[name] = v
    return out

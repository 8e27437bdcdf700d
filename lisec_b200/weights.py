"""VFE weight packs keyed by Keras layer names, in the creation order of the reference's createModel
(model_training.py:229-235). The shipped blob SampleModel/15SampleEpoch0.h5 is absent from the
reference checkout and h5py is not in this image, so packs travel as .npz with the same keys an .h5 would carry:

    dense/kernel (6,16)                batch_normalization/{gamma,beta,moving_mean,moving_variance} (16,)
    dense_1/kernel (32,32)             batch_normalization_1/{...} (32,)
    dense_2/kernel (64,64)             batch_normalization_2/{...} (64,)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Architecture:
    """Which of the reference's two graphs a weight set belongs to.

    post_dense = False  the current code: FCN = Dense -> BatchNormalization -> ReLU (model_training.py:169-174) and
                        Conv2D -> BatchNormalization -> ReLU (:201-208); VFE widths 16 | 32 | 64 (:231-233).
    post_dense = True   the graph model.png shows, i.e. the lines commented out at :172 and :205 switched on: FCN =
                        Dense -> BatchNormalization -> Dense(units, relu, no bias) and Conv2D -> BatchNormalization ->
                        Dense(relu); VFE widths 16 | 64 | 128, so the grid and the first Conv3D's input carry 128 channels.
    A Keras .h5 embeds its own graph and Predict.py:51-52 uses load_model(), so either may arrive; detect_architecture()
    tells them apart by the kernels' shapes. Keras numbers Dense layers in creation order, so the names shift:
    post_dense puts two Dense layers into every FCN (dense, dense_1 | dense_2, dense_3 | dense_4, dense_5), three behind
    the Conv3D blocks (dense_6..8) and sixteen behind the RPN's Conv2D layers (dense_9..24): 25, as model.png counts."""
    c1: int = 16
    c2: int = 32
    c3: int = 64
    post_dense: bool = False

    @property
    def widths(self):
        return (self.c1, self.c2, self.c3)


CURRENT = Architecture()
MODEL_PNG = Architecture(16, 64, 128, True)
SUPPORTED_WIDTHS = ((16, 32, 64), (16, 64, 128))  # what liblisec_b200.so instantiates (lisec_create checks)


def _suffix(base: str, i: int) -> str:
    return base if i == 0 else "%s_%d" % (base, i)


def vfe_layers(arch: Architecture = CURRENT):
    """[(dense name, bn name, post-dense name or None, cin, cout)] of the three FCNs (:231-233) in creation order."""
    out, n = [], 0
    for i, (cin, cout) in enumerate(((6, arch.c1), (2 * arch.c1, arch.c2), (2 * arch.c2, arch.c3))):
        d = _suffix("dense", n)
        n += 1
        post = None
        if arch.post_dense:
            post = _suffix("dense", n)
            n += 1
        out.append((d, _suffix("batch_normalization", i), post, cin, cout))
    return out


def _dense_after_vfe(arch: Architecture) -> int:
    return 6 if arch.post_dense else 3


def detect_architecture(pack: dict) -> Architecture:
    """The graph a Keras-named weight set was saved from, by the shapes of its first Dense kernels."""
    k0 = np.shape(pack["dense/kernel"])
    k1 = np.shape(pack["dense_1/kernel"])
    if len(k0) != 2 or k0[0] != 6 or len(k1) != 2:
        raise ValueError("dense/kernel %s, dense_1/kernel %s: not a VoxelNet front end of this reference" % (k0, k1))
    c1 = int(k0[1])
    if k1 == (c1, c1):  # the FCN's second Dense
        if "dense_5/kernel" not in pack:
            raise ValueError("dense_1/kernel is (%d,%d) — the Dense-BN-Dense FCN — but dense_5/kernel is missing" % k1)
        return Architecture(c1, int(np.shape(pack["dense_2/kernel"])[1]), int(np.shape(pack["dense_4/kernel"])[1]), True)
    if k1[0] != 2 * c1:
        raise ValueError("dense_1/kernel has %d rows: expected %d (concat) or %d (second FCN Dense)" % (k1[0], 2 * c1, c1))
    return Architecture(c1, int(k1[1]), int(np.shape(pack["dense_2/kernel"])[1]), False)


VFE_DENSE = ("dense", "dense_1", "dense_2")
VFE_BN = ("batch_normalization", "batch_normalization_1", "batch_normalization_2")
VFE_SHAPES = ((6, 16), (32, 32), (64, 64))
BN_FIELDS = ("gamma", "beta", "moving_mean", "moving_variance")


def vfe_keys(arch: Architecture = CURRENT):
    keys = []
    for d, b, post, _, _ in vfe_layers(arch):
        keys.append(d + "/kernel")
        keys.extend(b + "/" + f for f in BN_FIELDS)
        if post:
            keys.append(post + "/kernel")
    return keys


def synthetic_vfe_pack(seed: int = 0, arch: Architecture = CURRENT) -> dict:
    """Seeded stand-in for the missing .h5: Glorot-uniform kernels (the Keras Dense default) and non-trivial BN
    statistics, so that c_empty != 0 and the pad rows really do take part in the max-pools."""
    rng = np.random.default_rng(seed)
    pack = {}
    for d, b, post, cin, cout in vfe_layers(arch):
        lim = np.sqrt(6.0 / (cin + cout))
        pack[d + "/kernel"] = rng.uniform(-lim, lim, size=(cin, cout)).astype(np.float32)
        pack[b + "/gamma"] = rng.uniform(0.5, 1.5, size=cout).astype(np.float32)
        pack[b + "/beta"] = rng.uniform(-0.3, 0.3, size=cout).astype(np.float32)
        pack[b + "/moving_mean"] = rng.uniform(-0.5, 0.5, size=cout).astype(np.float32)
        pack[b + "/moving_variance"] = rng.uniform(0.3, 2.0, size=cout).astype(np.float32)
        if post:
            lim = np.sqrt(6.0 / (2 * cout))
            pack[post + "/kernel"] = rng.uniform(-lim, lim, size=(cout, cout)).astype(np.float32)
    return pack


def validate_vfe_pack(pack: dict, arch: Architecture = CURRENT) -> dict:
    out = {}
    for d, b, post, cin, cout in vfe_layers(arch):
        for name, shape in ((d, (cin, cout)),) + (((post, (cout, cout)),) if post else ()):
            if name + "/kernel" not in pack:
                raise KeyError("weight pack has no %s/kernel" % name)
            k = np.ascontiguousarray(pack[name + "/kernel"], dtype=np.float32)
            if k.shape != shape:
                raise ValueError("%s/kernel has shape %s, expected %s" % (name, k.shape, shape))
            out[name + "/kernel"] = k
        for f in BN_FIELDS:
            v = np.ascontiguousarray(pack[b + "/" + f], dtype=np.float32)
            if v.shape != (cout,):
                raise ValueError("%s/%s has shape %s, expected (%d,)" % (b, f, v.shape, cout))
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
def conv3d_blocks(arch: Architecture = CURRENT):
    """[(conv3d name, bn name, dense name, stride (z,x,y), pad (z,x,y))] — addConv3DLayer calls at :236-238."""
    geo = [((2, 1, 1), (1, 1, 1)), ((1, 1, 1), (0, 1, 1)), ((2, 1, 1), (1, 1, 1))]
    d0 = _dense_after_vfe(arch)
    return [(_suffix("conv3d", i), "batch_normalization_%d" % (3 + i), "dense_%d" % (d0 + i), s, p)
            for i, (s, p) in enumerate(geo)]


def conv2d_post_dense(arch: Architecture = CURRENT) -> dict:
    """{conv2d name: name of the Dense(relu) behind its BatchNormalization} — the variant of addConv2DLayer commented out
    at :205 (empty for the current code, whose Conv2D layers end in Activation('relu'))."""
    if not arch.post_dense:
        return {}
    d0 = _dense_after_vfe(arch) + 3
    return {_suffix("conv2d", i): "dense_%d" % (d0 + i) for i in range(16)}


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


def network_shapes(arch: Architecture = CURRENT) -> dict:
    shapes = {}
    for i, (c, b, d, _, _) in enumerate(conv3d_blocks(arch)):
        shapes[c + "/kernel"] = (3, 3, 3, arch.c3 if i == 0 else 64, 64)
        shapes[c + "/bias"] = (64,)
        for f in BN_FIELDS:
            shapes[b + "/" + f] = (64,)
        shapes[d + "/kernel"] = (64, 64)
    post = conv2d_post_dense(arch)
    for convs, (tname, k, s, cin) in rpn_blocks():
        for c, b, ci, co, _ in convs:
            shapes[c + "/kernel"] = (3, 3, ci, co)
            shapes[c + "/bias"] = (co,)
            for f in BN_FIELDS:
                shapes[b + "/" + f] = (co,)
            if c in post:
                shapes[post[c] + "/kernel"] = (co, co)
        shapes[tname + "/kernel"] = (k, k, 256, cin)
        shapes[tname + "/bias"] = (256,)
    shapes["ClassificationLayer/kernel"] = (1, 1, 768, 2)
    shapes["ClassificationLayer/bias"] = (2,)
    shapes["RegressionLayer/kernel"] = (1, 1, 768, 14)
    shapes["RegressionLayer/bias"] = (14,)
    return shapes


def synthetic_network_pack(seed: int = 0, arch: Architecture = CURRENT) -> dict:
    """Seeded stand-in for the dense half of the missing .h5: Glorot-uniform kernels (Keras default), small non-zero
    biases and non-trivial BN statistics. Merged with synthetic_vfe_pack(seed) it is a whole createModel()."""
    rng = np.random.default_rng(1000 + seed)
    pack = {}
    for name, shape in network_shapes(arch).items():
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


def synthetic_model_pack(seed: int = 0, arch: Architecture = CURRENT) -> dict:
    pack = synthetic_vfe_pack(seed, arch)
    pack.update(synthetic_network_pack(seed, arch))
    return pack


# model.png's graph predates the heads' names (:253-254): Keras auto-named them after the 16 RPN convolutions
HEAD_ALIASES = {"ClassificationLayer": "conv2d_16", "RegressionLayer": "conv2d_17"}


def validate_network_pack(pack: dict, arch: Architecture = CURRENT) -> dict:
    out = {}
    for name, shape in network_shapes(arch).items():
        layer, leaf = name.split("/")
        if name not in pack and layer in HEAD_ALIASES and HEAD_ALIASES[layer] + "/" + leaf in pack:
            pack = dict(pack)
            pack[name] = pack[HEAD_ALIASES[layer] + "/" + leaf]
        if name not in pack:
            raise KeyError("weight pack has no %s" % name)
        v = np.ascontiguousarray(pack[name], dtype=np.float32)
        if v.shape != shape:
            raise ValueError("%s has shape %s, expected %s" % (name, v.shape, shape))
        out[name] = v
    return out

"""Host logic for the two graphs the reference's .h5 may hold, on the CPU: Keras layer names in creation
order, telling the graphs apart by the kernels' shapes, the oracle's two FCN variants, and the .h5 round trip."""
import numpy as np
import pytest

from lisec_b200 import weights as W
from oracle import lisec_oracle as O
from oracle import network_oracle as NO


def test_layer_names_follow_keras_creation_order():
    assert [l[:3] for l in W.vfe_layers(W.CURRENT)] == [("dense", "batch_normalization", None),
                                                         ("dense_1", "batch_normalization_1", None),
                                                         ("dense_2", "batch_normalization_2", None)]
    assert [l[:3] for l in W.vfe_layers(W.MODEL_PNG)] == [("dense", "batch_normalization", "dense_1"),
                                                           ("dense_2", "batch_normalization_1", "dense_3"),
                                                           ("dense_4", "batch_normalization_2", "dense_5")]
    assert [b[2] for b in W.conv3d_blocks(W.CURRENT)] == ["dense_3", "dense_4", "dense_5"]
    assert [b[2] for b in W.conv3d_blocks(W.MODEL_PNG)] == ["dense_6", "dense_7", "dense_8"]
    post = W.conv2d_post_dense(W.MODEL_PNG)
    assert post["conv2d"] == "dense_9" and post["conv2d_15"] == "dense_24" and W.conv2d_post_dense(W.CURRENT) == {}
    pack = W.synthetic_model_pack(0, W.MODEL_PNG)
    assert sum(k.startswith("dense") for k in pack) == 25  # model.png: 25 dense_* layers
    assert pack["conv3d/kernel"].shape == (3, 3, 3, 128, 64)


@pytest.mark.parametrize("arch", [W.CURRENT, W.MODEL_PNG, W.Architecture(16, 32, 64, True), W.Architecture(16, 64, 128, False)])
def test_the_graph_is_read_off_the_kernel_shapes(arch):
    pack = W.synthetic_model_pack(3, arch)
    assert W.detect_architecture(pack) == arch
    W.validate_vfe_pack(pack, arch)
    W.validate_network_pack(pack, arch)
    other = W.MODEL_PNG if arch == W.CURRENT else W.CURRENT
    with pytest.raises((KeyError, ValueError)):
        W.validate_vfe_pack(pack, other)


def test_heads_of_the_older_graph_may_carry_their_auto_names():
    pack = W.synthetic_model_pack(0, W.MODEL_PNG)
    for head, alias in W.HEAD_ALIASES.items():
        for leaf in ("kernel", "bias"):
            pack[alias + "/" + leaf] = pack.pop(head + "/" + leaf)
    out = W.validate_network_pack(pack, W.MODEL_PNG)
    assert out["ClassificationLayer/kernel"].shape == (1, 1, 768, 2) and out["RegressionLayer/bias"].shape == (14,)


def test_oracle_fcn_variants():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4, 35, 6))
    x[:, 20:] = 0.0
    for arch in (W.CURRENT, W.MODEL_PNG):
        pack = W.synthetic_vfe_pack(1, arch)
        y = O.vfe_forward(x, pack)  # variant read off the shapes
        assert y.shape == (4, arch.c3) and np.array_equal(y, O.vfe_forward(x, pack, post_dense=arch.post_dense))
        # spelled out for voxel 0, float64: Dense -> BN -> (Dense) -> ReLU, max over all 35 rows, [pooled, pointwise]
        h = x[0]
        for i, (d, bn, post, _, _) in enumerate(W.vfe_layers(arch)):
            z = h @ pack[d + "/kernel"].astype(np.float64)
            g = pack[bn + "/gamma"].astype(np.float64) / np.sqrt(pack[bn + "/moving_variance"].astype(np.float64) + 1e-3)
            z = (z - pack[bn + "/moving_mean"]) * g + pack[bn + "/beta"]
            if post:
                z = z @ pack[post + "/kernel"].astype(np.float64)
            z = np.maximum(z, 0.0)
            h = z if i == 2 else np.concatenate([np.repeat(z.max(0, keepdims=True), 35, 0), z], axis=1)
        assert np.allclose(h.max(0), y[0], rtol=1e-12, atol=1e-12)
    # the two variants differ on the same first-layer weights
    p = W.synthetic_vfe_pack(1, W.MODEL_PNG)
    assert O.c_empty(p, 35).shape == (128,)


def test_network_oracle_older_graph_applies_the_post_dense():
    arch = W.MODEL_PNG
    pack = W.synthetic_model_pack(2, arch)
    grid = np.random.default_rng(1).normal(size=(1, 8, 8, 16, 128)).astype(np.float32)
    p0, r0 = NO.network_forward(grid, pack, arch=arch)
    assert p0.shape == (1, 4, 8, 2) and r0.shape == (1, 4, 8, 14)
    pack2 = dict(pack)
    pack2["dense_24/kernel"] = pack["dense_24/kernel"] * 0.5  # the Dense behind the last RPN convolution
    p1, _ = NO.network_forward(grid, pack2, arch=arch)
    assert not np.allclose(p0, p1)


def test_h5_round_trip_of_the_older_graph(tmp_path):
    from lisec_b200.h5weights import read_keras_weights
    from lisec_b200.h5write import write_keras_weights

    pack = W.synthetic_model_pack(4, W.MODEL_PNG)
    path = str(tmp_path / "older.h5")
    write_keras_weights(path, pack)
    back = read_keras_weights(path)
    assert set(back) == set(pack) and all(np.array_equal(back[k], pack[k]) for k in pack)
    assert W.detect_architecture(back) == W.MODEL_PNG

"""The training-step oracle (oracle/train_oracle.py; reference model_training.py:222-257, 295-299): consistency with the
inference oracles, gradients against finite differences, the Keras SGD formula, and that fit() steps reduce the loss."""
import numpy as np
import pytest
import torch

from lisec_b200.weights import synthetic_model_pack
from oracle import lisec_oracle as O
from oracle import network_oracle as NO
from oracle import train_oracle as TO

ARGS = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=5, maxVoxelX=8, maxVoxelY=4, maxVoxelZ=8)
GRID = (8, 16, 8)  # nz, nx, ny: the smallest grid the three stride-2 RPN blocks accept


def toy_input(seed=0):
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.uniform([-3.9, -0.9, 0.0], [3.9, 0.9, 1.95], size=(160, 3)),
                          rng.uniform([0.5, 0.25, 0.5], [1.0, 0.5, 0.75], size=(9, 3))])
    vox = O.voxelize_np(pts, **ARGS)
    ind, val = O.coo_from_voxels(vox, 5)
    return O.to_dense(ind, val, list(GRID) + [5, 6])[None]  # [1, 8, 16, 8, 5, 6]


def labels(seed=0):
    rng = np.random.default_rng(100 + seed)
    return (rng.integers(0, 3, size=(1, 8, 4, 2)).astype(np.float64), rng.normal(0, 0.5, size=(1, 8, 4, 14)))


def test_training_forward_with_batch_statistics_as_moving_statistics_is_the_inference_forward():
    """Training-mode BN normalises with the batch statistics; put those into moving_mean / moving_variance and the
    inference oracles (lisec_oracle.vfe_forward on the dense input, network_oracle.network_forward) must agree."""
    pack = synthetic_model_pack(3)
    x = toy_input()
    p = TO.to_params(pack)
    with torch.no_grad():
        prob, reg, stats, grid = TO.forward_train(torch.from_numpy(x), p)
    frozen = dict(pack)
    for bn, (mean, var) in stats.items():
        frozen[bn + "/moving_mean"], frozen[bn + "/moving_variance"] = mean.numpy(), var.numpy()
    grid_inf = O.vfe_forward(x[0], frozen)
    assert np.abs(grid_inf - grid[0].numpy()).max() < 1e-10
    want_p, want_r = NO.network_forward(grid_inf[None], frozen)
    assert np.abs(want_p - prob.numpy()).max() < 1e-9 and np.abs(want_r - reg.numpy()).max() < 1e-9
    assert len(stats) == 3 + 3 + 16  # VFE, Conv3D blocks, RPN convolutions


def test_gradients_match_finite_differences():
    pack = synthetic_model_pack(4)
    x, (yc, yr) = toy_input(1), labels(1)
    loss, _, _, grads = TO.train_step(pack, {}, 0, x, yc, yr)
    assert set(grads) == {k for k in pack if "moving_" not in k} and sum(g.size for g in grads.values()) == 6_491_024

    def loss_at(p2):
        with torch.no_grad():
            pr, rg, _, _ = TO.forward_train(torch.from_numpy(x), TO.to_params(p2))
            return float(TO.loss_mse2(pr, rg, torch.from_numpy(yc), torch.from_numpy(yr)))

    rng = np.random.default_rng(2)
    for name in ("dense/kernel", "batch_normalization_1/gamma", "dense_2/kernel", "conv3d/kernel", "dense_4/kernel",
                 "conv2d_7/bias", "batch_normalization_12/beta", "conv2d_transpose_1/kernel", "RegressionLayer/kernel"):
        g = grads[name]
        idx = np.unravel_index(int(np.argmax(np.abs(g))), g.shape)
        h = 1e-7 * max(1.0, float(np.abs(pack[name][idx])))  # small: the loss has kinks (ReLU, max) everywhere
        hi, lo = dict(pack), dict(pack)
        hi[name] = np.asarray(pack[name], dtype=np.float64).copy()
        lo[name] = hi[name].copy()
        hi[name][idx] += h
        lo[name][idx] -= h
        fd = (loss_at(hi) - loss_at(lo)) / (2 * h)
        assert abs(fd - g[idx]) <= 2e-3 * max(abs(fd), abs(g[idx])) + 1e-8, (name, fd, g[idx])


def test_keras_sgd_nesterov_update_formula():
    rng = np.random.default_rng(0)
    var, acc, g = (rng.normal(size=1000).astype(np.float32) for _ in range(3))
    v1, a1 = TO.sgd_nesterov_update(var, acc, g, iterations=7)
    assert v1.dtype == np.float32 and a1.dtype == np.float32
    lr_t = 0.01 / (1 + 1e-6 * 7)
    a_ref = 0.9 * acc.astype(np.float64) - lr_t * g
    v_ref = var + 0.9 * a_ref - lr_t * g
    assert np.abs(a1 - a_ref).max() < 1e-6 and np.abs(v1 - v_ref).max() < 1e-6
    v2, a2 = TO.sgd_nesterov_update(var, acc, g, iterations=7, nesterov=False)
    assert np.abs(v2 - (var + a_ref)).max() < 1e-6 and np.array_equal(a1, a2)


def test_a_step_along_the_negative_gradient_descends_and_fit_moves_the_moving_statistics():
    """With the seeded synthetic weights the loss surface is steep (|grad|^2 ~ 1e7: training-mode BatchNormalization over
    a mostly-zero dense tensor divides by tiny batch deviations), so the first-order regime is checked where it holds."""
    pack = synthetic_model_pack(5)
    x, (yc, yr) = toy_input(2), labels(2)
    loss, new_pack, accum, grads = TO.train_step(pack, {}, 0, x, yc, yr)
    gn2 = sum(float((g ** 2).sum()) for g in grads.values())
    eps = 1e-3 / gn2  # predicted decrease 1e-3
    moved = {k: (np.asarray(v, dtype=np.float64) - eps * grads[k] if k in grads else v) for k, v in pack.items()}
    with torch.no_grad():
        pr, rg, _, _ = TO.forward_train(torch.from_numpy(x), TO.to_params(moved))
        loss2 = float(TO.loss_mse2(pr, rg, torch.from_numpy(yc), torch.from_numpy(yr)))
    assert 0.5e-3 < loss - loss2 < 1.5e-3, (loss, loss2)
    # the Keras update itself: first step from zero accumulators moves every weight by -(1 + momentum) * lr_t * grad
    for k in ("dense/kernel", "conv2d_3/kernel", "RegressionLayer/bias"):
        assert np.allclose(new_pack[k], np.asarray(pack[k], dtype=np.float64) - 1.9 * 0.01 * grads[k], rtol=0, atol=1e-12)
        assert np.allclose(accum[k], -0.01 * grads[k], rtol=0, atol=1e-14)
    for bn in ("batch_normalization", "batch_normalization_4", "batch_normalization_20"):
        for f in ("moving_mean", "moving_variance"):
            assert np.abs(new_pack[bn + "/" + f] - pack[bn + "/" + f]).max() > 0
    # second step: lr_t decays with the iteration count, the accumulator carries over
    _, pack2, accum2, grads2 = TO.train_step(new_pack, accum, 1, x, yc, yr)
    lr1 = 0.01 / (1 + 1e-6)
    k = "dense_2/kernel"
    assert np.allclose(accum2[k], 0.9 * accum[k] - lr1 * grads2[k], rtol=0, atol=1e-13)


def test_vfe_on_rows_with_multiplicities_equals_the_dense_graph():
    """kept rows (weight 1) + one virtual pad row per non-full voxel (weight T - s) + one empty row (weight
    35 * n_empty) reproduce the dense VFE graph — outputs, batch statistics and every parameter gradient."""
    pack = synthetic_model_pack(6)
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform([-3.9, -0.9, 0.0], [3.9, 0.9, 1.95], size=(160, 3)),
                          rng.uniform([0.5, 0.25, 0.5], [1.0, 0.5, 0.75], size=(9, 3)),
                          np.tile([[1.3, 0.3, 0.6]], (3, 1))])  # duplicate points: exact ties between kept rows
    vox = O.voxelize_np(pts, **ARGS)
    T = 5
    assert (vox["counts"] > T).any() and (vox["counts"] < T).any()
    ind, val = O.coo_from_voxels(vox, T)
    dense = torch.from_numpy(O.to_dense(ind, val, list(GRID) + [T, 6])[None])
    names = [k for k in pack if k.split("/")[0] in TO.VFE_DENSE + TO.VFE_BN and "moving_" not in k]
    gw = torch.from_numpy(rng.normal(size=(1,) + GRID + (64,)))  # an arbitrary upstream gradient on the grid

    p1 = TO.to_params(pack)
    _, _, stats1, grid1 = TO.forward_train(dense, p1)
    g1 = torch.autograd.grad((grid1 * gw).sum(), [p1[k] for k in names])

    kept = np.minimum(vox["counts"], T)
    feats = vox["features"]  # [V, T, 6], zero padded
    rows = torch.from_numpy(np.concatenate([feats[v, :kept[v]] for v in range(len(kept))]))
    row_voxel = torch.from_numpy(np.repeat(np.arange(len(kept)), kept))
    p2 = TO.to_params(pack)
    n_cells = int(np.prod(GRID))
    vout, eout, stats2 = TO.forward_train_rows(rows, row_voxel, torch.from_numpy(kept), n_cells, T, p2)
    grid2 = eout.expand(GRID + (64,)).clone()
    c = vox["coords"]
    grid2[c[:, 0], c[:, 1], c[:, 2]] = vout
    assert float((grid2 - grid1[0]).detach().abs().max()) < 1e-12
    for bn in TO.VFE_BN:
        assert float((stats1[bn][0] - stats2[bn][0]).abs().max()) < 1e-13
        assert float((stats1[bn][1] - stats2[bn][1]).abs().max()) < 1e-13
    g2 = torch.autograd.grad((grid2 * gw[0]).sum(), [p2[k] for k in names])
    for k, a, b in zip(names, g1, g2):
        assert float((a - b).abs().max()) <= 1e-10 * max(1.0, float(a.abs().max())), k


def test_teacher_forcing_replaces_values_and_keeps_the_gradient_path():
    """network_forward_train(teacher=...): the straight-through device the GPU backward-chain gate is built on
    (tests/test_gpu_train_network.py). Forcing the oracle's OWN outputs changes nothing; forcing other values moves the
    loss residual (and with it every gradient) while the gradient still reaches the layers in front of the forced one."""
    from lisec_b200.weights import synthetic_network_pack

    pack = synthetic_network_pack(1)
    g = torch.Generator().manual_seed(0)
    grid = torch.rand((1, 8, 8, 16, 64), generator=g, dtype=torch.float64)
    yc = torch.rand((1, 4, 8, 2), generator=g, dtype=torch.float64)
    yr = torch.rand((1, 4, 8, 14), generator=g, dtype=torch.float64)

    def run(teacher):
        p = TO.to_params(pack)
        prob, reg = TO.network_forward_train(grid, p, {}, teacher=teacher)
        loss = TO.loss_mse2(prob, reg, yc, yr)
        names = ["conv3d/kernel", "conv2d_7/kernel", "ClassificationLayer/bias"]
        return prob.detach(), reg.detach(), float(loss.detach()), torch.autograd.grad(loss, [p[k] for k in names])

    prob, reg, loss, grads = run(None)
    _, _, loss_same, grads_same = run({"ClassificationLayer": prob, "RegressionLayer": reg})
    assert loss_same == loss and all(torch.equal(a, b) for a, b in zip(grads, grads_same))
    p2, _, loss_forced, grads_forced = run({"ClassificationLayer": prob + 0.5})
    assert torch.equal(p2, prob + 0.5) and loss_forced != loss
    assert float(grads_forced[0].abs().max()) > 0 and not torch.allclose(grads_forced[0], grads[0])
    with pytest.raises(ValueError):
        run({"conv2d_7": prob})

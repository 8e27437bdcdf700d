"""The dense network behind the voxel grid in training mode, chained from the per-layer stages (lisec_b200/train.py:
DenseNetworkTrainer): forward, the two MSE terms and the gradient of every parameter against the float64 autograd oracle
(oracle/train_oracle.py: network_forward_train) from the same bf16 grid and bf16-representable weights."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(got, want):
    got, want = got.double().cpu(), want.double()
    return float(((got - want) ** 2).sum().sqrt() / ((want ** 2).sum().sqrt() + 1e-30))


def compare_gradients(net, grads, rel_l2, verbose=True):
    """The chain's parameter gradients (plan layouts) against the oracle's (Keras layouts): rel-L2, cosine, norm ratio."""
    worst, cosines, ratios = {}, {}, {}
    for k, gw in grads.items():
        layer, field = k.split("/")
        if layer in ("ClassificationLayer", "RegressionLayer"):
            rows = slice(0, 2) if layer == "ClassificationLayer" else slice(2, 16)
            got = net.grads["heads/kernel"][0, rows].t() if field == "kernel" else net.grads["heads/bias"][rows]
            want = gw[0, 0] if field == "kernel" else gw
        elif field != "kernel":
            got, want = net.grads[k], gw
            if field == "bias" and layer.startswith(("conv3d", "conv2d")) and "transpose" not in layer:
                continue  # a bias in front of a training-mode BatchNormalization: zero gradient up to rounding
        elif layer.startswith("conv3d"):
            got, want = net.grads[k].reshape(3, 3, 3, 64, 64).permute(0, 1, 2, 4, 3), gw
        elif layer.startswith("dense"):
            got, want = net.grads[k][0].t(), gw
        elif layer.startswith("conv2d_transpose"):
            G = net.grads[k]
            got = G.reshape(3, 3, G.shape[1], G.shape[2]).flip(0, 1) if G.dim() == 3 else G
            want = gw
        else:
            G = net.grads[k]
            got, want = G.reshape(3, 3, G.shape[1], G.shape[2]).permute(0, 1, 3, 2), gw
        assert tuple(got.shape) == tuple(want.shape), (k, tuple(got.shape), tuple(want.shape))
        worst[k] = rel_l2(got, want)
        gg, ww = got.double().cpu().reshape(-1), want.double().reshape(-1)
        cosines[k] = float((gg * ww).sum() / (gg.norm() * ww.norm() + 1e-30))
        ratios[k] = float(gg.norm() / (ww.norm() + 1e-30))
        if verbose:
            print("%-36s rel-L2 %.3f  cos %.4f  |got|/|want| %.3f" % (k, worst[k], cosines[k], ratios[k]))
    return worst, cosines, ratios


@pytest.mark.parametrize("nx,ny,B,keras_init,bf16_oracle", [(24, 40, 2, False, False), (48, 80, 2, True, False),
                                                            (24, 40, 2, False, True)])
def test_dense_network_training_forward_loss_and_gradients(nx, ny, B, keras_init, bf16_oracle):
    from lisec_b200.train import DenseNetworkTrainer
    from lisec_b200.weights import synthetic_network_pack
    from oracle import train_oracle as TO

    pack = {k: np.asarray(v, dtype=np.float32) for k, v in synthetic_network_pack(3).items()}
    if keras_init:  # what createModel() starts train() from: Glorot kernels, zero biases, gamma 1, beta 0
        rng = np.random.default_rng(7)
        for k, v in pack.items():
            if k.endswith("/kernel"):
                fan_in, fan_out = int(np.prod(v.shape[:-1])), int(np.prod(v.shape[:-2]) * v.shape[-1])
                if "transpose" in k:
                    fan_in, fan_out = int(np.prod(v.shape[:2]) * v.shape[3]), int(np.prod(v.shape[:2]) * v.shape[2])
                pack[k] = rng.normal(0, np.sqrt(2.0 / (fan_in + fan_out)), size=v.shape).astype(np.float32)
            elif k.endswith(("/bias", "/beta", "/moving_mean")):
                pack[k] = np.zeros_like(v)
            else:
                pack[k] = np.ones_like(v)
    for k in pack:  # bf16-representable weights: the GPU's operand copies then equal the oracle's weights
        if k.endswith("/kernel"):
            pack[k] = torch.from_numpy(pack[k]).to(torch.bfloat16).float().numpy()
    g = torch.Generator(device="cpu").manual_seed(43)
    grid = torch.rand((B, 8, nx, ny, 64), generator=g).to(torch.bfloat16)
    yc = torch.randint(0, 3, (B, nx // 2, ny // 2, 2), generator=g).float()
    yr = torch.randn((B, nx // 2, ny // 2, 14), generator=g) * 0.5
    net = DenseNetworkTrainer(pack, B, nx, ny)
    prob, reg = net.forward(grid.cuda())
    loss = net.loss_and_backward(yc.cuda(), yr.cuda())
    torch.cuda.synchronize()

    p = TO.to_params(pack)
    want_p, want_r = TO.network_forward_train(grid.double(), p, {}, bf16_activations=bf16_oracle)
    want_loss = TO.loss_mse2(want_p, want_r, yc.double(), yr.double())
    names = [k for k, t in p.items() if t.requires_grad]
    grads = dict(zip(names, torch.autograd.grad(want_loss, [p[k] for k in names])))

    ep, er = rel_l2(prob, want_p.detach()), rel_l2(reg, want_r.detach())
    print("forward rel-L2: prob %.3e regress %.3e; loss %.6f vs %.6f" % (ep, er, float(loss), float(want_loss.detach())))
    # 22 layers of bf16 activations, each renormalised by its own batch statistics: a few percent at the heads
    assert ep <= 8e-2 and er <= 8e-2
    assert abs(float(loss) - float(want_loss.detach())) <= 5e-2 * float(want_loss.detach())
    # parameter gradients, mapped back from the plans' layouts to the Keras ones
    worst, cosines, ratios = compare_gradients(net, grads, rel_l2)
    # What the chain reproduces tightly: everything whose gradient does not pass through a BatchNormalization backward fed
    # by a bf16 gradient tensor — the heads, the three transposed convolutions, the last BN of every RPN block.
    tight = [k for k in worst if k.split("/")[0] in ("ClassificationLayer", "RegressionLayer", "conv2d_transpose",
                                                      "conv2d_transpose_1", "conv2d_transpose_2", "batch_normalization_9",
                                                      "batch_normalization_15", "batch_normalization_21")]
    assert len(tight) == 16 and all(worst[k] <= 0.15 for k in tight), {k: worst[k] for k in tight}
    # Everything further back keeps direction and size (cosine 0.73-0.93 against float64, norms within 25 %) but not more:
    # the chained forward drifts from float64 by x1.17 per stage (0.17 % after one bf16 rounding, 11 % after 22 stages —
    # tools/bisect_train_chain.py: smooth growth, no jump, i.e. no semantic difference), because every layer is
    # renormalised by its own batch statistics; the parameter gradients there are weak correlations and amplify it.
    # This is the distance between two FREE-RUNNING networks, on record here; it is not the parity gate of the backward
    # pass. That is the test below: at the chain's own operating point every gradient is within 1 % of exact.
    rest = [k for k in worst if k not in tight]
    assert all(cosines[k] >= 0.7 for k in rest), {k: cosines[k] for k in rest if cosines[k] < 0.7}
    assert all(0.8 <= ratios[k] <= 1.25 for k in rest), {k: ratios[k] for k in rest if not 0.8 <= ratios[k] <= 1.25}
    net.close()


def gpu_activations(net):
    """{layer name: the output the GPU chain stored for that layer group}, in the oracle's layouts (train_oracle.
    network_forward_train's `teacher`)."""
    t = {}
    for st, conv, _, _ in net.c3:
        t[conv] = st.y.double().cpu().permute(0, 4, 1, 2, 3).contiguous()
    for bi, (stages, tail, tname, s, dy_t, x_out) in enumerate(net.blocks):
        for st, conv, _ in stages:
            t[conv] = st.bn.y.double().cpu()[:, 0].permute(0, 3, 1, 2).contiguous()
        t[tname] = net.concat.double().cpu()[:, 0, :, :, 256 * bi:256 * bi + 256].permute(0, 3, 1, 2).contiguous()
    y = net.heads.y.double().cpu()[:, 0]
    t["ClassificationLayer"], t["RegressionLayer"] = y[..., :2].contiguous(), y[..., 2:].contiguous()
    return t


@pytest.mark.parametrize("nx,ny,B,seed,bf16_oracle", [(24, 40, 2, 3, True), (48, 80, 2, 5, True), (24, 40, 2, 3, False)])
def test_chain_gradients_are_the_exact_gradients_at_the_chains_own_operating_point(nx, ny, B, seed, bf16_oracle):
    """The parity gate of the backward chain. The float64 oracle is evaluated AT THE GPU'S ACTIVATIONS: every layer group's
    output value is replaced by what the GPU chain stored for it, gradients still flow through the float64 layers
    (straight-through; oracle/train_oracle.py: network_forward_train(teacher=...)). Autograd then gives the exact gradient
    of every parameter for the ReLU masks, batch statistics and loss residual the GPU actually had. A wrong backward
    kernel, a mis-wired gradient tensor or a wrong layout would still disagree here; forward drift amplified by 22
    batch-normalised layers — the reason the distance to the free-running float64 network is large (the test above) —
    cannot. With bf16_oracle the oracle also rounds the tensors the chain stores in bf16 INSIDE a layer group (the
    convolution output in front of the BatchNormalization, straight-through): what remains is the rounding of the backward
    pass itself — bf16 operands in the data- and weight-gradient GEMMs, float32 sums — and every parameter gradient of the
    network must be within north_star's bf16 bar, 2e-2 (measured: worst 1.0e-2, median 5e-3, every cosine 1.0000).
    Without it (float64 inside the layer groups) the same comparison reads 5-15 % (cosines >= 0.989): the size of that one
    rounding's effect on batch statistics and weak-correlation gradients, recorded with a loose bar."""
    from lisec_b200.train import DenseNetworkTrainer
    from lisec_b200.weights import synthetic_network_pack
    from oracle import train_oracle as TO

    pack = {k: np.asarray(v, dtype=np.float32) for k, v in synthetic_network_pack(seed).items()}
    for k in pack:
        if k.endswith("/kernel"):
            pack[k] = torch.from_numpy(pack[k]).to(torch.bfloat16).float().numpy()
    g = torch.Generator(device="cpu").manual_seed(40 + seed)
    grid = torch.rand((B, 8, nx, ny, 64), generator=g).to(torch.bfloat16)
    yc = torch.randint(0, 3, (B, nx // 2, ny // 2, 2), generator=g).float()
    yr = torch.randn((B, nx // 2, ny // 2, 14), generator=g) * 0.5
    net = DenseNetworkTrainer(pack, B, nx, ny, need_grid_grad=True)
    net.forward(grid.cuda())
    loss = net.loss_and_backward(yc.cuda(), yr.cuda())
    torch.cuda.synchronize()

    p = TO.to_params(pack)
    grid64 = grid.double().requires_grad_()
    want_p, want_r = TO.network_forward_train(grid64, p, {}, teacher=gpu_activations(net), bf16_activations=bf16_oracle)
    want_loss = TO.loss_mse2(want_p, want_r, yc.double(), yr.double())
    assert abs(float(loss) - float(want_loss.detach())) <= 1e-5 * float(want_loss.detach())  # same outputs, same loss
    names = [k for k, t in p.items() if t.requires_grad]
    all_grads = torch.autograd.grad(want_loss, [p[k] for k in names] + [grid64])
    grads = dict(zip(names, all_grads[:-1]))
    worst, cosines, ratios = compare_gradients(net, grads, rel_l2, verbose=False)
    # d loss / d grid: what the VFE stack's backward pass receives from this chain (TrainStep)
    worst["d loss / d grid"] = rel_l2(net.grid_grad, all_grads[-1])
    gg, ww = net.grid_grad.double().cpu().reshape(-1), all_grads[-1].reshape(-1)
    cosines["d loss / d grid"] = float((gg * ww).sum() / (gg.norm() * ww.norm()))
    bar, cos_bar = (2e-2, 0.9995) if bf16_oracle else (0.2, 0.98)
    bad = {k: (round(worst[k], 4), round(cosines[k], 4)) for k in worst if worst[k] > bar or cosines[k] < cos_bar}
    print("teacher-forced (bf16 oracle %s): worst rel-L2 %.3e (%s), median %.3e, min cosine %.4f" %
          (bf16_oracle, max(worst.values()), max(worst, key=worst.get), float(np.median(list(worst.values()))),
           min(cosines.values())))
    assert not bad, bad
    net.close()

"""Training-step pieces (lisec_b200/train.py): flat buffers and the gradient all-reduce on CPU (gloo, world_size 2); the
SGD-Nesterov and MSE kernels on the GPU against oracle/train_oracle.py."""
import os

import numpy as np
import pytest
import torch

from lisec_b200.weights import synthetic_model_pack
from oracle import train_oracle as TO


def test_flat_parameters_layout():
    from lisec_b200.train import FlatParameters, trainable_names

    pack = synthetic_model_pack(0)
    fp = FlatParameters(pack, device="cpu")
    assert fp.numel == 6_491_024 and fp.numel_padded % 4 == 0 and fp.numel_padded >= fp.numel
    assert fp.names == trainable_names(pack) and all("moving_" not in k for k in fp.names)
    for k in ("dense/kernel", "conv3d_1/bias", "RegressionLayer/kernel"):
        assert fp.offsets[k] % 4 == 0
        assert np.array_equal(fp.view(fp.var, k).numpy(), pack[k].astype(np.float32))
    back = fp.to_pack()
    assert all(np.array_equal(back[k], pack[k].astype(np.float32)) for k in fp.names)


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from lisec_b200.train import allreduce_gradients

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    for w in allreduce_gradients(g, bucket_elems=300):  # 4 asynchronous buckets
        w.wait()
    # the update's grad_scale = 1 / world turns the sum into the mean
    var, acc = TO.sgd_nesterov_update(np.zeros(1000, np.float32), np.zeros(1000, np.float32), g.numpy(), 0,
                                      grad_scale=1.0 / world)
    if rank == 0:
        torch.save({"sum": g, "var": torch.from_numpy(var)}, out)
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29631, out), nprocs=2, join=True)
    res = torch.load(out)
    base = torch.arange(1000, dtype=torch.float32)
    assert torch.equal(res["sum"], base * 3)  # ranks contributed 1x and 2x
    mean = (base * 3).numpy() * np.float32(0.5)
    assert np.array_equal(res["var"].numpy(), (np.float32(0) + (-(np.float32(0.01) * mean) * np.float32(0.9) - np.float32(0.01) * mean)))


@pytest.mark.gpu
def test_gpu_sgd_nesterov_bit_exact_against_the_oracle():
    from lisec_b200.train import FlatParameters, SgdNesterov

    pack = synthetic_model_pack(1)
    fp = FlatParameters(pack)
    opt = SgdNesterov(fp)
    rng = np.random.default_rng(0)
    var = fp.var.cpu().numpy().copy()
    acc = np.zeros_like(var)
    for it, world in enumerate((1, 8, 2)):
        g = (rng.normal(size=fp.numel_padded) * 10 ** rng.uniform(-6, 2, size=fp.numel_padded)).astype(np.float32)
        fp.grad.copy_(torch.from_numpy(g))
        opt.step(world_size=world)
        var, acc = TO.sgd_nesterov_update(var, acc, g, it, grad_scale=1.0 / world)
        assert fp.var.cpu().numpy().tobytes() == var.tobytes() and fp.accum.cpu().numpy().tobytes() == acc.tobytes()
    assert opt.iterations == 3
    # plain momentum, odd length (scalar tail)
    import ctypes as C

    from lisec_b200 import _native as N

    lib = N.load()
    n = 1027
    v, a, g = (torch.from_numpy(rng.normal(size=n).astype(np.float32)).cuda() for _ in range(3))
    wv, wa = TO.sgd_nesterov_update(v.cpu().numpy(), a.cpu().numpy(), g.cpu().numpy(), 5, nesterov=False)
    st = lib.lisec_sgd_nesterov(v.data_ptr(), a.data_ptr(), g.data_ptr(), n, C.c_float(1.0),
                                C.c_float(np.float32(0.01 / (1 + 1e-6 * 5))), C.c_float(0.9), 0, None)
    torch.cuda.synchronize()
    assert st == 0 and v.cpu().numpy().tobytes() == wv.tobytes() and a.cpu().numpy().tobytes() == wa.tobytes()
    assert lib.lisec_sgd_nesterov(v.data_ptr() + 4, a.data_ptr(), g.data_ptr(), 8, C.c_float(1.0), C.c_float(0.01),
                                  C.c_float(0.9), 1, None) == -1  # LISEC_ERR_BAD_ARG


@pytest.mark.gpu
def test_gpu_mse_loss_and_gradient():
    from lisec_b200.train import mse_loss_grad

    rng = np.random.default_rng(1)
    for shape in ((2, 100, 200, 14), (1, 100, 200, 2), (3, 7)):
        y = rng.normal(size=shape).astype(np.float32)
        t = rng.integers(0, 3, size=shape).astype(np.float32)
        loss, dy = mse_loss_grad(torch.from_numpy(y).cuda(), torch.from_numpy(t).cuda())
        want = ((y - t).astype(np.float64) ** 2).mean()  # the difference is float32's, squares and sum are float64's
        assert abs(float(loss) - want) <= 1e-12 * max(1.0, want)
        want_dy = (y - t) * np.float32(2.0 / y.size)
        assert dy.cpu().numpy().tobytes() == want_dy.astype(np.float32).tobytes()
    loss, dy = mse_loss_grad(torch.from_numpy(y).cuda(), torch.from_numpy(t).cuda(), want_grad=False)
    assert dy is None


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    dict(B=2, D=4, H=24, W=40, C=64, N=64, k=(3, 3, 3), sd=1, pad=(0, 1, 1)),    # conv3d_1's geometry, 27 taps (odd units)
    dict(B=1, D=8, H=17, W=21, C=64, N=64, k=(3, 3, 3), sd=2, pad=(1, 1, 1)),    # stride 2 in depth, ragged tiles
    dict(B=2, D=1, H=24, W=40, C=128, N=128, k=(1, 3, 3), sd=1, pad=(0, 1, 1)),  # RPN 128 -> 128: 18 units, 3 passes
    dict(B=1, D=1, H=12, W=20, C=256, N=256, k=(1, 3, 3), sd=1, pad=(0, 1, 1)),  # RPN 256 -> 256: 9 passes
    dict(B=3, D=1, H=16, W=16, C=64, N=128, k=(1, 1, 1), sd=1, pad=(0, 0, 0)),   # 1x1: a single (half-used) pair
    dict(B=2, D=1, H=24, W=40, C=64, N=128, k=(1, 3, 3), sd=1, pad=(0, 1, 1), shw=2),  # first conv of an RPN block: stride 2
    dict(B=1, D=1, H=18, W=22, C=128, N=256, k=(1, 3, 3), sd=1, pad=(0, 1, 1), shw=2),
    dict(B=2, D=1, H=12, W=20, C=768, N=64, k=(1, 1, 1), sd=1, pad=(0, 0, 0)),   # the heads' 768 input channels (dy padded to 64)
    dict(B=2, D=1, H=3, W=5, C=256, N=256, k=(1, 3, 3), sd=1, pad=(0, 1, 1)),    # a map smaller than one tile
])
def test_gpu_conv_wgrad_matches_autograd(case):
    """dW from conv_wgrad_kernel against torch CPU float64 autograd of the same convolution on the same bf16 tensors."""
    import torch.nn.functional as F

    from lisec_b200.train import ConvWgrad

    c = case
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn((c["B"], c["D"], c["H"], c["W"], c["C"]), generator=g).to(torch.bfloat16)
    k = c["k"]
    shw = c.get("shw", 1)
    OD = (c["D"] + 2 * c["pad"][0] - k[0]) // c["sd"] + 1
    OH, OW = (c["H"] + 2 * c["pad"][1] - k[1]) // shw + 1, (c["W"] + 2 * c["pad"][2] - k[2]) // shw + 1
    dy = (torch.randn((c["B"], OD, OH, OW, c["N"]), generator=g) * 0.5).to(torch.bfloat16)
    wg = ConvWgrad(x.cuda(), dy.cuda(), k, c["sd"], c["pad"], stride_hw=shw)
    wg.dw.fill_(float("nan"))
    got = wg.run()
    torch.cuda.synchronize()
    got2 = wg.run().clone()  # a second launch: same bits (no atomics anywhere)
    torch.cuda.synchronize()
    w = torch.zeros((c["N"], c["C"]) + tuple(k), dtype=torch.float64, requires_grad=True)
    y = F.conv3d(x.double().permute(0, 4, 1, 2, 3), w, None, stride=(c["sd"], shw, shw), padding=c["pad"])
    (y * dy.double().permute(0, 4, 1, 2, 3)).sum().backward()
    want = w.grad.permute(2, 3, 4, 0, 1).reshape(k[0] * k[1] * k[2], c["N"], c["C"])  # [tap][co][ci]
    got = got.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max() / want.abs().max()
    assert err <= 2e-5, err  # bf16 products are exact in float32; what differs is the float32 summation order
    assert torch.equal(got2.cpu().double(), got)
    wg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    dict(B=2, D=4, H=24, W=40, C=64, N=64, k=(3, 3, 3), pad=(0, 1, 1)),    # conv3d_1: depth shrinks 4 -> 2, grows back
    dict(B=2, D=1, H=24, W=40, C=128, N=128, k=(1, 3, 3), pad=(0, 1, 1)),  # RPN 128 -> 128
    dict(B=1, D=1, H=12, W=20, C=128, N=256, k=(1, 3, 3), pad=(0, 1, 1)),  # 128 -> 256: dX has 128 channels
    dict(B=3, D=1, H=16, W=16, C=64, N=64, k=(1, 1, 1), pad=(0, 0, 0)),    # the Dense(64 -> 64) behind a Conv3D
])
def test_gpu_conv_dgrad_matches_autograd(case):
    """dX through a forward plan on dy with flipped, transposed weights, against torch CPU float64 autograd."""
    import torch.nn.functional as F

    from lisec_b200.train import ConvDgrad

    c = case
    k = c["k"]
    g = torch.Generator(device="cpu").manual_seed(9)
    OD = c["D"] + 2 * c["pad"][0] - k[0] + 1
    OH, OW = c["H"] + 2 * c["pad"][1] - k[1] + 1, c["W"] + 2 * c["pad"][2] - k[2] + 1
    dy = torch.randn((c["B"], OD, OH, OW, c["N"]), generator=g).to(torch.bfloat16)
    w = (torch.randn((k[0] * k[1] * k[2], c["N"], c["C"]), generator=g) * 0.05).to(torch.bfloat16).float()
    dg = ConvDgrad(dy.cuda(), w.cuda(), k, c["pad"], out_dtype=torch.float32)
    dg.dx.fill_(float("nan"))
    got = dg.run().cpu().double()
    x = torch.zeros((c["B"], c["C"], c["D"], c["H"], c["W"]), dtype=torch.float64, requires_grad=True)
    wt = w.double().reshape(k[0], k[1], k[2], c["N"], c["C"]).permute(3, 4, 0, 1, 2)
    y = F.conv3d(x, wt, None, stride=1, padding=c["pad"])
    (y * dy.double().permute(0, 4, 1, 2, 3)).sum().backward()
    want = x.grad.permute(0, 2, 3, 4, 1)
    assert got.shape == want.shape and torch.isfinite(got).all()
    err = (got - want).abs().max() / want.abs().max()
    assert err <= 2e-5, err
    dg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,relu", [((2, 1, 24, 40, 128), True), ((1, 4, 17, 21, 64), False), ((3, 1, 9, 7, 256), True),
                                        ((1, 1, 50, 100, 16), True)])
def test_gpu_batchnorm_train_forward_backward(shape, relu):
    """Training-mode BN (+ReLU) kernels against torch float64 autograd of the same formula on the same bf16 tensors."""
    from lisec_b200.train import BatchNormTrain

    g = torch.Generator(device="cpu").manual_seed(13)
    C = shape[-1]
    x = (torch.randn(shape, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    dy = torch.randn(shape, generator=g).to(torch.bfloat16)
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.2
    mm, mv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    bn = BatchNormTrain(x.cuda(), gamma.cuda(), beta.cuda(), mm.clone().cuda(), mv.clone().cuda(), relu=relu)
    y = bn.forward().float().cpu()
    dx = bn.backward(dy.cuda()).float().cpu()
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    axes = tuple(range(len(shape) - 1))
    mean, var = xd.mean(dim=axes), xd.var(dim=axes, unbiased=False)
    yd = (xd - mean) / torch.sqrt(var + 1e-3) * gd + bd
    if relu:
        yd = torch.relu(yd)
    assert (y.double() - yd.detach()).abs().max() <= 2.0 ** -8 * max(1.0, float(yd.detach().abs().max()))  # one bf16 rounding
    assert torch.allclose(bn.mean.cpu().double(), mean.detach(), rtol=0, atol=1e-6)
    assert torch.allclose(bn.invstd.cpu().double(), 1 / torch.sqrt(var.detach() + 1e-3), rtol=1e-6, atol=0)
    assert torch.allclose(bn.moving_mean.cpu().double(), mm.double() * 0.99 + mean.detach() * 0.01, rtol=0, atol=1e-6)
    assert torch.allclose(bn.moving_var.cpu().double(), mv.double() * 0.99 + var.detach() * 0.01, rtol=1e-6, atol=1e-7)
    # the kernel masks with ITS y (bf16): use the same mask in the reference so that borderline zeros cannot differ
    mask = (y > 0).double() if relu else torch.ones_like(y, dtype=torch.float64)
    yl = (xd - mean) / torch.sqrt(var + 1e-3) * gd + bd
    (yl * (dy.double() * mask)).sum().backward()
    scale = float(xd.grad.abs().max())
    assert (dx.double() - xd.grad).abs().max() <= 2.0 ** -8 * scale + 1e-6
    assert torch.allclose(bn.dgamma.cpu().double(), gd.grad, rtol=1e-5, atol=1e-4)
    assert torch.allclose(bn.dbeta.cpu().double(), bd.grad, rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    dict(B=2, D=1, H=24, W=40, C=128, N=128, k=(1, 3, 3), pad=(0, 1, 1)),  # addConv2DLayer(128, 128, 3, 1, 1)
    dict(B=2, D=4, H=16, W=24, C=64, N=64, k=(3, 3, 3), pad=(0, 1, 1)),    # Conv3D + BN of the second middle block
    dict(B=2, D=1, H=32, W=48, C=64, N=128, k=(1, 3, 3), pad=(0, 1, 1), s=2),  # addConv2DLayer(64, 128, 3, 2, 1): stride 2
])
def test_gpu_conv_bn_relu_layer_trains_like_autograd(case):
    """One conv(bias) -> BN(train) -> ReLU stage, forward and backward, then one SGD-Nesterov step on its weights:
    against torch float64 autograd from the same bf16 input and bf16-representable weights. The stage keeps the
    convolution output and dz in bf16 (2^-9): agreement 2-4e-3 of each tensor's scale, bar 1e-2."""
    import torch.nn.functional as F

    from lisec_b200.train import ConvBnReluTrain

    c = case
    k, pad = c["k"], c["pad"]
    g = torch.Generator(device="cpu").manual_seed(17)
    x = torch.randn((c["B"], c["D"], c["H"], c["W"], c["C"]), generator=g).to(torch.bfloat16)
    taps = k[0] * k[1] * k[2]
    w = (torch.randn((taps, c["N"], c["C"]), generator=g) * (1.0 / np.sqrt(taps * c["C"]))).to(torch.bfloat16).float()
    bias, beta = torch.randn(c["N"], generator=g) * 0.1, torch.randn(c["N"], generator=g) * 0.1
    gamma = torch.rand(c["N"], generator=g) + 0.5
    shw = c.get("s", 1)
    layer = ConvBnReluTrain(x.cuda(), w.cuda(), bias.cuda(), gamma.cuda(), beta.cuda(), k, pad, stride_hw=shw)
    y = layer.forward().float().cpu()
    dy = torch.randn(y.shape, generator=g).to(torch.bfloat16)
    dx = layer.backward(dy.cuda()).float().cpu()
    torch.cuda.synchronize()

    xd = x.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wd = w.double().reshape(k[0], k[1], k[2], c["N"], c["C"]).permute(3, 4, 0, 1, 2).requires_grad_(True)
    bd, gd, btd = bias.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    z = F.conv3d(xd, wd, bd, stride=(1, shw, shw), padding=pad)
    mean = z.mean(dim=(0, 2, 3, 4), keepdim=True)
    var = z.var(dim=(0, 2, 3, 4), unbiased=False, keepdim=True)
    lin = (z - mean) / torch.sqrt(var + 1e-3) * gd.view(1, -1, 1, 1, 1) + btd.view(1, -1, 1, 1, 1)
    yd = torch.relu(lin)
    # ReLU's gradient is a step: where |lin| is within bf16 rounding of zero the stage's mask (from ITS y) may differ from
    # float64's, an O(1) change at isolated elements that says nothing about the kernels — the reference takes the stage's mask
    mask = (y > 0).double().permute(0, 4, 1, 2, 3)
    assert float(((lin.detach() > 0).double() - mask).abs().mean()) < 5e-3
    (lin * mask * dy.double().permute(0, 4, 1, 2, 3)).sum().backward()

    def close(got, want, tol):
        got, want = got.double(), want.double()
        emax = float((got - want).abs().max()) / float(want.abs().max())
        el2 = float(((got - want) ** 2).sum().sqrt() / (want ** 2).sum().sqrt())
        print("max err / max |ref| = %.3e, rel-L2 = %.3e" % (emax, el2))
        return emax <= tol and el2 <= tol / 2

    assert close(y.permute(0, 4, 1, 2, 3), yd.detach(), 1e-2)
    assert close(dx.permute(0, 4, 1, 2, 3), xd.grad, 1e-2)
    want_dw = wd.grad.permute(2, 3, 4, 0, 1).reshape(taps, c["N"], c["C"])
    assert close(layer.dw.cpu(), want_dw, 1e-2)
    assert close(layer.bn.dgamma.cpu(), gd.grad, 1e-2) and close(layer.bn.dbeta.cpu(), btd.grad, 1e-2)
    # a bias in front of a training-mode BN has no effect: its gradient is zero up to rounding
    assert float(layer.dbias.abs().max()) <= 1e-2 * float(layer.bn.dbeta.abs().max()) and float(bd.grad.abs().max()) < 1e-9
    layer.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    dict(B=1, D=8, H=16, W=24, C=64, N=64, k=(3, 3, 3), pad=(1, 1, 1), sd=2, s=1),    # first / third Conv3D: stride 2 in depth
    dict(B=2, D=1, H=24, W=40, C=64, N=128, k=(1, 3, 3), pad=(0, 1, 1), sd=1, s=2),   # first conv of an RPN block
    dict(B=1, D=1, H=17, W=23, C=128, N=128, k=(1, 3, 3), pad=(0, 1, 1), sd=1, s=2),  # odd sizes: an untouched last row / column
])
def test_gpu_strided_conv_dgrad_matches_autograd(case):
    import torch.nn.functional as F

    from lisec_b200.train import ConvDgradStrided, relu_backward

    c = case
    k = c["k"]
    g = torch.Generator(device="cpu").manual_seed(23)
    OD = (c["D"] + 2 * c["pad"][0] - k[0]) // c["sd"] + 1
    OH, OW = (c["H"] + 2 * c["pad"][1] - k[1]) // c["s"] + 1, (c["W"] + 2 * c["pad"][2] - k[2]) // c["s"] + 1
    dy = torch.randn((c["B"], OD, OH, OW, c["N"]), generator=g).to(torch.bfloat16)
    w = (torch.randn((k[0] * k[1] * k[2], c["N"], c["C"]), generator=g) * 0.05).to(torch.bfloat16).float()
    dg = ConvDgradStrided(dy.cuda(), w.cuda(), k, c["pad"], c["sd"], c["s"], (c["D"], c["H"], c["W"]), out_dtype=torch.float32)
    got = dg.run().cpu().double()
    got2 = dg.run().cpu().double()  # the dilated buffer's zeros are still zeros
    x = torch.zeros((c["B"], c["C"], c["D"], c["H"], c["W"]), dtype=torch.float64, requires_grad=True)
    wt = w.double().reshape(k[0], k[1], k[2], c["N"], c["C"]).permute(3, 4, 0, 1, 2)
    y = F.conv3d(x, wt, None, stride=(c["sd"], c["s"], c["s"]), padding=c["pad"])
    (y * dy.double().permute(0, 4, 1, 2, 3)).sum().backward()
    want = x.grad.permute(0, 2, 3, 4, 1)
    assert got.shape == want.shape
    assert (got - want).abs().max() / want.abs().max() <= 2e-5 and torch.equal(got, got2)
    dg.close()
    # ReLU backward
    yy = torch.randn((3, 5, 64), generator=g).to(torch.bfloat16)
    dd = torch.randn((3, 5, 64), generator=g).to(torch.bfloat16)
    out = relu_backward(dd.cuda(), yy.cuda()).cpu()
    assert torch.equal(out, torch.where(yy > 0, dd, torch.zeros_like(dd)))


@pytest.mark.gpu
@pytest.mark.parametrize("sd,pad,D", [(2, (1, 1, 1), 8), (1, (0, 1, 1), 4)])
def test_gpu_conv3d_block_trains_like_autograd(sd, pad, D):
    """addConv3DLayer in training mode — Conv3D(bias) -> BN(batch statistics) -> Dense(64, no bias) -> ReLU — forward and
    backward against torch float64 autograd (the ReLU mask taken from the stage's own output, see the Conv2D test)."""
    import torch.nn.functional as F

    from lisec_b200.train import Conv3dBlockTrain

    g = torch.Generator(device="cpu").manual_seed(29)
    B, H, W, Cn = 2, 16, 24, 64
    k = (3, 3, 3)
    x = torch.randn((B, D, H, W, Cn), generator=g).to(torch.bfloat16)
    w = (torch.randn((27, Cn, Cn), generator=g) / np.sqrt(27 * Cn)).to(torch.bfloat16).float()
    wd = (torch.randn((1, Cn, Cn), generator=g) / np.sqrt(Cn)).to(torch.bfloat16).float()
    bias, beta = torch.randn(Cn, generator=g) * 0.1, torch.randn(Cn, generator=g) * 0.1
    gamma = torch.rand(Cn, generator=g) + 0.5
    blk = Conv3dBlockTrain(x.cuda(), w.cuda(), bias.cuda(), gamma.cuda(), beta.cuda(), wd.cuda(), k, pad, stride_d=sd)
    y = blk.forward().float().cpu()
    dy = torch.randn(y.shape, generator=g).to(torch.bfloat16)
    dx = blk.backward(dy.cuda()).float().cpu()
    torch.cuda.synchronize()

    xd = x.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wv = w.double().reshape(3, 3, 3, Cn, Cn).permute(3, 4, 0, 1, 2).requires_grad_(True)
    wdv = wd.double()[0].requires_grad_(True)  # [out][in]
    gd, btd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    z = F.conv3d(xd, wv, bias.double(), stride=(sd, 1, 1), padding=pad)
    mean, var = z.mean(dim=(0, 2, 3, 4), keepdim=True), z.var(dim=(0, 2, 3, 4), unbiased=False, keepdim=True)
    u = (z - mean) / torch.sqrt(var + 1e-3) * gd.view(1, -1, 1, 1, 1) + btd.view(1, -1, 1, 1, 1)
    lin = torch.einsum("ncdhw,kc->nkdhw", u, wdv)
    mask = (y > 0).double().permute(0, 4, 1, 2, 3)
    assert float(((lin.detach() > 0).double() - mask).abs().mean()) < 5e-3
    (lin * mask * dy.double().permute(0, 4, 1, 2, 3)).sum().backward()

    def close(got, want, tol=1e-2):
        got, want = got.double(), want.double()
        return float((got - want).abs().max()) <= tol * float(want.abs().max())

    assert close(y.permute(0, 4, 1, 2, 3), torch.relu(lin.detach()))
    assert close(dx.permute(0, 4, 1, 2, 3), xd.grad)
    assert close(blk.dw.cpu(), wv.grad.permute(2, 3, 4, 0, 1).reshape(27, Cn, Cn))
    assert close(blk.dwd.cpu()[0], wdv.grad)
    assert close(blk.bn.dgamma.cpu(), gd.grad) and close(blk.bn.dbeta.cpu(), btd.grad)
    blk.close()


@pytest.mark.gpu
def test_gpu_heads_train_forward_backward():
    """The two 1x1 head convolutions (768 -> 2 + 14, bias) with the MSE loss on top: forward, loss, and the gradients with
    respect to the head kernels, biases and the concat tensor, against torch float64 autograd."""
    from lisec_b200.train import HeadsTrain, mse_loss_grad

    g = torch.Generator(device="cpu").manual_seed(31)
    B, H, W = 2, 24, 40
    x = torch.randn((B, 1, H, W, 768), generator=g).to(torch.bfloat16)
    w = (torch.randn((1, 16, 768), generator=g) / np.sqrt(768)).to(torch.bfloat16).float()
    bias = torch.randn(16, generator=g) * 0.1
    tc = torch.randint(0, 3, (B, 1, H, W, 2), generator=g).float()
    tr = torch.randn((B, 1, H, W, 14), generator=g)
    heads = HeadsTrain(x.cuda(), w.cuda(), bias.cuda())
    y = heads.forward()
    lc, dyc = mse_loss_grad(y[..., :2].contiguous(), tc.cuda())
    lr, dyr = mse_loss_grad(y[..., 2:].contiguous(), tr.cuda())
    dy = torch.cat([dyc, dyr], dim=-1)
    dx = heads.backward(dy).float().cpu()
    torch.cuda.synchronize()

    xd = x.double().requires_grad_(True)
    wd, bd = w.double()[0].requires_grad_(True), bias.double().requires_grad_(True)
    yd = xd @ wd.t() + bd
    loss = ((yd[..., :2] - tc.double()) ** 2).mean() + ((yd[..., 2:] - tr.double()) ** 2).mean()
    loss.backward()
    assert abs(float(lc + lr) - float(loss.detach())) <= 1e-5 * float(loss.detach())
    assert float((y.cpu().double() - yd.detach()).abs().max()) <= 1e-4 * float(yd.detach().abs().max())

    def close(got, want, tol):
        return float((got.double() - want).abs().max()) <= tol * float(want.abs().max())

    # dy is rounded to bf16 on its way into the tensor-core operands: 2^-9
    assert close(heads.dw.cpu()[0], wd.grad, 5e-3) and close(heads.dbias.cpu(), bd.grad, 5e-3)
    assert close(dx, xd.grad, 1e-2)
    heads.close()
    # the same gradient as three dense 256-channel tensors (the dy of the three transposed-convolution stages), and a += b
    from lisec_b200.train import add_

    h3 = HeadsTrain(x.cuda(), w.cuda(), bias.cuda(), dense_slices=True)
    h3.forward()
    parts = h3.backward(dy)
    assert len(parts) == 3 and torch.equal(torch.cat([p.float().cpu() for p in parts], dim=-1), dx)
    acc = parts[0].clone()
    add_(acc, parts[1])
    assert torch.equal(acc.float().cpu(), (parts[0].float() + parts[1].float()).to(torch.bfloat16).float().cpu())
    h3.close()


@pytest.mark.gpu
def test_gpu_transposed_k3s1_stage_into_a_concat_slice():
    """RPN block 1's Conv2DTranspose(256, k3, s1, 'same') as its flipped-kernel convolution: forward into channels
    256..511 of a 768-channel buffer, backward from a dense dy, against float64 autograd of F.conv_transpose2d itself."""
    import torch.nn.functional as F

    from lisec_b200.train import ConvBiasTrain

    g = torch.Generator(device="cpu").manual_seed(37)
    B, H, W, Cin, Co = 2, 24, 40, 128, 256
    x = torch.randn((B, 1, H, W, Cin), generator=g).to(torch.bfloat16)
    Fk = (torch.randn((3, 3, Co, Cin), generator=g) / np.sqrt(9 * Cin)).to(torch.bfloat16).float()  # Keras (kh, kw, out, in)
    bias = torch.randn(Co, generator=g) * 0.1
    w = Fk.flip(0, 1).reshape(9, Co, Cin).contiguous()  # the plans' [tap][out][in] of the equivalent convolution
    concat = torch.full((B, 1, H, W, 768), 7.0, dtype=torch.bfloat16, device="cuda")
    dy = torch.randn((B, 1, H, W, Co), generator=g).to(torch.bfloat16)
    st = ConvBiasTrain(x.cuda(), w.cuda(), bias.cuda(), (1, 3, 3), (0, 1, 1), concat, 256, dy.cuda())
    st.forward()
    dx = st.backward().float().cpu()
    torch.cuda.synchronize()
    xd = x.double()[:, 0].permute(0, 3, 1, 2).requires_grad_(True)
    Fd = Fk.double().permute(3, 2, 0, 1).requires_grad_(True)  # torch conv_transpose2d weight: (in, out, kh, kw)
    bd = bias.double().requires_grad_(True)
    yd = F.conv_transpose2d(xd, Fd, bd, stride=1, padding=1)
    (yd * dy.double()[:, 0].permute(0, 3, 1, 2)).sum().backward()
    got_y = concat[:, 0, :, :, 256:512].float().cpu().permute(0, 3, 1, 2)
    assert float((got_y.double() - yd.detach()).abs().max()) <= 2.0 ** -8 * float(yd.detach().abs().max())
    assert bool((concat[..., :256] == 7).all()) and bool((concat[..., 512:] == 7).all())  # the neighbours are untouched

    def close(got, want, tol=2e-5):
        return float((got.double() - want).abs().max()) <= tol * float(want.abs().max())

    want_dw = Fd.grad.permute(2, 3, 1, 0).flip(0, 1).reshape(9, Co, Cin)  # back into the flipped [tap][out][in] layout
    assert close(st.dw.cpu(), want_dw) and close(st.dbias.cpu(), bd.grad, 1e-5)
    assert close(dx[:, 0].permute(0, 3, 1, 2), xd.grad, 2.0 ** -8)
    st.close()


@pytest.mark.gpu
@pytest.mark.parametrize("s,Ci,H,W", [(2, 128, 12, 20), (4, 256, 6, 10)])
def test_gpu_transposed_kernel_equals_stride_backward(s, Ci, H, W):
    """RPN block 2's Conv2DTranspose(256, k2, s2) and block 3's (k4, s4): dx, dF and dbias through the [B, H, s, W, s*Co]
    view of dy, against float64 autograd of F.conv_transpose2d."""
    import torch.nn.functional as F

    from lisec_b200.train import ConvTransposeBackward

    g = torch.Generator(device="cpu").manual_seed(41)
    B, Co = 2, 256
    x = torch.randn((B, 1, H, W, Ci), generator=g).to(torch.bfloat16)
    Fk = (torch.randn((s, s, Co, Ci), generator=g) / np.sqrt(Ci)).to(torch.bfloat16).float()
    dy = torch.randn((B, 1, s * H, s * W, Co), generator=g).to(torch.bfloat16)
    tb = ConvTransposeBackward(x.cuda(), dy.cuda(), Fk.cuda(), s)
    dx = tb.backward().float().cpu()
    torch.cuda.synchronize()
    xd = x.double()[:, 0].permute(0, 3, 1, 2).requires_grad_(True)
    Fd = Fk.double().permute(3, 2, 0, 1).requires_grad_(True)  # (in, out, kh, kw)
    bd = torch.zeros(Co, dtype=torch.float64, requires_grad=True)
    yd = F.conv_transpose2d(xd, Fd, bd, stride=s)
    (yd * dy.double()[:, 0].permute(0, 3, 1, 2)).sum().backward()

    def close(got, want, tol):
        return float((got.double() - want).abs().max()) <= tol * float(want.abs().max())

    assert close(dx[:, 0].permute(0, 3, 1, 2), xd.grad, 2.0 ** -8)
    assert close(tb.dF.cpu(), Fd.grad.permute(2, 3, 1, 0), 2e-5)
    assert close(tb.dbias.cpu(), bd.grad, 1e-5)
    tb.close()

"""CPU restatement of the reference's dense network behind the voxel grid — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module; the product
(lisec_b200/) never does. PARITY UNPINNED: the arithmetic of these layers lives in TensorFlow/Keras (version unpinned
by the reference, not installable here) and the reference ships no golden outputs or weights; this file restates the
published Keras layer semantics with torch CPU ops.

Follows reference model_training.py:
  addConv3DLayer  :191-196   ZeroPadding3D(p) -> Conv3D(64, 3, strides=s, 'valid', bias) -> BatchNormalization()
                              -> Dense(64, relu, use_bias=False) via addDenseLayer :178-186
  createModel     :236-238   three blocks, strides (2,1,1) (1,1,1) (2,1,1), pads (1,1,1) (0,1,1) (1,1,1)
                  :242-243   Permute((2,3,4,1)) + Reshape -> (nx, ny, 64 * 1)
  addConv2DLayer  :201-208   ZeroPadding2D(p) -> Conv2D(k3, stride, bias) -> BatchNormalization() -> ReLU
  addRPNConvLayer :211-215   one stride-2 layer then q stride-1 layers
  createModel     :245-254   blocks (128,q=3) (128,q=5) (256,q=5); Conv2DTranspose(256, k3 s1 / k2 s2 / k4 s4, 'same');
                              Concatenate; ClassificationLayer (2) and RegressionLayer (14): 1x1, linear
Keras defaults: BatchNormalization(axis=-1, epsilon=1e-3) in inference mode; channels_last everywhere.
arch.post_dense (lisec_b200.weights.Architecture) switches on the lines the reference has commented out — :205, Conv2D ->
BatchNormalization -> Dense(cout, relu, no bias) instead of -> ReLU — and the 128-channel grid: the graph model.png shows.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from lisec_b200.weights import CURRENT, Architecture, conv2d_post_dense, conv3d_blocks, rpn_blocks

BN_EPS = 1e-3


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _bn(x, pack, name, dtype):
    """Keras inference BatchNormalization on a channels-first tensor: y = (x - mean) * gamma / sqrt(var + eps) + beta."""
    g, b, m, v = (_t(pack["%s/%s" % (name, f)], dtype) for f in ("gamma", "beta", "moving_mean", "moving_variance"))
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - m.view(shape)) * (g / torch.sqrt(v + BN_EPS)).view(shape) + b.view(shape)


def middle_forward(grid: np.ndarray, pack: dict, dtype=torch.float64, arch: Architecture = CURRENT) -> torch.Tensor:
    """grid [N, nz, nx, ny, C3] (the VFE output) -> [N, 64, nx, ny] (channels-first) after the three Conv3D blocks."""
    x = _t(grid, dtype).permute(0, 4, 1, 2, 3)  # N C D(z) H(x) W(y)
    for conv, bn, dense, stride, pad in conv3d_blocks(arch):
        w = _t(pack[conv + "/kernel"], dtype).permute(4, 3, 0, 1, 2)  # (kd,kh,kw,cin,cout) -> (cout,cin,kd,kh,kw)
        x = F.conv3d(x, w, _t(pack[conv + "/bias"], dtype), stride=stride, padding=pad)
        x = _bn(x, pack, bn, dtype)
        wd = _t(pack[dense + "/kernel"], dtype)  # (cin, cout), no bias, relu
        x = torch.relu(torch.einsum("ncdhw,ck->nkdhw", x, wd))
    if x.shape[2] != 1:
        raise ValueError("the Conv3D stack must collapse z to 1 (nz = 8), got %d" % x.shape[2])
    # Permute((2,3,4,1)) + Reshape: [N,1,nx,ny,64] -> [N,nx,ny,64*1]
    return x[:, :, 0]


def rpn_forward(x: torch.Tensor, pack: dict, dtype=torch.float64, arch: Architecture = CURRENT):
    """[N, 64, nx, ny] -> prob [N, nx/2, ny/2, 2], regress [N, nx/2, ny/2, 14] (channels-last, like model.predict)."""
    ups = []
    post = conv2d_post_dense(arch)
    for convs, (tname, k, s, _) in rpn_blocks():
        for conv, bn, _, _, stride in convs:
            w = _t(pack[conv + "/kernel"], dtype).permute(3, 2, 0, 1)  # (kh,kw,cin,cout) -> (cout,cin,kh,kw)
            x = F.conv2d(x, w, _t(pack[conv + "/bias"], dtype), stride=stride, padding=1)
            x = _bn(x, pack, bn, dtype)
            if conv in post:  # addDenseLayer(layer, cout, 'relu') (:205)
                x = torch.einsum("nchw,ck->nkhw", x, _t(pack[post[conv] + "/kernel"], dtype))
            x = torch.relu(x)
        # Keras Conv2DTranspose kernel (kh,kw,cout,cin); 'same' => output = input * stride: k3 s1 crops 1, k == s crops 0
        wt = _t(pack[tname + "/kernel"], dtype).permute(3, 2, 0, 1)  # -> torch (cin, cout, kh, kw)
        ups.append(F.conv_transpose2d(x, wt, _t(pack[tname + "/bias"], dtype), stride=s, padding=(k - s) // 2))
    cat = torch.cat(ups, dim=1)
    outs = []
    for head in ("ClassificationLayer", "RegressionLayer"):
        w = _t(pack[head + "/kernel"], dtype).permute(3, 2, 0, 1)
        outs.append(F.conv2d(cat, w, _t(pack[head + "/bias"], dtype)).permute(0, 2, 3, 1).contiguous())
    return outs[0], outs[1]


def network_forward(grid: np.ndarray, pack: dict, dtype=torch.float64, arch: Architecture = CURRENT):
    """Everything behind MaxPoolingVFELayer(combine=True) (:235): numpy prob, regress."""
    with torch.no_grad():
        p, r = rpn_forward(middle_forward(grid, pack, dtype, arch), pack, dtype, arch)
    return p.numpy(), r.numpy()

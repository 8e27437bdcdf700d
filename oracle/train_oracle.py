"""CPU restatement of the reference's TRAINING step — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module.
PARITY UNPINNED: every number here comes out of TensorFlow/Keras in the reference (version unpinned, not installable
here; no golden outputs, no weights). This file restates the published semantics with torch CPU float64 + autograd.

Follows reference model_training.py:
  createModel          :222-257   the whole graph on the DENSE input [N, nz, nx, ny, T, 6], layer for layer — including
                                  the VFE stack on all nz*nx*ny*T slots, which is what Keras differentiates
  addVFELayer/addFCN   :155-174   Dense (no bias) -> BatchNormalization -> ReLU; MaxPoolingVFELayer (K.max over axis -2,
                                  :44-56) -> RepeatLayer (:32-40) -> Concatenate([pooling, layer])
  addConv3DLayer       :191-196   ZeroPadding3D -> Conv3D(bias) -> BatchNormalization -> Dense(relu, no bias)
  addConv2DLayer       :201-208   ZeroPadding2D -> Conv2D(bias) -> BatchNormalization -> ReLU
  train                :295-299   SGD(lr=0.01, decay=1e-6, momentum=0.9, nesterov=True), loss=['mse', 'mse'],
                                  fit(batch_size=1): every step sees ONE sweep, BatchNormalization in training mode

Keras semantics restated:
  BatchNormalization(training)  y = gamma * (x - mean_B) / sqrt(var_B + 1e-3) + beta with the batch mean and the BIASED
                                batch variance over every axis but the last; moving_x <- 0.99 * moving_x + 0.01 * batch_x
                                (moving_variance from the biased variance: the non-fused path, which rank-5/6 inputs
                                always take; the fused rank-4 path of some TF versions uses the unbiased one — `unbiased_4d`)
  loss 'mse' x 2                mean over all elements of (y - t)^2 per output, summed over the two outputs
  SGD (optimizer_v2, resource_apply_keras_momentum)
                                lr_t = lr / (1 + decay * iterations);  accum <- momentum * accum - lr_t * grad;
                                var <- var + momentum * accum - lr_t * grad   (nesterov)
  max-pool gradient             K.max -> reduce_max: the gradient is split EQUALLY among tied maxima (TensorFlow's
                                _MinOrMaxGrad); torch.amax has the same rule. Ties are the norm here: the T - count pad
                                rows of a voxel are identical, and ReLU zeros tie across rows.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from lisec_b200.weights import conv3d_blocks, rpn_blocks

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
VFE_DENSE = ("dense", "dense_1", "dense_2")
VFE_BN = ("batch_normalization", "batch_normalization_1", "batch_normalization_2")


def to_params(pack: Dict[str, np.ndarray], dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """Keras-named arrays -> torch tensors; everything but the moving statistics requires grad."""
    out = {}
    for k, v in pack.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
        if "moving_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def _bn_train(x, p, name, stats, channels_last=True, unbiased_moving=False):
    """Training-mode BatchNormalization; records (batch mean, batch variance used for the moving update) in `stats`."""
    axes = tuple(range(x.dim() - 1)) if channels_last else (0,) + tuple(range(2, x.dim()))
    mean = x.mean(dim=axes)
    var = x.var(dim=axes, unbiased=False)
    n = x.numel() // mean.numel()
    stats[name] = (mean.detach(), (var * n / (n - 1) if unbiased_moving else var).detach())
    shape = [1] * x.dim()
    shape[-1 if channels_last else 1] = -1
    xhat = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    return xhat * p[name + "/gamma"].view(shape) + p[name + "/beta"].view(shape)


def forward_train(dense_input: torch.Tensor, p: Dict[str, torch.Tensor], unbiased_4d: bool = False):
    """dense_input [N, nz, nx, ny, T, 6] -> (prob [N,nx/2,ny/2,2], regress [N,nx/2,ny/2,14], batch statistics)."""
    stats: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
    T = dense_input.shape[-2]
    x = dense_input
    for i in range(2):  # addVFELayer x 2 (:231-232)
        h = torch.relu(_bn_train(x @ p[VFE_DENSE[i] + "/kernel"], p, VFE_BN[i], stats))
        pooled = h.amax(dim=-2, keepdim=True)  # MaxPoolingVFELayer(), pad rows included (nothing is masked)
        x = torch.cat([pooled.expand(*h.shape[:-2], T, h.shape[-1]), h], dim=-1)  # RepeatLayer + Concatenate
    h = torch.relu(_bn_train(x @ p[VFE_DENSE[2] + "/kernel"], p, VFE_BN[2], stats))  # addFCN(., 64, 64) (:233)
    grid = h.amax(dim=-2)  # MaxPoolingVFELayer(combine=True) (:235): [N, nz, nx, ny, 64]
    prob, reg = network_forward_train(grid, p, stats, unbiased_4d)
    return prob, reg, stats, grid


def _bf16_ste(x):
    """Round to bfloat16 in the forward pass, identity in the backward pass (straight-through): the float64 gradient of a
    network whose stored activations are bf16 — what a mixed-precision implementation computes up to float32 effects."""
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


def network_forward_train(grid: torch.Tensor, p: Dict[str, torch.Tensor], stats: dict, unbiased_4d: bool = False,
                          bf16_activations: bool = False, teacher: Dict[str, torch.Tensor] = None):
    """Everything behind the voxel grid [N, nz, nx, ny, 64] in training mode (:236-254): prob, regress.
    bf16_activations: every tensor the GPU chain stores in bf16 (convolution outputs, BN outputs, Dense outputs, the
    concat tensor) is rounded to bf16 on the way forward.
    teacher: {layer name: that layer group's OUTPUT as another implementation computed it} (channels-first for the
    convolution stages — keyed by the conv3d / conv2d / conv2d_transpose name —, channels-last for the two heads). Where
    given, the stage's output VALUE is replaced by the teacher's while the gradient still flows through the float64 stage
    (straight-through): autograd then yields the exact float64 gradient of every parameter AT THE OTHER IMPLEMENTATION'S
    OPERATING POINT — the same ReLU masks, batch statistics of the same inputs, the same loss residual. It separates a
    wrong backward pass (which would still disagree) from forward drift amplified by the network (which disappears)."""
    q = _bf16_ste if bf16_activations else (lambda t: t)
    teacher = teacher or {}

    def force(name, t):
        g = teacher.get(name)
        if g is None:
            return t
        if tuple(g.shape) != tuple(t.shape):
            raise ValueError("teacher[%s] has shape %s, the stage's output %s" % (name, tuple(g.shape), tuple(t.shape)))
        return g.to(t.dtype) + (t - t.detach())

    x = grid.permute(0, 4, 1, 2, 3)
    for conv, bn, dense, stride, pad in conv3d_blocks():
        w = p[conv + "/kernel"].permute(4, 3, 0, 1, 2)
        x = q(F.conv3d(x, w, p[conv + "/bias"], stride=stride, padding=pad))
        x = q(_bn_train(x, p, bn, stats, channels_last=False))
        x = force(conv, q(torch.relu(torch.einsum("ncdhw,ck->nkdhw", x, p[dense + "/kernel"]))))
    x = x[:, :, 0]
    ups = []
    for convs, (tname, k, s, _) in rpn_blocks():
        for conv, bn, _, _, stride in convs:
            w = p[conv + "/kernel"].permute(3, 2, 0, 1)
            x = q(F.conv2d(x, w, p[conv + "/bias"], stride=stride, padding=1))
            x = force(conv, q(torch.relu(_bn_train(x, p, bn, stats, channels_last=False, unbiased_moving=unbiased_4d))))
        wt = p[tname + "/kernel"].permute(3, 2, 0, 1)
        ups.append(force(tname, q(F.conv_transpose2d(x, wt, p[tname + "/bias"], stride=s, padding=(k - s) // 2))))
    cat = torch.cat(ups, dim=1)
    outs = []
    for head in ("ClassificationLayer", "RegressionLayer"):
        w = p[head + "/kernel"].permute(3, 2, 0, 1)
        outs.append(force(head, F.conv2d(cat, w, p[head + "/bias"]).permute(0, 2, 3, 1)))
    return outs[0], outs[1]


def loss_mse2(prob, regress, y_class, y_regress):
    """loss=['mse', 'mse'] (:296): the two mean-squared errors, summed."""
    return ((prob - y_class) ** 2).mean() + ((regress - y_regress) ** 2).mean()


def sgd_nesterov_update(var, accum, grad, iterations: int, lr=0.01, decay=1e-6, momentum=0.9, nesterov=True,
                        grad_scale=1.0):
    """One Keras SGD update (optimizer_v2 / resource_apply_keras_momentum) on arrays of any float dtype, evaluated in that
    dtype operation by operation: lr_t rounded once, then accum*momentum, lr_t*grad, their difference, ... — the order
    lisec_sgd_nesterov (lisec_b200/csrc/train.cu) follows. Returns (var, accum)."""
    dt = np.asarray(var).dtype.type
    lr_t = dt(lr / (1.0 + decay * iterations))
    m = dt(momentum)
    step = lr_t * (grad * dt(grad_scale))  # grad_scale = 1 / world_size after a summing all-reduce
    accum = accum * m - step
    var = var + (accum * m - step) if nesterov else var + accum
    return var, accum


def train_step(pack: Dict[str, np.ndarray], accum: Dict[str, np.ndarray], iterations: int, dense_input: np.ndarray,
               y_class: np.ndarray, y_regress: np.ndarray, lr=0.01, decay=1e-6, momentum=0.9, nesterov=True,
               unbiased_4d: bool = False):
    """One fit() step (:299) in float64: returns (loss, new pack incl. moving statistics, new accumulators, gradients)."""
    p = to_params(pack)
    prob, reg, stats, _ = forward_train(torch.from_numpy(dense_input).double(), p, unbiased_4d)
    loss = loss_mse2(prob, reg, torch.from_numpy(y_class).double(), torch.from_numpy(y_regress).double())
    names = [k for k, t in p.items() if t.requires_grad]
    grads = torch.autograd.grad(loss, [p[k] for k in names])
    new_pack, new_accum, gdict = {}, {}, {}
    for k, g in zip(names, grads):
        g = g.numpy()
        gdict[k] = g
        v, a = sgd_nesterov_update(np.asarray(pack[k], dtype=np.float64), np.asarray(accum.get(k, 0.0) * np.ones_like(g)),
                                   g, iterations, lr, decay, momentum, nesterov)
        new_pack[k], new_accum[k] = v, a
    for k, v in pack.items():
        if "moving_" in k:
            bn, field = k.rsplit("/", 1)
            mean, var = stats[bn]
            batch = (mean if field == "moving_mean" else var).numpy()
            new_pack[k] = np.asarray(v, dtype=np.float64) * BN_MOMENTUM + batch * (1.0 - BN_MOMENTUM)
    return float(loss.detach()), new_pack, new_accum, gdict


# ---- the VFE stack on ROWS WITH MULTIPLICITIES: the formulation the GPU training step will use ----------
# The dense input holds three classes of rows: kept points (weight 1), per non-full voxel ONE virtual pad row standing for
# its T - s identical zero rows (weight T - s), and ONE empty row standing for the 35 * n_empty rows of all empty voxels.
# forward_train_rows() evaluates the same graph as forward_train()'s VFE section on that compact representation; the only
# non-standard piece is the max over a voxel's rows, whose gradient must count a row's copies among the ties.
class _WeightedSegmentMax(torch.autograd.Function):
    """values [R, C], seg [R] (segment id per row, segments contiguous not required), weight [R] (copies per row) ->
    per-segment max [S, C]. Backward: TensorFlow's reduce_max rule on the EXPANDED rows — the gradient is split equally
    among tied copies — so compact row r receives weight_r * g / n_ties with n_ties = sum of the weights of tied rows."""

    @staticmethod
    def forward(ctx, values, seg, weight, n_seg):
        out = torch.full((n_seg, values.shape[1]), -float("inf"), dtype=values.dtype)
        out = out.scatter_reduce(0, seg.view(-1, 1).expand_as(values), values, reduce="amax", include_self=True)
        ctx.save_for_backward(values, seg, weight, out)
        return out

    @staticmethod
    def backward(ctx, g):
        values, seg, weight, out = ctx.saved_tensors
        tied = (values == out[seg]).to(values.dtype) * weight.view(-1, 1)
        n_ties = torch.zeros_like(out).index_add_(0, seg, tied)
        return tied * (g / n_ties)[seg], None, None, None


def forward_train_rows(feat_rows, row_voxel, counts, n_cells_total, T, p, stats=None):
    """feat_rows [R, 6] float64: the kept rows of all occupied voxels (voxel-major); row_voxel [R]; counts [V] = kept rows
    per voxel (= min(points, T)); n_cells_total = N * nz * nx * ny. Returns (voxel_out [V, 64], empty_out [64], stats)."""
    stats = {} if stats is None else stats
    V = len(counts)
    dt = feat_rows.dtype
    nonfull = torch.nonzero(counts < T).view(-1)
    n_empty = n_cells_total - V
    M = float(n_cells_total * T)
    # compact rows: [kept | virtual (one per non-full voxel) | empty (one)], weights and segment ids; the empty row is its
    # own segment V
    w = torch.cat([torch.ones(len(feat_rows), dtype=dt), (T - counts[nonfull]).to(dt), torch.tensor([float(T * n_empty)], dtype=dt)])
    seg = torch.cat([row_voxel, nonfull, torch.tensor([V])])
    x = torch.cat([feat_rows, torch.zeros((len(nonfull) + 1, feat_rows.shape[1]), dtype=dt)])

    def bn_relu(u, name):
        mean = (w.view(-1, 1) * u).sum(0) / M
        var = (w.view(-1, 1) * u * u).sum(0) / M - mean * mean
        stats[name] = (mean.detach(), var.detach())
        return torch.relu((u - mean) / torch.sqrt(var + BN_EPS) * p[name + "/gamma"] + p[name + "/beta"])

    for i in range(2):
        h = bn_relu(x @ p[VFE_DENSE[i] + "/kernel"], VFE_BN[i])
        pooled = _WeightedSegmentMax.apply(h, seg, w, V + 1)
        x = torch.cat([pooled[seg], h], dim=1)
    h = bn_relu(x @ p[VFE_DENSE[2] + "/kernel"], VFE_BN[2])
    out = _WeightedSegmentMax.apply(h, seg, w, V + 1)
    return out[:V], out[V], stats

"""The pieces of model_training.train()'s step (model_training.py:295-299) that exist so far: flat parameter / gradient /
accumulator buffers in the model's Keras weight order, the step's ONE collective (all-reduce of the 6 491 024-element
float32 gradient: NCCL over NVLink on the GPUs, any torch.distributed backend for the host logic) and the Keras
SGD-Nesterov update as one kernel over the flat buffer (lisec_sgd_nesterov), plus the 'mse' loss head
(lisec_mse_loss_grad). The backward pass is NOT built: compat.train() still raises.

Sweeps shard over ranks (lisec_b200/sharding.py), every rank holds the full model, gradients are summed across ranks and
the 1 / world_size lands inside the update kernel (grad_scale) — one pass over the flat buffers per step."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _native as N


def trainable_names(pack: Dict[str, np.ndarray]) -> List[str]:
    """Everything but BatchNormalization's moving statistics, in the pack's (= Keras creation) order."""
    return [k for k in pack if "moving_" not in k]


class FlatParameters:
    """var / accum / grad as three flat float32 tensors, with per-weight views under the Keras names. Offsets are padded
    to 4 elements so every weight starts 16-byte aligned."""

    def __init__(self, pack: Dict[str, np.ndarray], device="cuda"):
        self.names = trainable_names(pack)
        self.shapes = {k: tuple(np.shape(pack[k])) for k in self.names}
        self.offsets, off = {}, 0
        for k in self.names:
            self.offsets[k] = off
            off += (int(np.prod(self.shapes[k])) + 3) // 4 * 4
        self.numel_padded = off
        self.numel = sum(int(np.prod(s)) for s in self.shapes.values())
        self.device = torch.device(device)
        self.var = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.accum = torch.zeros_like(self.var)
        self.grad = torch.zeros_like(self.var)
        for k in self.names:
            self.view(self.var, k).copy_(torch.from_numpy(np.ascontiguousarray(pack[k], dtype=np.float32)))

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o, s = self.offsets[name], self.shapes[name]
        return flat[o:o + int(np.prod(s))].view(s)

    def to_pack(self) -> Dict[str, np.ndarray]:
        return {k: self.view(self.var, k).cpu().numpy() for k in self.names}


def allreduce_gradients(flat_grad: torch.Tensor, group=None, bucket_elems: int = 0):
    """The step's collective: SUM the flat gradient over the ranks (the 1 / world_size is applied by the update kernel).
    bucket_elems > 0 splits it into asynchronous chunks (what a backward pass would overlap with); returns the works."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    n = flat_grad.numel()
    step = n if bucket_elems <= 0 else bucket_elems
    return [dist.all_reduce(flat_grad[a:min(a + step, n)], op=dist.ReduceOp.SUM, group=group, async_op=True)
            for a in range(0, n, step)]


class SgdNesterov:
    """optimizers.SGD(lr=0.01, decay=1e-6, momentum=0.9, nesterov=True) (model_training.py:295) on FlatParameters."""

    def __init__(self, params: FlatParameters, lr=0.01, decay=1e-6, momentum=0.9, nesterov=True):
        if params.device.type != "cuda":
            raise RuntimeError("lisec_b200 has no CPU fallback: the update runs in lisec_sgd_nesterov on a CUDA device")
        self._lib = N.load()
        self.params, self.lr, self.decay, self.momentum, self.nesterov = params, lr, decay, momentum, nesterov
        self.iterations = 0

    def step(self, world_size: int = 1) -> None:
        p = self.params
        lr_t = np.float32(self.lr / (1.0 + self.decay * self.iterations))
        with torch.cuda.device(p.device):
            st = self._lib.lisec_sgd_nesterov(
                C.c_void_p(p.var.data_ptr()), C.c_void_p(p.accum.data_ptr()), C.c_void_p(p.grad.data_ptr()),
                p.numel_padded, C.c_float(1.0 / world_size), C.c_float(lr_t), C.c_float(self.momentum),
                1 if self.nesterov else 0, C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))
        self.iterations += 1


def mse_loss_grad(y: torch.Tensor, target: torch.Tensor, want_grad: bool = True):
    """One 'mse' term of loss=['mse','mse'] (:296): returns (loss as a 0-d float64 device tensor, d loss / d y or None)."""
    if y.dtype != torch.float32 or target.dtype != torch.float32 or y.shape != target.shape or not y.is_cuda:
        raise ValueError("y and target: cuda float32 tensors of one shape")
    lib = N.load()
    y, target = y.contiguous(), target.contiguous()
    dy = torch.empty_like(y) if want_grad else None
    acc = torch.zeros((), dtype=torch.float64, device=y.device)
    with torch.cuda.device(y.device):
        st = lib.lisec_mse_loss_grad(C.c_void_p(y.data_ptr()), C.c_void_p(target.data_ptr()), y.numel(),
                                     C.c_void_p(dy.data_ptr() if want_grad else 0), C.c_void_p(acc.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream(y.device).cuda_stream))
    if st != N.LISEC_OK:
        raise N.LisecError(st, lib.lisec_train_last_error().decode("utf-8", "replace"))
    return acc / y.numel(), dy


class _RefreshBatch:
    """While `recording` is a list, the operand refreshes of the stages (float32 master weights -> the plans' bf16 operand
    copies: a cast, or the data-gradient plans' flip + transpose) are RECORDED instead of launched: (src, dst, elements,
    kd, kh, kw, out_c, in_c, mode). DenseNetworkTrainer records its plain stages once and then refreshes all of them with
    one lisec_refresh_operands launch per step."""
    recording: Optional[list] = None


def _refresh_cast(lib, src: torch.Tensor, dst: torch.Tensor, stream) -> int:
    if _RefreshBatch.recording is not None:
        _RefreshBatch.recording.append((src.data_ptr(), dst.data_ptr(), src.numel(), 1, 1, 1, 0, 0, 0))
        return N.LISEC_OK
    return lib.lisec_cast_f32_to_bf16(C.c_void_p(src.data_ptr()), src.numel(), C.c_void_p(dst.data_ptr()), stream)


def _refresh_flip(lib, w: torch.Tensor, k, n_out: int, c_in: int, dst: torch.Tensor, stream) -> int:
    if _RefreshBatch.recording is not None:
        _RefreshBatch.recording.append((w.data_ptr(), dst.data_ptr(), w.numel(), k[0], k[1], k[2], n_out, c_in, 1))
        return N.LISEC_OK
    return lib.lisec_weights_flip_transpose(C.c_void_p(w.data_ptr()), k[0], k[1], k[2], n_out, c_in, C.c_void_p(dst.data_ptr()),
                                            stream)


class _GradSink:
    """Where the stage under construction should write its parameter gradients: {"dw" | "dwd" | "dgamma" | "dbeta":
    float32 CUDA tensor}. DenseNetworkTrainer fills it with views of the caller's flat gradient buffer (TrainStep), so that
    the weight-gradient plans and the BatchNormalization backward write straight into the buffer the all-reduce and the
    optimizer read — no per-tensor copies after the backward pass. Empty: every stage owns its gradient tensors."""
    current: Dict[str, torch.Tensor] = {}

    @staticmethod
    def take(key: str, shape, device) -> torch.Tensor:
        t = _GradSink.current.get(key)
        if t is None:
            return torch.empty(shape, dtype=torch.float32, device=device)
        if t.numel() != int(np.prod(shape)) or not t.is_contiguous() or t.dtype != torch.float32:
            raise ValueError("gradient sink %r: %s does not hold shape %s" % (key, tuple(t.shape), tuple(shape)))
        return t.view(shape)


class ConvWgrad:
    """dW of one convolution layer (lisec_conv_wgrad_plan_*, lisec_b200/csrc/wgrad.cu): x bf16 [B,D,H,W,C], dy bf16
    [B,OD,OH,OW,N] -> dw float32 [taps, N, C] (the forward plans' weight layout). Buffers are bound at construction."""

    def __init__(self, x: torch.Tensor, dy: torch.Tensor, k, stride_d: int, pad, tile=(16, 8), stride_hw: int = 1,
                 sink: Optional[str] = None):
        """sink: the _GradSink key this plan's dw should be taken from (None: a tensor of its own)."""
        if x.dtype != torch.bfloat16 or dy.dtype != torch.bfloat16 or not x.is_cuda or x.dim() != 5 or dy.dim() != 5:
            raise ValueError("x, dy: cuda bf16 [B, D, H, W, C]")
        self._lib = N.load()
        B, D, H, W, Cin = x.shape
        self.x, self.dy = x.contiguous(), dy.contiguous()
        self.desc = N.lisec_conv_desc(
            batch=B, in_d=D, in_h=H, in_w=W, in_c=Cin, kd=k[0], kh=k[1], kw=k[2], stride_d=stride_d, stride_hw=stride_hw,
            pad_d=pad[0], pad_h=pad[1], pad_w=pad[2], out_c=dy.shape[-1], n_tiles=1, shuffle=1, out_pitch=dy.shape[-1],
            out_ch_off=0, relu=0, out_dtype=N.LISEC_F32, tile_w=tile[0], tile_h=tile[1], m_tiles=1, in_dtype=N.LISEC_BF16,
            out_split=0, group_kh=0, reserved=0)
        taps = k[0] * k[1] * k[2]
        shape = (taps, dy.shape[-1], Cin)
        self.dw = _GradSink.take(sink, shape, x.device) if sink else torch.empty(shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = int(self._lib.lisec_conv_wgrad_workspace_bytes(C.byref(self.desc)))
            self.workspace = torch.empty(ws // 4, dtype=torch.float32, device=x.device)
            self.plan = C.c_void_p()
            st = self._lib.lisec_conv_wgrad_plan_create(C.byref(self.desc), C.c_void_p(self.x.data_ptr()),
                                                        C.c_void_p(self.dy.data_ptr()),
                                                        C.c_void_p(self.workspace.data_ptr()),
                                                        C.c_void_p(self.dw.data_ptr()), C.byref(self.plan))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_wgrad_last_error().decode("utf-8", "replace"))

    def run(self) -> torch.Tensor:
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_conv_wgrad_plan_run(self.plan, C.c_void_p(torch.cuda.current_stream(self.x.device).cuda_stream))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_wgrad_last_error().decode("utf-8", "replace"))
        return self.dw

    def close(self):
        if getattr(self, "plan", None):
            self._lib.lisec_conv_wgrad_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ConvDgrad:
    """dX of a stride-1 convolution layer: a FORWARD plan (lisec_conv_plan_*, the same TMA + tcgen05 kernels) over dy with
    the kernel flipped and its channel roles swapped (lisec_weights_flip_transpose) and pad' = k - 1 - pad.
    dy bf16 [B,OD,OH,OW,N] -> dx bf16 [B,D,H,W,C]; `w` is the layer's float32 master weight [taps, N, C]. Call
    refresh_weights() after every optimizer step (the operand is a bf16 copy), then run()."""

    def __init__(self, dy: torch.Tensor, w: torch.Tensor, k, pad, out_dtype=torch.bfloat16, tile=None):
        if dy.dtype != torch.bfloat16 or not dy.is_cuda or dy.dim() != 5 or w.dtype != torch.float32 or w.dim() != 3:
            raise ValueError("dy: cuda bf16 [B,OD,OH,OW,N]; w: cuda float32 [taps, N, C]")
        self._lib = N.load()
        B, OD, OH, OW, Nout = dy.shape
        taps, n2, Cin = w.shape
        if n2 != Nout or taps != k[0] * k[1] * k[2]:
            raise ValueError("w does not match dy / k")
        self.dy, self.w, self.k = dy.contiguous(), w, tuple(k)
        p2 = tuple(k[i] - 1 - pad[i] for i in range(3))
        D, H, W = (OD + 2 * p2[0] - k[0] + 1, OH + 2 * p2[1] - k[1] + 1, OW + 2 * p2[2] - k[2] + 1)
        self.dx = torch.empty((B, D, H, W, Cin), dtype=out_dtype, device=dy.device)
        self.wt = torch.empty((taps, Cin, Nout), dtype=torch.bfloat16, device=dy.device)
        self.scale = torch.ones(Cin, dtype=torch.float32, device=dy.device)
        self.shift = torch.zeros(Cin, dtype=torch.float32, device=dy.device)
        self.refresh_weights()
        self.plan = C.c_void_p()
        # the halo plans (one input box per (kd, 64 channels) serves all nine taps, two M-tiles) where they apply, else a
        # plain one-M-tile plan
        cands = [((8, 16), 2, 2)] if (k[1] == 3 and k[2] == 3 and Cin <= 128 and tile is None) else []
        cands.append((tile or ((16, 8) if W >= 16 else (8, 16)), 1, 0))
        st = N.LISEC_ERR_BAD_CONFIG
        for tl, mt, gk in cands:
            self.desc = N.lisec_conv_desc(
                batch=B, in_d=OD, in_h=OH, in_w=OW, in_c=Nout, kd=k[0], kh=k[1], kw=k[2], stride_d=1, stride_hw=1,
                pad_d=p2[0], pad_h=p2[1], pad_w=p2[2], out_c=Cin, n_tiles=1, shuffle=1, out_pitch=Cin, out_ch_off=0,
                relu=0, out_dtype=N.LISEC_BF16 if out_dtype == torch.bfloat16 else N.LISEC_F32, tile_w=tl[0],
                tile_h=tl[1], m_tiles=mt, in_dtype=N.LISEC_BF16, out_split=0, group_kh=gk, reserved=0)
            with torch.cuda.device(dy.device):
                st = self._lib.lisec_conv_plan_create(C.byref(self.desc), C.c_void_p(self.dy.data_ptr()),
                                                      C.c_void_p(self.wt.data_ptr()), C.c_void_p(self.scale.data_ptr()),
                                                      C.c_void_p(self.shift.data_ptr()), C.c_void_p(self.dx.data_ptr()),
                                                      C.byref(self.plan))
            if st != N.LISEC_ERR_BAD_CONFIG:
                break
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dy.device).cuda_stream)

    def refresh_weights(self) -> None:
        taps, Nout, Cin = self.w.shape
        with torch.cuda.device(self.dy.device):
            st = _refresh_flip(self._lib, self.w, self.k, Nout, Cin, self.wt, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))

    def run(self) -> torch.Tensor:
        with torch.cuda.device(self.dy.device):
            st = self._lib.lisec_conv_plan_run(self.plan, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
        return self.dx

    def close(self):
        if getattr(self, "plan", None):
            self._lib.lisec_conv_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchNormTrain:
    """Training-mode BatchNormalization (+ optional ReLU) on a channels-last bf16 activation (lisec_bn_train_*): forward()
    leaves y and the batch statistics, backward(dy) returns dx and fills dgamma / dbeta. gamma, beta, moving_mean,
    moving_var: float32 device vectors owned by the caller (views into FlatParameters / the model's state)."""

    def __init__(self, x: torch.Tensor, gamma, beta, moving_mean=None, moving_var=None, relu=True, eps=1e-3, momentum=0.99,
                 unbiased_moving: bool = False):
        """unbiased_moving: update moving_variance with the Bessel-corrected batch variance var * P / (P - 1), as Keras's
        FUSED BatchNormalization does (the path rank-4 inputs take; rank-5/6 inputs take the non-fused one with the biased
        variance in the TensorFlow versions of the reference's era). The normalisation itself always uses the biased one."""
        if x.dtype != torch.bfloat16 or not x.is_cuda:
            raise ValueError("x: cuda bf16, channels last")
        self._lib = N.load()
        self.x, self.relu, self.eps, self.momentum = x, bool(relu), eps, momentum
        self.unbiased_moving = bool(unbiased_moving)
        self.C = x.shape[-1]
        self.P = x.numel() // self.C
        self.gamma, self.beta, self.moving_mean, self.moving_var = gamma, beta, moving_mean, moving_var
        dev = x.device
        ws = int(self._lib.lisec_bn_workspace_bytes(self.P, self.C))
        if ws < 0:
            raise ValueError("unsupported channel count %d" % self.C)
        self.workspace = torch.empty(ws // 8, dtype=torch.float64, device=dev)
        vec = lambda: torch.empty(self.C, dtype=torch.float32, device=dev)  # noqa: E731
        self.mean, self.invstd, self.scale, self.shift = vec(), vec(), vec(), vec()
        self.dgamma, self.dbeta = _GradSink.take("dgamma", (self.C,), dev), _GradSink.take("dbeta", (self.C,), dev)
        self._mg, self._mgx = vec(), vec()
        self.y = torch.empty_like(x)
        self.dx = torch.empty_like(x)

    def _p(self, t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.x.device).cuda_stream)

    def forward(self) -> torch.Tensor:
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_bn_train_forward(
                self._p(self.x), self.P, self.C, self._p(self.gamma), self._p(self.beta), C.c_float(self.eps),
                C.c_float(self.momentum), self._p(self.moving_mean), self._p(self.moving_var),
                int(self.relu) | (2 if self.unbiased_moving else 0),  # bit 1: Bessel-corrected moving variance (in the kernel)
                self._p(self.y), self._p(self.mean), self._p(self.invstd), self._p(self.scale), self._p(self.shift),
                self._p(self.workspace), self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_bn_last_error().decode("utf-8", "replace"))
        return self.y

    def backward(self, dy: torch.Tensor) -> torch.Tensor:
        if dy.dtype not in (torch.bfloat16, torch.float32) or dy.shape != self.x.shape:
            raise ValueError("dy: bf16 or float32 of x's shape")
        fn = self._lib.lisec_bn_train_backward if dy.dtype == torch.bfloat16 else self._lib.lisec_bn_train_backward_f32
        with torch.cuda.device(self.x.device):
            st = fn(
                self._p(self.x), self._p(dy.contiguous()), self._p(self.y), self.P, self.C, self._p(self.gamma),
                self._p(self.mean), self._p(self.invstd), int(self.relu), self._p(self.dx), self._p(self.dgamma),
                self._p(self.dbeta), self._p(self._mg), self._p(self._mgx), self._p(self.workspace), self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_bn_last_error().decode("utf-8", "replace"))
        return self.dx


class ConvBnReluTrain:
    """One addConv2DLayer / Conv3D + BN stage in TRAINING mode (model_training.py:191-208 under fit): convolution with
    bias (tensor-core plan on a bf16 copy of the float32 master weights) -> training-mode BatchNormalization -> ReLU,
    and its backward: BN backward -> weight gradient (conv_wgrad_kernel), bias gradient (column sums), data gradient
    (forward plan on dz with flipped, transposed weights; zero-dilated dz for the stride-2 layers). Master weights: w float32
    [taps, N, C], bias / gamma / beta float32 [N] (views into FlatParameters in a full model)."""

    def __init__(self, x: torch.Tensor, w, bias, gamma, beta, k, pad, relu=True, moving_mean=None, moving_var=None,
                 need_dx=True, tile=None, stride_hw: int = 1, grad_dtype=torch.bfloat16):
        self._lib = N.load()
        B, D, H, W, Cin = x.shape
        taps, Nout, c2 = w.shape
        if c2 != Cin or taps != k[0] * k[1] * k[2]:
            raise ValueError("w does not match x / k")
        dev = x.device
        self.x, self.w, self.bias, self.k, self.pad = x, w, bias, tuple(k), tuple(pad)
        s = stride_hw
        OD, OH, OW = D + 2 * pad[0] - k[0] + 1, (H + 2 * pad[1] - k[1]) // s + 1, (W + 2 * pad[2] - k[2]) // s + 1
        self.w16 = torch.empty((taps, Nout, Cin), dtype=torch.bfloat16, device=dev)
        self.z = torch.empty((B, OD, OH, OW, Nout), dtype=torch.bfloat16, device=dev)
        self.ones = torch.ones(Nout, dtype=torch.float32, device=dev)
        if tile is None:
            tile = (16, 8) if OW >= 16 else (8, 16)
        self.desc = N.lisec_conv_desc(
            batch=B, in_d=D, in_h=H, in_w=W, in_c=Cin, kd=k[0], kh=k[1], kw=k[2], stride_d=1, stride_hw=s, pad_d=pad[0],
            pad_h=pad[1], pad_w=pad[2], out_c=Nout, n_tiles=1, shuffle=1, out_pitch=Nout, out_ch_off=0, relu=0,
            out_dtype=N.LISEC_BF16, tile_w=tile[0], tile_h=tile[1], m_tiles=1, in_dtype=N.LISEC_BF16, out_split=0,
            group_kh=0, reserved=0)  # (the bias in front of a training-mode BN never changes: its gradient is zero)
        self.refresh_weights()
        self.plan = C.c_void_p()
        with torch.cuda.device(dev):
            st = self._lib.lisec_conv_plan_create(C.byref(self.desc), C.c_void_p(x.data_ptr()),
                                                  C.c_void_p(self.w16.data_ptr()), C.c_void_p(self.ones.data_ptr()),
                                                  C.c_void_p(bias.data_ptr()), C.c_void_p(self.z.data_ptr()),
                                                  C.byref(self.plan))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
        self.bn = BatchNormTrain(self.z, gamma, beta, moving_mean, moving_var, relu=relu,
                                 unbiased_moving=(k[0] == 1))  # Conv2D stages: rank-4 input, Keras's fused path
        self.wgrad = ConvWgrad(x, self.bn.dx, k, 1, pad, tile=tile, stride_hw=s, sink="dw")  # dz lands in bn.dx
        self.dgrad = None
        if need_dx:
            self.dgrad = (ConvDgrad(self.bn.dx, w, k, pad, out_dtype=grad_dtype) if s == 1 else
                          ConvDgradStrided(self.bn.dx, w, k, pad, 1, s, (D, H, W), out_dtype=grad_dtype))
        self.dbias = torch.zeros(Nout, dtype=torch.float32, device=dev)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.x.device).cuda_stream)

    def refresh_weights(self) -> None:
        """After an optimizer step: the bf16 operand copies of the master weights."""
        with torch.cuda.device(self.x.device):
            st = _refresh_cast(self._lib, self.w, self.w16, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))
        if getattr(self, "dgrad", None) is not None:
            self.dgrad.refresh_weights()

    def forward(self) -> torch.Tensor:
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_conv_plan_run(self.plan, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
        return self.bn.forward()

    def backward(self, dy: torch.Tensor):
        """dy: gradient with respect to this stage's output. Returns dx (or None); dw / dbias / bn.dgamma / bn.dbeta hold
        the parameter gradients afterwards."""
        self.bn.backward(dy)  # dz lands in bn.dx, which the weight- and data-gradient plans read
        self.dw = self.wgrad.run()
        # dbias stays zero: a bias in front of a training-mode BatchNormalization has no effect on anything behind it (the
        # batch mean removes it), and its gradient, the per-channel sum of dz, is zero by construction of the BN backward —
        # what a float32 framework computes there is rounding noise (~2^-24 |dz| sqrt(P)); a sum over this chain's
        # bf16-stored dz would be noise 2^15 times larger. Exact zero is the closer of the two, and saves a pass over dz.
        return self.dgrad.run() if self.dgrad is not None else None

    def close(self):
        if getattr(self, "plan", None):
            self._lib.lisec_conv_plan_destroy(self.plan)
            self.plan = None
        self.wgrad.close()
        if self.dgrad is not None:
            self.dgrad.close()


class ConvDgradStrided:
    """dX of a STRIDED convolution layer (stride_d and / or stride_hw = 2): dy is zero-dilated (lisec_dilate) and goes
    through the stride-1 data-gradient plan. in_dhw = (D, H, W) of the layer's input. First version: the plan multiplies
    the inserted zeros too (2x / 4x the MACs); the phase-decomposed form is the optimisation to come."""

    def __init__(self, dy: torch.Tensor, w: torch.Tensor, k, pad, stride_d: int, stride_hw: int, in_dhw,
                 out_dtype=torch.bfloat16):
        self._lib = N.load()
        B, OD, OH, OW, Nout = dy.shape
        sizes = []
        for n_out, n_in, kk, pp, ss in ((OD, in_dhw[0], k[0], pad[0], stride_d), (OH, in_dhw[1], k[1], pad[1], stride_hw),
                                        (OW, in_dhw[2], k[2], pad[2], stride_hw)):
            if (n_in + 2 * pp - kk) // ss + 1 != n_out:
                raise ValueError("dy does not match the layer's geometry")
            sizes.append((n_out - 1) * ss + 1 + (n_in + 2 * pp - kk) % ss)
        self.dy, self.sd, self.s = dy.contiguous(), stride_d, stride_hw
        self.dilated = torch.zeros((B, sizes[0], sizes[1], sizes[2], Nout), dtype=torch.bfloat16, device=dy.device)
        self.inner = ConvDgrad(self.dilated, w, k, pad, out_dtype=out_dtype)
        if tuple(self.inner.dx.shape[1:4]) != tuple(in_dhw):
            raise RuntimeError("dilated data gradient has shape %s, expected %s" % (tuple(self.inner.dx.shape[1:4]), tuple(in_dhw)))
        self.dx = self.inner.dx

    def refresh_weights(self) -> None:
        self.inner.refresh_weights()

    def run(self) -> torch.Tensor:
        B, OD, OH, OW, Nout = self.dy.shape
        _, D2, H2, W2, _ = self.dilated.shape
        with torch.cuda.device(self.dy.device):
            st = self._lib.lisec_dilate(C.c_void_p(self.dy.data_ptr()), B, OD, OH, OW, Nout, self.sd, self.s, D2, H2, W2,
                                        C.c_void_p(self.dilated.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream(self.dy.device).cuda_stream))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))
        return self.inner.run()

    def close(self):
        self.inner.close()


def add_(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a += b on bf16 device tensors (lisec_add_bf16): gradient accumulation where a tensor has two consumers."""
    if a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape:
        return a.add_(b)  # float32 gradient tensors: torch's elementwise add (plumbing between stages)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or a.shape != b.shape or not a.is_cuda or not a.is_contiguous():
        raise ValueError("a, b: contiguous cuda bf16 tensors of one shape")
    lib = N.load()
    with torch.cuda.device(a.device):
        st = lib.lisec_add_bf16(C.c_void_p(a.data_ptr()), C.c_void_p(b.contiguous().data_ptr()), a.numel(),
                                C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
    if st != N.LISEC_OK:
        raise N.LisecError(st, lib.lisec_train_last_error().decode("utf-8", "replace"))
    return a


def relu_backward(dy: torch.Tensor, y: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dy masked by y > 0 (bf16): the ReLU behind the Dense of a Conv3D block (model_training.py:195)."""
    if dy.dtype not in (torch.bfloat16, torch.float32) or y.dtype != torch.bfloat16 or dy.shape != y.shape or not dy.is_cuda:
        raise ValueError("dy (bf16 or float32), y (bf16): cuda tensors of one shape")
    lib = N.load()
    out = torch.empty_like(y) if out is None else out
    fn = lib.lisec_relu_backward if dy.dtype == torch.bfloat16 else lib.lisec_relu_backward_f32
    with torch.cuda.device(dy.device):
        st = fn(C.c_void_p(dy.contiguous().data_ptr()), C.c_void_p(y.contiguous().data_ptr()), dy.numel(),
                C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(dy.device).cuda_stream))
    if st != N.LISEC_OK:
        raise N.LisecError(st, lib.lisec_train_last_error().decode("utf-8", "replace"))
    return out


class Conv3dBlockTrain:
    """One addConv3DLayer block in TRAINING mode (model_training.py:191-196): ZeroPadding3D -> Conv3D(bias, stride (sd,1,1))
    -> BatchNormalization (batch statistics) -> Dense(64, no bias) -> ReLU, forward and backward, from the pieces above.
    Master weights: w [27, N, C], bias [N], gamma / beta [N], wd [1, N2, N] (the Dense kernel as a 1x1 convolution in the
    plans' [tap][out][in] layout), all float32 on the device."""

    def __init__(self, x: torch.Tensor, w, bias, gamma, beta, wd, k, pad, stride_d=1, moving_mean=None, moving_var=None,
                 need_dx=True, grad_dtype=torch.bfloat16):
        self._lib = N.load()
        B, D, H, W, Cin = x.shape
        taps, Nout, _ = w.shape
        N2 = wd.shape[1]
        dev = x.device
        self.x, self.w, self.wd, self.bias = x, w, wd, bias
        OD = (D + 2 * pad[0] - k[0]) // stride_d + 1
        OH, OW = H + 2 * pad[1] - k[1] + 1, W + 2 * pad[2] - k[2] + 1
        tile = (16, 8) if OW >= 16 else (8, 16)
        self.w16 = torch.empty((taps, Nout, Cin), dtype=torch.bfloat16, device=dev)
        self.wd16 = torch.empty((1, N2, Nout), dtype=torch.bfloat16, device=dev)
        self.z = torch.empty((B, OD, OH, OW, Nout), dtype=torch.bfloat16, device=dev)
        self.y = torch.empty((B, OD, OH, OW, N2), dtype=torch.bfloat16, device=dev)
        self.dv = torch.empty_like(self.y)
        self.ones = torch.ones(max(Nout, N2), dtype=torch.float32, device=dev)
        self.zeros = torch.zeros(N2, dtype=torch.float32, device=dev)
        self.bn = BatchNormTrain(self.z, gamma, beta, moving_mean, moving_var, relu=False)

        def plan(src, wt, shift, dst, kk, pp, sd, cin, cout, relu):
            desc = N.lisec_conv_desc(
                batch=B, in_d=src.shape[1], in_h=src.shape[2], in_w=src.shape[3], in_c=cin, kd=kk[0], kh=kk[1], kw=kk[2],
                stride_d=sd, stride_hw=1, pad_d=pp[0], pad_h=pp[1], pad_w=pp[2], out_c=cout, n_tiles=1, shuffle=1,
                out_pitch=cout, out_ch_off=0, relu=relu, out_dtype=N.LISEC_BF16, tile_w=tile[0], tile_h=tile[1], m_tiles=1,
                in_dtype=N.LISEC_BF16, out_split=0, group_kh=0, reserved=0)  # (the bias in front of a training-mode BN never changes: its gradient is zero)
            h = C.c_void_p()
            with torch.cuda.device(dev):
                st = self._lib.lisec_conv_plan_create(C.byref(desc), C.c_void_p(src.data_ptr()), C.c_void_p(wt.data_ptr()),
                                                      C.c_void_p(self.ones.data_ptr()), C.c_void_p(shift.data_ptr()),
                                                      C.c_void_p(dst.data_ptr()), C.byref(h))
            if st != N.LISEC_OK:
                raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
            return h

        self._dgrads = []
        self.refresh_weights()
        self.conv_plan = plan(x, self.w16, bias, self.z, k, pad, stride_d, Cin, Nout, 0)
        self.dense_plan = plan(self.bn.y, self.wd16, self.zeros, self.y, (1, 1, 1), (0, 0, 0), 1, Nout, N2, 1)
        self.dense_wgrad = ConvWgrad(self.bn.y, self.dv, (1, 1, 1), 1, (0, 0, 0), tile=tile, sink="dwd")
        self.dense_dgrad = ConvDgrad(self.dv, wd, (1, 1, 1), (0, 0, 0), out_dtype=grad_dtype)  # du: gradient at the BN output
        self.conv_wgrad = ConvWgrad(x, self.bn.dx, k, stride_d, pad, tile=tile, sink="dw")    # dz lands in bn.dx
        self.conv_dgrad = None
        if need_dx:
            self.conv_dgrad = (ConvDgrad(self.bn.dx, w, k, pad, out_dtype=grad_dtype) if stride_d == 1 else
                               ConvDgradStrided(self.bn.dx, w, k, pad, stride_d, 1, (D, H, W), out_dtype=grad_dtype))
        self._dgrads = [self.dense_dgrad] + ([self.conv_dgrad] if self.conv_dgrad is not None else [])
        self.dbias = torch.zeros(Nout, dtype=torch.float32, device=dev)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.x.device).cuda_stream)

    def refresh_weights(self) -> None:
        with torch.cuda.device(self.x.device):
            for src, dst in ((self.w, self.w16), (self.wd, self.wd16)):
                st = _refresh_cast(self._lib, src, dst, self._stream())
                if st != N.LISEC_OK:
                    raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))
        for dg in self._dgrads:
            dg.refresh_weights()

    def _run(self, plan):
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_conv_plan_run(plan, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))

    def forward(self) -> torch.Tensor:
        self._run(self.conv_plan)
        self.bn.forward()
        self._run(self.dense_plan)
        return self.y

    def backward(self, dy: torch.Tensor):
        relu_backward(dy, self.y, out=self.dv)
        self.dwd = self.dense_wgrad.run()
        du = self.dense_dgrad.run()
        self.bn.backward(du)
        self.dw = self.conv_wgrad.run()
        # (dbias stays zero: the bias sits in front of a training-mode BatchNormalization — see ConvBnReluTrain.backward)
        return self.conv_dgrad.run() if self.conv_dgrad is not None else None

    def close(self):
        for h in ("conv_plan", "dense_plan"):
            if getattr(self, h, None):
                self._lib.lisec_conv_plan_destroy(getattr(self, h))
                setattr(self, h, None)
        for o in (self.dense_wgrad, self.conv_wgrad, self.dense_dgrad, self.conv_dgrad):
            if o is not None:
                o.close()


class HeadsTrain:
    """ClassificationLayer + RegressionLayer in training (model_training.py:253-254): one 1x1 convolution 768 -> 2 + 14
    with bias, float32 outputs; its backward from the float32 loss gradient dy [B,1,H,W,16]: dy is widened to 64 bf16
    channels (lisec_pad_channels_bf16), the weight gradient is conv_wgrad_kernel's (rows 0..15 of its 64), the bias
    gradient the column sums, the data gradient a plan with three 256-column N-tiles. w: float32 [1, 16, 768] ([tap][out][in]),
    bias float32 [16]; x: bf16 [B,1,H,W,768] (the concat tensor)."""

    def __init__(self, x: torch.Tensor, w, bias, dense_slices: bool = False):
        self._lib = N.load()
        B, D, H, W, Cin = x.shape
        dev = x.device
        self.x, self.w, self.bias = x, w, bias
        self.w16 = torch.empty((1, 16, Cin), dtype=torch.bfloat16, device=dev)
        self.y = torch.empty((B, D, H, W, 16), dtype=torch.float32, device=dev)
        self.dy64 = torch.empty((B, D, H, W, 64), dtype=torch.bfloat16, device=dev)
        self.w64 = torch.zeros((1, 64, Cin), dtype=torch.float32, device=dev)  # rows 16.. stay zero
        self.ones = torch.ones(Cin, dtype=torch.float32, device=dev)
        self.zeros = torch.zeros(Cin, dtype=torch.float32, device=dev)
        tile = (16, 8) if W >= 16 else (8, 16)

        def plan(src, wt, scale, shift, dst, cin, cout, n_tiles, out_dtype):
            desc = N.lisec_conv_desc(
                batch=B, in_d=D, in_h=H, in_w=W, in_c=cin, kd=1, kh=1, kw=1, stride_d=1, stride_hw=1, pad_d=0, pad_h=0,
                pad_w=0, out_c=cout, n_tiles=n_tiles, shuffle=1, out_pitch=cout * n_tiles, out_ch_off=0, relu=0,
                out_dtype=out_dtype, tile_w=tile[0], tile_h=tile[1], m_tiles=1, in_dtype=N.LISEC_BF16, out_split=0,
                group_kh=0, reserved=1)  # bit 0: the bias is a trainable weight, rewritten between runs
            h = C.c_void_p()
            with torch.cuda.device(dev):
                st = self._lib.lisec_conv_plan_create(C.byref(desc), C.c_void_p(src.data_ptr()), C.c_void_p(wt.data_ptr()),
                                                      C.c_void_p(scale.data_ptr()), C.c_void_p(shift.data_ptr()),
                                                      C.c_void_p(dst.data_ptr()), C.byref(h))
            if st != N.LISEC_OK:
                raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
            return h

        self.fwd_plan = plan(x, self.w16, self.ones, bias, self.y, Cin, 16, 1, N.LISEC_F32)
        self.wgrad = ConvWgrad(x, self.dy64, (1, 1, 1), 1, (0, 0, 0), tile=tile)
        # data gradient: dx[p][ci] = sum_n dy64[p][n] * w64[n][ci]: weights [1][ci = 768][n = 64] as three N-tiles of 256
        # (the concat tensor is three 256-channel blocks with three different producers: with dense_slices the gradient
        # leaves as three dense [.., 256] tensors, one plan each, ready to be the dy of the three transposed stages)
        self.wt16 = torch.empty((1, Cin, 64), dtype=torch.bfloat16, device=dev)
        self.refresh_weights()
        if dense_slices:
            self.dx = [torch.empty((B, D, H, W, 256), dtype=torch.bfloat16, device=dev) for _ in range(Cin // 256)]
            self.dgrad_plans = [plan(self.dy64, self.wt16[:, 256 * i:256 * i + 256], self.ones, self.zeros, self.dx[i], 64, 256, 1,
                                     N.LISEC_BF16) for i in range(Cin // 256)]
        else:
            self.dx = torch.empty((B, D, H, W, Cin), dtype=torch.bfloat16, device=dev)
            self.dgrad_plans = [plan(self.dy64, self.wt16, self.ones, self.zeros, self.dx, 64, 256, Cin // 256, N.LISEC_BF16)]
        self.bn_ws = torch.empty(int(self._lib.lisec_bn_workspace_bytes(B * D * H * W, 64)) // 8, dtype=torch.float64, device=dev)
        self.dbias64 = torch.empty(64, dtype=torch.float32, device=dev)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.x.device).cuda_stream)

    def refresh_weights(self) -> None:
        self.w64[:, :16].copy_(self.w)
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_cast_f32_to_bf16(C.c_void_p(self.w.data_ptr()), self.w.numel(), C.c_void_p(self.w16.data_ptr()),
                                                  self._stream())
            if st == N.LISEC_OK:
                st = self._lib.lisec_weights_flip_transpose(C.c_void_p(self.w64.data_ptr()), 1, 1, 1, 64, self.w64.shape[2],
                                                            C.c_void_p(self.wt16.data_ptr()), self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))

    def _run(self, plan):
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_conv_plan_run(plan, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))

    def forward(self) -> torch.Tensor:
        self._run(self.fwd_plan)
        return self.y

    def backward(self, dy: torch.Tensor) -> torch.Tensor:
        """dy: float32 [B,1,H,W,16] (mse_loss_grad's output for the two heads, concatenated). Returns dx (bf16); dw
        [1,16,768] and dbias [16] hold the parameter gradients."""
        P = dy.numel() // 16
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_pad_channels_bf16(C.c_void_p(dy.contiguous().data_ptr()), P, 16, 64,
                                                   C.c_void_p(self.dy64.data_ptr()), self._stream())
            if st == N.LISEC_OK:
                st = self._lib.lisec_channel_sums(C.c_void_p(self.dy64.data_ptr()), P, 64, C.c_void_p(self.dbias64.data_ptr()),
                                                  C.c_void_p(self.bn_ws.data_ptr()), self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, (self._lib.lisec_train_last_error() or self._lib.lisec_bn_last_error()).decode("utf-8", "replace"))
        self.dw = self.wgrad.run()[:, :16]
        self.dbias = self.dbias64[:16]
        for pl in self.dgrad_plans:
            self._run(pl)
        return self.dx

    def close(self):
        for pl in [getattr(self, "fwd_plan", None)] + list(getattr(self, "dgrad_plans", [])):
            if pl:
                self._lib.lisec_conv_plan_destroy(pl)
        self.fwd_plan, self.dgrad_plans = None, []
        self.wgrad.close()


class ConvBiasTrain:
    """A convolution with bias and nothing behind it, in training: the k3 s1 'same' Conv2DTranspose of RPN block 1
    (model_training.py:247) is a 3x3 convolution with the kernel flipped (lisec_b200/network.py), so its training stage is
    this one. Forward writes into a channel slice of a wider buffer (out_pitch / out_ch_off: the concat tensor); backward
    takes a DENSE dy [B,D,H,W,N]. w: float32 [taps, N, C] in the plans' layout, bias float32 [N]."""

    def __init__(self, x: torch.Tensor, w, bias, k, pad, out: torch.Tensor, out_ch_off: int, dy: torch.Tensor, need_dx=True,
                 grad_dtype=torch.bfloat16):
        self._lib = N.load()
        B, D, H, W, Cin = x.shape
        taps, Nout, _ = w.shape
        dev = x.device
        self.x, self.w, self.bias, self.dy = x, w, bias, dy
        self.w16 = torch.empty((taps, Nout, Cin), dtype=torch.bfloat16, device=dev)
        self.ones = torch.ones(Nout, dtype=torch.float32, device=dev)
        tile = (16, 8) if W >= 16 else (8, 16)
        self.desc = N.lisec_conv_desc(
            batch=B, in_d=D, in_h=H, in_w=W, in_c=Cin, kd=k[0], kh=k[1], kw=k[2], stride_d=1, stride_hw=1, pad_d=pad[0],
            pad_h=pad[1], pad_w=pad[2], out_c=Nout, n_tiles=1, shuffle=1, out_pitch=out.shape[-1], out_ch_off=out_ch_off,
            relu=0, out_dtype=N.LISEC_BF16, tile_w=tile[0], tile_h=tile[1], m_tiles=1, in_dtype=N.LISEC_BF16, out_split=0,
            group_kh=0, reserved=1)  # bit 0: the bias is a trainable weight, rewritten between runs
        self.dgrad = ConvDgrad(dy, w, k, pad, out_dtype=grad_dtype) if need_dx else None
        self.refresh_weights()
        self.plan = C.c_void_p()
        with torch.cuda.device(dev):
            st = self._lib.lisec_conv_plan_create(C.byref(self.desc), C.c_void_p(x.data_ptr()), C.c_void_p(self.w16.data_ptr()),
                                                  C.c_void_p(self.ones.data_ptr()), C.c_void_p(bias.data_ptr()),
                                                  C.c_void_p(out.data_ptr()), C.byref(self.plan))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
        self.wgrad = ConvWgrad(x, dy, k, 1, pad, tile=tile, sink="dw")
        self.P = dy.numel() // Nout
        self.ws = torch.empty(int(self._lib.lisec_bn_workspace_bytes(self.P, Nout)) // 8, dtype=torch.float64, device=dev)
        self.dbias = _GradSink.take("dbias", (Nout,), dev)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.x.device).cuda_stream)

    def refresh_weights(self) -> None:
        with torch.cuda.device(self.x.device):
            st = _refresh_cast(self._lib, self.w, self.w16, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))
        if self.dgrad is not None:
            self.dgrad.refresh_weights()

    def forward(self) -> None:
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_conv_plan_run(self.plan, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))

    def backward(self):
        """From self.dy (filled by the stage behind). Returns dx; dw / dbias hold the parameter gradients."""
        self.dw = self.wgrad.run()
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_channel_sums(C.c_void_p(self.dy.data_ptr()), self.P, self.dy.shape[-1],
                                              C.c_void_p(self.dbias.data_ptr()), C.c_void_p(self.ws.data_ptr()), self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_bn_last_error().decode("utf-8", "replace"))
        return self.dgrad.run() if self.dgrad is not None else None

    def close(self):
        if getattr(self, "plan", None):
            self._lib.lisec_conv_plan_destroy(self.plan)
            self.plan = None
        self.wgrad.close()
        if self.dgrad is not None:
            self.dgrad.close()


class ConvTransposeBackward:
    """Backward of a Conv2DTranspose whose kernel equals its stride s (RPN blocks 2 and 3, model_training.py:249, 251):
    y[s*h + i, s*w + j] = x[h, w] . F[i, j] + b. With dy viewed as [B, H, s, W, s*Co] (the same memory: row s*h + i of the
    output holds, for each w, the s*Co values of columns s*w .. s*w + s - 1), dx is a convolution over that view with
    kh = s taps and no padding, and dF the matching weight gradient — both on the existing tensor-core kernels.
    F: float32 [s, s, Co, Ci] (the Keras layout); dy: bf16 dense [B, 1, s*H, s*W, Co]; x: bf16 [B, 1, H, W, Ci]."""

    def __init__(self, x: torch.Tensor, dy: torch.Tensor, F: torch.Tensor, s: int, grad_dtype=torch.bfloat16):
        self._lib = N.load()
        B, _, H, W, Ci = x.shape
        Co = dy.shape[-1]
        if tuple(dy.shape) != (B, 1, s * H, s * W, Co) or tuple(F.shape) != (s, s, Co, Ci) or s not in (2, 3, 4):
            raise ValueError("shapes do not describe a kernel = stride transposed convolution (s = 2, 3 or 4)")
        dev = x.device
        self.s, self.F, self.dy = s, F, dy
        self.dy_view = dy.view(B, H, s, W, s * Co)
        self.x_view = x.view(B, H, 1, W, Ci)
        tile = (128, 1)
        # dx: weights [tap = i][n = ci][c = (j, co)]
        self.w = torch.empty((s, Ci, s * Co), dtype=torch.float32, device=dev)
        self.w16 = torch.empty((s, Ci, s * Co), dtype=torch.bfloat16, device=dev)
        self.dx = torch.empty((B, H, 1, W, Ci), dtype=grad_dtype, device=dev)
        self.ones = torch.ones(Ci, dtype=torch.float32, device=dev)
        self.zeros = torch.zeros(Ci, dtype=torch.float32, device=dev)
        self.refresh_weights()
        desc = N.lisec_conv_desc(
            batch=B, in_d=H, in_h=s, in_w=W, in_c=s * Co, kd=1, kh=s, kw=1, stride_d=1, stride_hw=1, pad_d=0, pad_h=0,
            pad_w=0, out_c=Ci, n_tiles=1, shuffle=1, out_pitch=Ci, out_ch_off=0, relu=0,
            out_dtype=N.LISEC_BF16 if grad_dtype == torch.bfloat16 else N.LISEC_F32,
            tile_w=tile[0], tile_h=tile[1], m_tiles=1, in_dtype=N.LISEC_BF16, out_split=0, group_kh=0, reserved=0)
        self.plan = C.c_void_p()
        with torch.cuda.device(dev):
            st = self._lib.lisec_conv_plan_create(C.byref(desc), C.c_void_p(self.dy_view.data_ptr()),
                                                  C.c_void_p(self.w16.data_ptr()), C.c_void_p(self.ones.data_ptr()),
                                                  C.c_void_p(self.zeros.data_ptr()), C.c_void_p(self.dx.data_ptr()),
                                                  C.byref(self.plan))
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
        self.wgrad = ConvWgrad(self.dy_view, self.x_view, (1, s, 1), 1, (0, 0, 0), tile=tile)  # -> [i][ci][(j, co)]
        self.P = dy.numel() // Co
        self.ws = torch.empty(int(self._lib.lisec_bn_workspace_bytes(self.P, Co)) // 8, dtype=torch.float64, device=dev)
        self.dbias = torch.empty(Co, dtype=torch.float32, device=dev)
        self.dy = dy

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dy.device).cuda_stream)

    def refresh_weights(self) -> None:
        s = self.s
        # F[i, j, co, ci] -> w[i][ci][j*Co + co]: a permutation of a <= 1 MB tensor (torch indexing: plumbing, not arithmetic)
        self.w.copy_(self.F.permute(0, 3, 1, 2).reshape(s, self.F.shape[3], s * self.F.shape[2]))
        with torch.cuda.device(self.F.device):
            st = self._lib.lisec_cast_f32_to_bf16(C.c_void_p(self.w.data_ptr()), self.w.numel(), C.c_void_p(self.w16.data_ptr()),
                                                  self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))

    def backward(self):
        """Returns dx [B,1,H,W,Ci] (bf16); dF [s,s,Co,Ci] and dbias [Co] hold the parameter gradients."""
        s, Co, Ci = self.s, self.F.shape[2], self.F.shape[3]
        with torch.cuda.device(self.dy.device):
            st = self._lib.lisec_conv_plan_run(self.plan, self._stream())
            if st != N.LISEC_OK:
                raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))
            st = self._lib.lisec_channel_sums(C.c_void_p(self.dy.data_ptr()), self.P, Co, C.c_void_p(self.dbias.data_ptr()),
                                              C.c_void_p(self.ws.data_ptr()), self._stream())
            if st != N.LISEC_OK:
                raise N.LisecError(st, self._lib.lisec_bn_last_error().decode("utf-8", "replace"))
        dw = self.wgrad.run()  # [i][ci][(j, co)]
        self.dF = dw.view(s, Ci, s, Co).permute(0, 2, 3, 1)
        return self.dx.view(self.dx.shape[0], 1, self.dx.shape[1], self.dx.shape[3], Ci)

    def close(self):
        if getattr(self, "plan", None):
            self._lib.lisec_conv_plan_destroy(self.plan)
            self.plan = None
        self.wgrad.close()


def plan_layout(name: str, arr: np.ndarray) -> np.ndarray:
    """Keras layout -> the layout the training plans keep their float32 master weights in (DenseNetworkTrainer).
    A pure permutation of the elements, so the elementwise SGD update can run on either layout."""
    from .weights import conv3d_blocks, rpn_blocks

    a = np.asarray(arr, dtype=np.float32)
    leaf = name.split("/")[1]
    if leaf != "kernel":
        return np.ascontiguousarray(a)
    base = name.split("/")[0]
    c3 = {c: d for c, _, d, _, _ in conv3d_blocks()}
    if base in c3:
        return np.ascontiguousarray(a.reshape(27, 64, 64).transpose(0, 2, 1))
    if base in c3.values():
        return np.ascontiguousarray(a.T[None])
    for convs, (tname, k, s, cin) in rpn_blocks():
        for conv, _, ci, co, _ in convs:
            if base == conv:
                return np.ascontiguousarray(a.reshape(9, ci, co).transpose(0, 2, 1))
        if base == tname:
            return np.ascontiguousarray(a[::-1, ::-1].reshape(9, 256, cin)) if s == 1 else np.ascontiguousarray(a)
    raise KeyError(name)


def keras_layout(name: str, arr: np.ndarray) -> np.ndarray:
    """The inverse of plan_layout()."""
    from .weights import conv3d_blocks, rpn_blocks

    a = np.asarray(arr, dtype=np.float32)
    leaf = name.split("/")[1]
    if leaf != "kernel":
        return np.ascontiguousarray(a)
    base = name.split("/")[0]
    c3 = {c: d for c, _, d, _, _ in conv3d_blocks()}
    if base in c3:
        return np.ascontiguousarray(a.transpose(0, 2, 1).reshape(3, 3, 3, 64, 64))
    if base in c3.values():
        return np.ascontiguousarray(a[0].T)
    for convs, (tname, k, s, cin) in rpn_blocks():
        for conv, _, ci, co, _ in convs:
            if base == conv:
                return np.ascontiguousarray(a.transpose(0, 2, 1).reshape(3, 3, ci, co))
        if base == tname:
            return np.ascontiguousarray(a.reshape(3, 3, 256, cin)[::-1, ::-1]) if s == 1 else np.ascontiguousarray(a)
    raise KeyError(name)


class FlatStore:
    """One flat float32 buffer each for the variables, the optimizer accumulators and the gradients of the WHOLE model,
    carved into per-weight views as the trainers ask for them (every view 16-byte aligned): one SGD kernel launch and one
    all-reduce per step cover everything. Duck-compatible with FlatParameters for SgdNesterov."""

    def __init__(self, capacity: int, device):
        self.device = torch.device(device)
        self.var = torch.zeros(capacity, dtype=torch.float32, device=self.device)
        self.accum = torch.zeros_like(self.var)
        self.grad = torch.zeros_like(self.var)
        self.offsets, self.shapes, self.used = {}, {}, 0

    def alloc(self, name: str, arr: np.ndarray) -> torch.Tensor:
        a = np.ascontiguousarray(arr, dtype=np.float32)
        n = a.size
        if self.used + n > self.var.numel():
            raise RuntimeError("FlatStore capacity exceeded")
        self.offsets[name], self.shapes[name] = self.used, a.shape
        v = self.var[self.used:self.used + n].view(a.shape)
        v.copy_(torch.from_numpy(a))
        self.used += (n + 3) // 4 * 4
        return v

    def grad_view(self, name: str) -> torch.Tensor:
        o, sh = self.offsets[name], self.shapes[name]
        return self.grad[o:o + int(np.prod(sh))].view(sh)

    def var_view(self, name: str) -> torch.Tensor:
        o, sh = self.offsets[name], self.shapes[name]
        return self.var[o:o + int(np.prod(sh))].view(sh)

    @property
    def numel_padded(self) -> int:
        return self.used


class VfeTrainer:
    """The VFE stack (the first 23 Keras layers, model_training.py:229-235) in TRAINING mode on one GPU: forward with
    batch statistics into the dense grid, backward from the grid's gradient to the gradients of dense*/kernel and
    batch_normalization*/{gamma,beta} — lisec_vfe_train_forward / _backward (lisec_b200/csrc/vfe_train.cu), evaluated on
    rows with multiplicities exactly as oracle/train_oracle.py: forward_train_rows() states it.
    `params`: Keras-named float32 CUDA tensors (dense/kernel (6,16), batch_normalization/gamma, ..., moving statistics
    updated in place); `grads`: tensors of the same names and shapes that receive the gradients."""

    def __init__(self, frontend, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], eps=1e-3, momentum=0.99):
        from .weights import VFE_BN, VFE_DENSE

        self.fe, self._lib = frontend, N.load()
        self.params, self.grads = params, grads
        P, G = N.lisec_vfe_train_params(), N.lisec_vfe_train_grads()
        for i, (d, b) in enumerate(zip(VFE_DENSE, VFE_BN)):
            for t in (params[d + "/kernel"], params[b + "/gamma"], params[b + "/beta"], grads[d + "/kernel"],
                      grads[b + "/gamma"], grads[b + "/beta"]):
                if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                    raise ValueError("VfeTrainer wants contiguous float32 CUDA tensors")
            P.dense_kernel[i] = params[d + "/kernel"].data_ptr()
            P.bn_gamma[i] = params[b + "/gamma"].data_ptr()
            P.bn_beta[i] = params[b + "/beta"].data_ptr()
            mm, mv = params.get(b + "/moving_mean"), params.get(b + "/moving_variance")
            P.moving_mean[i] = mm.data_ptr() if mm is not None else None
            P.moving_var[i] = mv.data_ptr() if mv is not None else None
            G.dkernel[i] = grads[d + "/kernel"].data_ptr()
            G.dgamma[i] = grads[b + "/gamma"].data_ptr()
            G.dbeta[i] = grads[b + "/beta"].data_ptr()
        P.bn_epsilon, P.bn_momentum = eps, momentum
        self._P, self._G = P, G

    def forward(self, points, sweep_offsets, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """voxelize + the VFE stack with batch statistics -> grid [n_sweeps, nz, nx, ny, 64] (the handle's grid dtype)."""
        fe = self.fe
        fe.voxelize(points, sweep_offsets)
        if out is None:
            out = fe.new_grid(fe._n_sweeps)
        with torch.cuda.device(fe.device):
            fe._check(self._lib.lisec_vfe_train_forward(fe._h, C.byref(self._P), C.c_void_p(out.data_ptr()), fe._stream()))
        return out

    def backward(self, dgrid: torch.Tensor) -> None:
        fe = self.fe
        if dgrid.dtype != torch.float32 or not dgrid.is_contiguous():
            raise ValueError("dgrid: contiguous float32 [n_sweeps, nz, nx, ny, 64]")
        with torch.cuda.device(fe.device):
            fe._check(self._lib.lisec_vfe_train_backward(fe._h, C.byref(self._P), C.c_void_p(dgrid.data_ptr()),
                                                         C.byref(self._G), fe._stream()))

    def read_layer(self, layer: int, n_voxels: int):
        """(per-voxel output rows [n_voxels + 1, C], batch mean [C], batch variance [C]) of `layer` — a test aid."""
        Cl = (16, 32, 64)[layer]
        rows = np.zeros((n_voxels + 1, Cl), np.float32)
        mean, inv = np.zeros(Cl, np.float32), np.zeros(Cl, np.float32)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
        self.fe._check(self._lib.lisec_vfe_train_read(self.fe._h, layer, fp(rows), n_voxels + 1, fp(mean), fp(inv)))
        return rows, mean, 1.0 / inv.astype(np.float64) ** 2 - 1e-3


class DenseNetworkTrainer:
    """The dense network behind the voxel grid in TRAINING mode (model_training.py:236-254 under fit): three Conv3D blocks,
    the RPN's sixteen Conv2D + BN + ReLU stages, the three transposed convolutions into the concat tensor, the heads and the
    two MSE terms — forward, loss and the gradient of every one of its parameters, chained from the stages above.
    `pack`: Keras-named float arrays (lisec_b200/weights.py); the float32 master weights live in self.params[name] in the
    plans' layouts (see _to_plan_layout), their gradients in self.grads[name] after backward(). The VFE stack in front of
    the grid is not part of this class (its training kernels are not built yet): `grid` is an input."""

    def __init__(self, pack: dict, batch: int, nx: int, ny: int, nz: int = 8, device: int = 0, alloc=None,
                 need_grid_grad: bool = False, grad_alloc=None):
        """alloc(name, float32 array) -> CUDA tensor holding it: where a trainable weight's master copy lives (default:
        its own tensor; TrainStep hands out views of one flat buffer). need_grid_grad: also compute d loss / d grid
        (float32, self.grid_grad after loss_and_backward) for the VFE stack in front. grad_alloc(name) -> the float32 CUDA
        tensor that weight's gradient belongs in (TrainStep: a view of the flat gradient buffer): the weight-gradient plans
        and BatchNormalization backward kernels of the plain stages then write there directly (_GradSink)."""
        from .weights import conv3d_blocks, rpn_blocks

        self._lib = N.load()
        dev = torch.device("cuda", device)
        self.device = dev
        B = batch
        own = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)  # noqa: E731

        def mk(name, a):  # trainable weights go through alloc(); moving statistics stay in tensors of their own
            return alloc(name, a) if (alloc is not None and "moving_" not in name) else own(a)

        _GradSink.current = {}  # (nothing left over from a construction that failed half way)
        self.params, self.stages = {}, []
        self.zero_grads = set()  # biases in front of a training-mode BatchNormalization: gradient identically zero
        self.grid_grad = None

        def sink(**names):  # the gradient tensors of the stage constructed next
            _GradSink.current = {k: grad_alloc(v) for k, v in names.items()} if grad_alloc is not None else {}
        self.grid = torch.zeros((B, nz, nx, ny, 64), dtype=torch.bfloat16, device=dev)
        P = self.params
        x, d = self.grid, nz
        self.c3 = []
        for i, (conv, bn, dense, stride, pad) in enumerate(conv3d_blocks()):
            P[conv + "/kernel"] = mk(conv + "/kernel", np.asarray(pack[conv + "/kernel"]).reshape(27, 64, 64).transpose(0, 2, 1))
            P[dense + "/kernel"] = mk(dense + "/kernel", np.asarray(pack[dense + "/kernel"]).T[None])
            for f in ("bias",):
                P[conv + "/" + f] = mk(conv + "/" + f, pack[conv + "/" + f])
            for f in ("gamma", "beta", "moving_mean", "moving_variance"):
                P[bn + "/" + f] = mk(bn + "/" + f, pack[bn + "/" + f])
            sink(dw=conv + "/kernel", dwd=dense + "/kernel", dgamma=bn + "/gamma", dbeta=bn + "/beta")
            self.zero_grads.add(conv + "/bias")
            st = Conv3dBlockTrain(x, P[conv + "/kernel"], P[conv + "/bias"], P[bn + "/gamma"], P[bn + "/beta"],
                                  P[dense + "/kernel"], (3, 3, 3), pad, stride_d=stride[0],
                                  moving_mean=P[bn + "/moving_mean"], moving_var=P[bn + "/moving_variance"],
                                  need_dx=i > 0 or need_grid_grad,
                                  grad_dtype=torch.float32)
            self.c3.append((st, conv, bn, dense))
            x = st.y
        assert x.shape[1] == 1
        self.concat = torch.zeros((B, 1, nx // 2, ny // 2, 768), dtype=torch.bfloat16, device=dev)
        self.blocks = []
        for bi, (convs, (tname, k, s, tc_in)) in enumerate(rpn_blocks()):
            stages = []
            for conv, bn, cin, cout, stride in convs:
                P[conv + "/kernel"] = mk(conv + "/kernel", np.asarray(pack[conv + "/kernel"]).reshape(9, cin, cout).transpose(0, 2, 1))
                P[conv + "/bias"] = mk(conv + "/bias", pack[conv + "/bias"])
                for f in ("gamma", "beta", "moving_mean", "moving_variance"):
                    P[bn + "/" + f] = mk(bn + "/" + f, pack[bn + "/" + f])
                sink(dw=conv + "/kernel", dgamma=bn + "/gamma", dbeta=bn + "/beta")
                self.zero_grads.add(conv + "/bias")
                st = ConvBnReluTrain(x, P[conv + "/kernel"], P[conv + "/bias"], P[bn + "/gamma"], P[bn + "/beta"], (1, 3, 3),
                                     (0, 1, 1), moving_mean=P[bn + "/moving_mean"], moving_var=P[bn + "/moving_variance"],
                                     stride_hw=stride, grad_dtype=torch.float32)
                stages.append((st, conv, bn))
                x = st.bn.y
            F = np.asarray(pack[tname + "/kernel"], dtype=np.float32)  # (k, k, 256, cin)
            P[tname + "/bias"] = mk(tname + "/bias", pack[tname + "/bias"])
            dy_t = torch.zeros((B, 1, nx // 2, ny // 2, 256), dtype=torch.bfloat16, device=dev)
            if s == 1:
                P[tname + "/kernel"] = mk(tname + "/kernel", F[::-1, ::-1].reshape(9, 256, tc_in))  # the flipped-kernel convolution's layout
                sink(dw=tname + "/kernel", dbias=tname + "/bias")
                tail = ConvBiasTrain(x, P[tname + "/kernel"], P[tname + "/bias"], (1, 3, 3), (0, 1, 1), self.concat, 256 * bi, dy_t,
                                     grad_dtype=torch.float32)
            else:
                P[tname + "/kernel"] = mk(tname + "/kernel", F)  # Keras layout (k, k, 256, cin)
                sink()
                tail = _ShuffleTail(self._lib, x, P[tname + "/kernel"], P[tname + "/bias"], s, self.concat, 256 * bi, dy_t)
            self.blocks.append((stages, tail, tname, s, dy_t, x))
        Kh = np.concatenate([np.asarray(pack["ClassificationLayer/kernel"])[0, 0], np.asarray(pack["RegressionLayer/kernel"])[0, 0]], axis=1)
        P["heads/kernel"] = mk("heads/kernel", Kh.T[None])  # [1][16][768]: rows 0-1 ClassificationLayer, 2-15 RegressionLayer
        P["heads/bias"] = mk("heads/bias", np.concatenate([pack["ClassificationLayer/bias"], pack["RegressionLayer/bias"]]))
        sink()
        self.heads = HeadsTrain(self.concat, P["heads/kernel"], P["heads/bias"], dense_slices=True)
        _GradSink.current = {}
        self.grads = {}

    def forward(self, grid: Optional[torch.Tensor] = None):
        if grid is not None:
            self.grid.copy_(grid)
        for st, *_ in self.c3:
            st.forward()
        for stages, tail, *_ in self.blocks:
            for st, *_ in stages:
                st.forward()
            tail.forward()
        y = self.heads.forward()
        return y[:, 0, :, :, :2], y[:, 0, :, :, 2:]

    def loss_and_backward(self, y_class: torch.Tensor, y_regress: torch.Tensor) -> torch.Tensor:
        """loss=['mse','mse'] (:296) on the last forward(), then the backward pass: self.grads[...] afterwards."""
        y = self.heads.y
        lc, dc = mse_loss_grad(y[..., :2].contiguous(), y_class.reshape(y[..., :2].shape).contiguous())
        lr, dr = mse_loss_grad(y[..., 2:].contiguous(), y_regress.reshape(y[..., 2:].shape).contiguous())
        parts = self.heads.backward(torch.cat([dc, dr], dim=-1))
        G = self.grads
        G["heads/kernel"], G["heads/bias"] = self.heads.dw, self.heads.dbias
        carry = None  # gradient arriving at a block's output from the NEXT block
        for bi in (2, 1, 0):
            stages, tail, tname, s, dy_t, x_out = self.blocks[bi]
            dy_t.copy_(parts[bi])
            dx = tail.backward()
            G[tname + "/kernel"], G[tname + "/bias"] = tail.dw, tail.dbias
            if carry is not None:
                add_(dx, carry)  # the block's output has two consumers
            for st, conv, bn in reversed(stages):
                dx = st.backward(dx)
                G[conv + "/kernel"], G[conv + "/bias"] = st.dw, st.dbias
                G[bn + "/gamma"], G[bn + "/beta"] = st.bn.dgamma, st.bn.dbeta
            carry = dx
        dx = carry
        for st, conv, bn, dense in reversed(self.c3):
            dx = st.backward(dx)
            self.grid_grad = dx  # after the loop: the first block's data gradient = d loss / d grid (or None)
            G[conv + "/kernel"], G[conv + "/bias"], G[dense + "/kernel"] = st.dw, st.dbias, st.dwd
            G[bn + "/gamma"], G[bn + "/beta"] = st.bn.dgamma, st.bn.dbeta
        return lc + lr

    def refresh_weights(self) -> None:
        """After the float32 master weights changed (an optimizer step): re-derive the plans' bf16 operand copies. The
        plain stages' refreshes (casts and flip-transposes, ~45 of them) are recorded once and run as ONE launch."""
        plain = [st for st, *_ in self.c3]
        for stages, tail, *_ in self.blocks:
            plain += [st for st, *_ in stages]
            if isinstance(tail, ConvBiasTrain):
                plain.append(tail)
        if getattr(self, "_refresh_table", None) is None:
            rec: list = []
            _RefreshBatch.recording = rec
            try:
                for st in plain:
                    st.refresh_weights()
            finally:
                _RefreshBatch.recording = None
            tab = np.zeros(len(rec), dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("first", "<i8"), ("kd", "<i4"), ("kh", "<i4"),
                                                     ("kw", "<i4"), ("out_c", "<i4"), ("in_c", "<i4"), ("mode", "<i4")]))
            first = 0
            for i, (src, dst, n, kd, kh, kw, oc, ic, mode) in enumerate(rec):
                tab[i] = (src, dst, first, kd, kh, kw, oc, ic, mode)
                first += n
            self._refresh_total = first
            self._refresh_table = torch.from_numpy(tab.view(np.uint8).copy()).to(self.device)
            self._refresh_n = len(rec)
        with torch.cuda.device(self.device):
            st_ = self._lib.lisec_refresh_operands(C.c_void_p(self._refresh_table.data_ptr()), self._refresh_n, self._refresh_total,
                                                   C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if st_ != N.LISEC_OK:
            raise N.LisecError(st_, self._lib.lisec_train_last_error().decode("utf-8", "replace"))
        for stages, tail, *_ in self.blocks:
            if not isinstance(tail, ConvBiasTrain):
                tail.refresh_weights()
        self.heads.refresh_weights()

    def close(self):
        for st, *_ in self.c3:
            st.close()
        for stages, tail, *_ in self.blocks:
            for st, *_ in stages:
                st.close()
            tail.close()
        self.heads.close()


class TrainStep:
    """One fit() step of model_training.train() (model_training.py:295-299) on one GPU, the whole model:
    voxelize -> VFE stack with batch statistics (VfeTrainer) -> dense network (DenseNetworkTrainer, bf16 tensor-core plans,
    float32 master weights) -> loss=['mse','mse'] -> backward through both -> [NCCL all-reduce of the flat gradient, summed;
    per-replica BatchNormalization statistics] -> optimizers.SGD(lr=0.01, decay=1e-6, momentum=0.9, nesterov=True) on one
    flat buffer -> the plans' operand copies refreshed. `pack`: Keras-named float arrays of the whole createModel()."""

    def __init__(self, pack: dict, batch: int, max_points: int, device: int = 0, nx: int = 200, ny: int = 400, nz: int = 8,
                 lr=0.01, decay=1e-6, momentum=0.9, nesterov=True, group=None, bucket_elems: int = 0, use_graph: bool = False):
        """use_graph (off by default; measured 12.18 -> 11.69 ms per 2-sweep step on B200 — the step is bound by its kernels,
        not by their launches): after one eager step, the dense network's forward + loss + backward + the gradient copies into the flat
        buffer (~330 launches of 5-50 us kernels at batch 2, issued from Python through ctypes) and the refresh of the plans'
        operand copies are captured into two CUDA graphs and replayed — the same kernels on the same buffers, without the
        launch overhead. The VFE stack (launch sizes depend on the step's point count), the all-reduce and the optimizer
        update (its learning rate changes every iteration) stay eager."""
        from .frontend import Frontend
        from .weights import VFE_BN, VFE_DENSE

        self.group, self.bucket_elems, self.batch = group, bucket_elems, batch
        self.use_graph = bool(use_graph)
        self.collective = True  # False: skip the all-reduce (a measurement aid: the step's compute alone)
        self._graph_dense = self._graph_refresh = None
        self._steps = 0
        dev = torch.device("cuda", device)
        n_train = sum(int(np.prod(np.shape(v))) for k, v in pack.items() if "moving_" not in k)
        self.store = FlatStore(n_train + 4 * len(pack), dev)
        self.fe = Frontend(device=device, max_points=max_points, max_sweeps=batch, grid_dtype="bf16",
                           max_voxel=(nx // 2, ny // 2, nz))
        # the VFE stack's parameters first (Keras creation order), Keras layout
        self.vfe_params, vfe_grads = {}, {}
        for d, b in zip(VFE_DENSE, VFE_BN):
            for k in (d + "/kernel", b + "/gamma", b + "/beta"):
                self.vfe_params[k] = self.store.alloc(k, pack[k])
            for f in ("moving_mean", "moving_variance"):
                self.vfe_params[b + "/" + f] = torch.from_numpy(np.ascontiguousarray(pack[b + "/" + f], dtype=np.float32)).to(dev)
        self._vfe_end = self.store.used  # flat layout: [VFE stack | dense network]
        self.dense = DenseNetworkTrainer(pack, batch, nx, ny, nz, device=device, alloc=self.store.alloc, need_grid_grad=True,
                                         grad_alloc=self.store.grad_view)
        for k in list(self.vfe_params):
            if "moving_" not in k:
                vfe_grads[k] = self.store.grad_view(k)
        self.vfe = VfeTrainer(self.fe, self.vfe_params, vfe_grads)
        self.opt = SgdNesterov(self.store, lr, decay, momentum, nesterov)
        self.pack_names = list(pack)
        self.exposed_allreduce_ms = None

    def step(self, points, sweep_offsets, y_class: torch.Tensor, y_regress: torch.Tensor) -> torch.Tensor:
        """points: (n, 3) host or device array of the batch's sweeps, concatenated; labels: float32 CUDA tensors
        (B, nx/2, ny/2, 2) and (B, nx/2, ny/2, 14). Returns the loss (0-d float64 device tensor)."""
        import torch.distributed as dist

        self.vfe.forward(points, sweep_offsets, out=self.dense.grid)
        if not self.use_graph:
            loss = self._dense_region(y_class, y_regress)
        else:
            if self._steps == 0:  # static label buffers the captured kernels read
                self._yc, self._yr = torch.empty_like(y_class), torch.empty_like(y_regress)
            self._yc.copy_(y_class)
            self._yr.copy_(y_regress)
            if self._graph_dense is None and self._steps >= 1:  # step 0 ran eagerly: every lazy workspace exists
                torch.cuda.synchronize(self.dense.device)
                self._graph_dense = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph_dense):
                    self._loss = self._dense_region(self._yc, self._yr)
            if self._graph_dense is not None:
                self._graph_dense.replay()
                loss = self._loss.clone()
            else:
                loss = self._dense_region(self._yc, self._yr)
        # The collective overlaps the rest of the backward pass: the dense network's gradients (25.9 MB, all but 21 KB of the
        # flat buffer) are complete here and their all-reduce runs on NCCL's stream WHILE the VFE stack's backward (~1.3 ms)
        # runs on the compute stream; only the VFE stack's own 21 KB are reduced after it.
        world, works = 1, []
        if self.collective and dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.group)
            works = allreduce_gradients(self.store.grad[self._vfe_end:self.store.numel_padded], self.group, self.bucket_elems)
        self.vfe.backward(self._grid_grad)
        if world > 1:
            works += allreduce_gradients(self.store.grad[:self._vfe_end], self.group)
        for w in works:
            w.wait()
        self.opt.step(world)
        if self.use_graph and self._graph_refresh is None and self._steps >= 1:
            torch.cuda.synchronize(self.dense.device)
            self._graph_refresh = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_refresh):
                self.dense.refresh_weights()
        if self._graph_refresh is not None:
            self._graph_refresh.replay()
        else:
            self.dense.refresh_weights()
        self._steps += 1
        return loss

    def _dense_region(self, y_class, y_regress) -> torch.Tensor:
        """The statically shaped part of the step: dense forward, the two MSE terms, the backward pass, the gradients into
        the flat buffer (plan layout). Leaves d loss / d grid in self._grid_grad for the VFE backward."""
        self.dense.forward()
        loss = self.dense.loss_and_backward(y_class, y_regress)
        self._grid_grad = self.dense.grid_grad.contiguous()
        for name, g in self.dense.grads.items():  # what the stages did not write into the flat buffer themselves
            dst = self.store.grad_view(name)
            if g.data_ptr() != dst.data_ptr() and name not in self.dense.zero_grads:
                dst.copy_(g.reshape(self.store.shapes[name]))
        return loss

    def to_pack(self) -> Dict[str, np.ndarray]:
        """The model's weights under the Keras names and in the Keras layouts (what model.save() persists)."""
        out = {}
        P = self.dense.params
        for k in self.pack_names:
            base, leaf = k.split("/")
            if k in self.vfe_params:
                out[k] = self.vfe_params[k].detach().cpu().numpy().copy()
            elif base in ("ClassificationLayer", "RegressionLayer"):
                sl = slice(0, 2) if base == "ClassificationLayer" else slice(2, 16)
                if leaf == "kernel":
                    out[k] = np.ascontiguousarray(P["heads/kernel"][0, sl].detach().cpu().numpy().T[None, None])
                else:
                    out[k] = P["heads/bias"][sl].detach().cpu().numpy().copy()
            else:
                out[k] = keras_layout(k, P[k].detach().cpu().numpy())
        return out

    def close(self):
        self.dense.close()
        self.fe.close()


class _ShuffleTail:
    """A kernel = stride Conv2DTranspose in training: forward = the pixel-shuffle plan (a 1x1 GEMM with s*s N-tiles of 256
    columns written to (s*h + i, s*w + j) of a concat slice), backward = ConvTransposeBackward."""

    def __init__(self, lib, x, F, bias, s, concat, ch_off, dy):
        self._lib, self.x, self.F, self.s = lib, x, F, s
        B, _, H, W, Cin = x.shape
        dev = x.device
        self.w16 = torch.empty((1, s * s * 256, Cin), dtype=torch.bfloat16, device=dev)
        self.wf = torch.empty((1, s * s * 256, Cin), dtype=torch.float32, device=dev)
        self.ones = torch.ones(256, dtype=torch.float32, device=dev)
        self.bwd = ConvTransposeBackward(x, dy, F, s, grad_dtype=torch.float32)
        self.refresh_weights()
        tile = (16, 8) if W >= 16 else (8, 16)
        desc = N.lisec_conv_desc(
            batch=B, in_d=1, in_h=H, in_w=W, in_c=Cin, kd=1, kh=1, kw=1, stride_d=1, stride_hw=1, pad_d=0, pad_h=0, pad_w=0,
            out_c=256, n_tiles=s * s, shuffle=s, out_pitch=concat.shape[-1], out_ch_off=ch_off, relu=0,
            out_dtype=N.LISEC_BF16, tile_w=tile[0], tile_h=tile[1], m_tiles=1, in_dtype=N.LISEC_BF16, out_split=0, group_kh=0,
            reserved=1)  # bit 0: the bias is a trainable weight, rewritten between runs
        self.plan = C.c_void_p()
        with torch.cuda.device(dev):
            st = lib.lisec_conv_plan_create(C.byref(desc), C.c_void_p(x.data_ptr()), C.c_void_p(self.w16.data_ptr()),
                                            C.c_void_p(self.ones.data_ptr()), C.c_void_p(bias.data_ptr()),
                                            C.c_void_p(concat.data_ptr()), C.byref(self.plan))
        if st != N.LISEC_OK:
            raise N.LisecError(st, lib.lisec_conv_last_error().decode("utf-8", "replace"))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.x.device).cuda_stream)

    def refresh_weights(self):
        self.wf.copy_(self.F.reshape(1, -1, self.F.shape[3]))  # (i, j, co) major, ci contiguous: the plan's [N][C]
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_cast_f32_to_bf16(C.c_void_p(self.wf.data_ptr()), self.wf.numel(), C.c_void_p(self.w16.data_ptr()),
                                                  self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_train_last_error().decode("utf-8", "replace"))
        self.bwd.refresh_weights()

    def forward(self):
        with torch.cuda.device(self.x.device):
            st = self._lib.lisec_conv_plan_run(self.plan, self._stream())
        if st != N.LISEC_OK:
            raise N.LisecError(st, self._lib.lisec_conv_last_error().decode("utf-8", "replace"))

    def backward(self):
        dx = self.bwd.backward()
        self.dw, self.dbias = self.bwd.dF, self.bwd.dbias
        return dx

    def close(self):
        if self.plan:
            self._lib.lisec_conv_plan_destroy(self.plan)
            self.plan = None
        self.bwd.close()

"""The dense network behind the voxel grid — middle Conv3D stack, RPN, heads (reference model_training.py:236-256) —
as a list of tensor-core convolution plans (lisec_conv_plan_*, lisec_b200/csrc/conv.cu).

Host side of SURVEY §8 row a12. What Keras builds layer by layer in createModel, this module builds plan by plan; the
affine tails are folded on the host in float64 before anything is rounded to the bf16 operands:

  addConv3DLayer (:191-196)  ZeroPadding3D + Conv3D(bias) + BatchNormalization + Dense(64, relu, no bias)
      conv -> BN -> Dense is linear up to the ReLU:   W'[tap,ci,n] = sum_c K[tap,ci,c] g[c] Wd[c,n],
      shift'[n] = sum_c ((bias[c] - mean[c]) g[c] + beta[c]) Wd[c,n],  g = gamma / sqrt(var + 1e-3)  -> one plan, ReLU
  addConv2DLayer (:201-208)  Conv2D(bias) + BatchNormalization + ReLU -> scale = g, shift = (bias - mean) g + beta
      with Architecture.post_dense (the line commented out at :205 switched on, the graph model.png shows): Conv2D + BN +
      Dense(cout, relu, no bias) folds exactly like addConv3DLayer; the grid and the first Conv3D then carry C3 = 128 channels
  Conv2DTranspose (:245,248,251, padding='same')  k3 s1 = 3x3 convolution with the kernel flipped; k2 s2 / k4 s4 do not
      overlap = 1x1 GEMMs with k*k groups of 256 columns, pixel-shuffled by the epilogue. Each writes its 256-channel
      slice of the Concatenate (:252) buffer directly.
  ClassificationLayer + RegressionLayer (:253-254)  one 1x1 plan with N = 2 + 14 = 16 columns, float32 out.

Activations are channels-last [B, D, H=x, W=y, C]; accumulation is float32 in TMEM. Two arithmetic modes:
  dtype="bf16"  bf16 operands and activations (north_star's 2e-2 bar), the fast path;
  dtype="f32"   every operand is a pair of float32 planes (hi = tf32-rounded, lo = remainder) and every product runs as
                3xTF32 (Ah*Bl + Al*Bh + Ah*Bh), which reproduces float32 products (north_star's 1e-5 bar); activations
                are stored as [2][B, D, H, W, C] hi/lo planes by each plan's epilogue.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _native
from .weights import CURRENT, Architecture, conv2d_post_dense, conv3d_blocks, rpn_blocks, validate_network_pack

BN_EPS = 1e-3


def best_tile(out_h: int, out_w: int, m_tiles: int = 1, min_w: int = 1) -> Tuple[int, int]:
    """tile_w x tile_h = 128 output positions per M-tile, tile_w a power of two: the shape whose CTA tiles
    (tile_w x m_tiles*tile_h) waste the fewest positions."""
    best = None
    for lg in range(0, 8):
        tw, th = 1 << lg, 128 >> lg
        if tw < min_w:
            continue
        cover = -(-out_w // tw) * tw * -(-out_h // (th * m_tiles)) * th * m_tiles
        key = (cover, abs(lg - 4))
        if best is None or key < best[0]:
            best = (key, (tw, th))
    return best[1]


def tf32_round(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest tf32 (10 mantissa bits), ties away from zero: what cvt.rna.tf32.f32 does on the device."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _bn_fold(pack, bn):
    g = pack[bn + "/gamma"].astype(np.float64) / np.sqrt(pack[bn + "/moving_variance"].astype(np.float64) + BN_EPS)
    return g, pack[bn + "/beta"].astype(np.float64) - pack[bn + "/moving_mean"].astype(np.float64) * g


class _Layer:
    def __init__(self, name, desc, w, scale, shift, src, dst):
        self.name, self.desc, self.w, self.scale, self.shift, self.src, self.dst = name, desc, w, scale, shift, src, dst
        self.plan = C.c_void_p()


def default_schedule(k, stride_hw: int, in_c: int, out_c: int, n_tiles: int) -> List[Tuple[int, int]]:
    """Candidate (m_tiles, group_kh) settings of a layer, best first; the first one the library accepts is used.
    The kernel is bound by L2 -> shared-memory delivery (profiles/conv_r1j_summary.txt): fetching the kh taps from one
    halo box cuts the input traffic, two M-tiles per weight box halve the weight traffic. Both need room — 2 * m_tiles *
    out_c accumulator columns <= 512 and at least two pipeline stages of shared memory — which the library checks."""
    mts = (2, 1) if out_c <= 128 else (1,)
    groups = (1, 0) if (stride_hw == 1 and k[1] == 3) else (0,)
    return [(mt, g) for g in groups for mt in mts]


def halo_schedule(k, stride_hw: int, in_c: int, out_c: int, n_tiles: int) -> List[Tuple[int, int]]:
    """default_schedule with the "halo" plans (group_kh = 2) first where they apply: one 18 x 18 input box per (kd, 64
    channels) serves all nine (kh, kw) taps — every tap is a descriptor into the same box (tools/umma_shift_probe.cu).
    Measured on B200: L2 -> SM bytes of the Conv3D blocks drop another 1.6x; with one MMA-issuing thread per M-tile the
    Conv3D blocks run at 1096 TFLOP/s (1040 with the kh-halo plans), the whole network 1.56 ms per 8 sweeps (1.60).
    This is the schedule DenseNetwork uses."""
    cands = default_schedule(k, stride_hw, in_c, out_c, n_tiles)
    if stride_hw == 1 and k[1] == 3 and k[2] == 3 and out_c <= 128:
        cands = [(2, 2)] + cands
    return cands


class DenseNetwork:
    """Plans and activation buffers for one batch size and grid. `grid` is the input buffer the front end writes."""

    def __init__(self, pack: dict, batch: int, nx: int = 200, ny: int = 400, nz: int = 8, device: int = 0,
                 schedule=None, dtype: str = "bf16", fuse_heads: bool = True, arch: Architecture = CURRENT):
        if nz != 8 or nx % 8 or ny % 8:
            raise ValueError("the Conv3D stack collapses nz = 8 to 1 and the RPN halves x, y three times: need nz = 8 "
                             "and nx, ny multiples of 8 (got %d, %d, %d)" % (nz, nx, ny))
        if not torch.cuda.is_available():
            raise RuntimeError("lisec_b200 has no CPU fallback: a CUDA device is required")
        self._lib = _native.load()
        self.device = torch.device("cuda", device)
        self.batch, self.nx, self.ny, self.nz = batch, nx, ny, nz
        if dtype not in ("bf16", "f32"):
            raise ValueError("dtype must be 'bf16' or 'f32', got %r" % (dtype,))
        self.f32 = dtype == "f32"
        self._schedule = (lambda *a: [(1, 0)]) if self.f32 else (schedule or halo_schedule)
        self._auto_schedule = schedule is None and not self.f32
        self.fuse_heads = bool(fuse_heads)
        self._sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.arch = arch
        pack = validate_network_pack(pack, arch)
        c3 = arch.c3
        dev = self.device
        f32 = self.f32

        def buf(*shape, dtype=None, planes=True):
            if f32:  # hi / lo planes in front
                return torch.zeros(((2,) if planes else ()) + shape, dtype=torch.float32, device=dev)
            return torch.zeros(shape, dtype=dtype or torch.bfloat16, device=dev)

        B = batch
        self.grid = buf(B, nz, nx, ny, c3, planes=False)  # what the front end writes: bf16, or plain float32
        self.grid_planes = buf(B, nz, nx, ny, c3) if f32 else self.grid
        self.layers: List[_Layer] = []
        # ---- middle: three Conv3D blocks (:236-238) ----
        src, d, cin3 = self.grid_planes, nz, c3
        for conv, bn, dense, stride, pad in conv3d_blocks(arch):
            K = pack[conv + "/kernel"].astype(np.float64).reshape(27, cin3, 64)  # [tap (kd,kh,kw)][ci][c]
            g, b0 = _bn_fold(pack, bn)
            Wd = pack[dense + "/kernel"].astype(np.float64)
            W = np.einsum("tic,c,cn->tni", K, g, Wd)
            shift = (pack[conv + "/bias"].astype(np.float64) * g + b0) @ Wd
            od = (d + 2 * pad[0] - 3) // stride[0] + 1
            dst = buf(B, od, nx, ny, 64)
            self._add(conv, src, dst, W, np.ones(64), shift, in_d=d, in_h=nx, in_w=ny, in_c=cin3, k=(3, 3, 3),
                      stride_d=stride[0], stride_hw=1, pad=pad, out_c=64, relu=1)
            src, d, cin3 = dst, od, 64
        assert d == 1
        # ---- RPN (:245-251) ----
        h, w = nx, ny
        # Tail of the network (:247-254). Conv2DTranspose has no activation and no BatchNormalization behind it, Concatenate is
        # a layout, the heads are 1x1 convolutions: per RPN block, transposed convolution -> its 256 rows of the head kernels
        # is ONE linear map. fuse_heads folds them on the host (float64): block 1 becomes a 3x3 plan with 16 output columns
        # at the output resolution, blocks 2 and 3 become 1x1 plans with s*s*16 columns at their own resolution, and
        # lisec_heads_combine adds the three. The [B,100,200,768] concat tensor (245 MB per 8 sweeps, written and re-read)
        # and 95 % of the tail's multiply-adds disappear; nothing is rounded to bf16 between the blocks and the heads.
        Kh = np.concatenate([pack["ClassificationLayer/kernel"][0, 0], pack["RegressionLayer/kernel"][0, 0]],
                            axis=1).astype(np.float64)  # (768, 16)
        bh = np.concatenate([pack["ClassificationLayer/bias"], pack["RegressionLayer/bias"]]).astype(np.float64)
        self.heads = buf(B, 1, nx // 2, ny // 2, 16, dtype=torch.float32, planes=False)
        self._parts = []
        if self.fuse_heads:
            for (_, (tname, k, s, tc_in)), bi in zip(rpn_blocks(), range(3)):
                bh = bh + pack[tname + "/bias"].astype(np.float64) @ Kh[256 * bi:256 * bi + 256]
        else:
            self.concat = buf(B, 1, nx // 2, ny // 2, 768)
        post = conv2d_post_dense(arch)
        for bi, (convs, (tname, k, s, tc_in)) in enumerate(rpn_blocks()):
            pp = None
            for conv, bn, cin, cout, stride in convs:
                K = pack[conv + "/kernel"].astype(np.float64).reshape(9, cin, cout)
                g, b0 = _bn_fold(pack, bn)
                W = K.transpose(0, 2, 1)
                shift = pack[conv + "/bias"].astype(np.float64) * g + b0
                if conv in post:  # Conv2D -> BN -> Dense(relu): linear up to the ReLU, folded like the Conv3D blocks
                    Wd = pack[post[conv] + "/kernel"].astype(np.float64)
                    W = np.einsum("tic,c,cn->tni", K, g, Wd)
                    shift = shift @ Wd
                    g = np.ones(cout)
                oh, ow = h // stride, w // stride
                if pp is None:
                    pp = [buf(B, 1, oh, ow, cout), buf(B, 1, oh, ow, cout)]
                dst = pp[0] if src is not pp[0] else pp[1]
                self._add(conv, src, dst, W, g, shift, in_d=1, in_h=h, in_w=w, in_c=cin, k=(1, 3, 3), stride_d=1,
                          stride_hw=stride, pad=(0, 1, 1), out_c=cout, relu=1)
                src, h, w = dst, oh, ow
            F = pack[tname + "/kernel"].astype(np.float64)  # (k, k, 256, cin)
            bias = pack[tname + "/bias"].astype(np.float64)
            if self.fuse_heads:
                Kb = Kh[256 * bi:256 * bi + 256]  # this block's rows of the head kernels
                if s == 1:  # flipped 3x3 kernel (see below), then 256 -> 16 columns; every bias rides on this plan
                    W = np.einsum("toc,on->tnc", F[::-1, ::-1].reshape(9, 256, tc_in), Kb)
                    part = buf(B, 1, h, w, 16, dtype=torch.float32, planes=False)
                    self._add(tname + "+heads", src, part, W, np.ones(16), bh, in_d=1, in_h=h, in_w=w, in_c=tc_in,
                              k=(1, 3, 3), stride_d=1, stride_hw=1, pad=(0, 1, 1), out_c=16, relu=0,
                              out_dtype=_native.LISEC_F32, out_split=0)
                else:  # kernel == stride: column group (i, j) of a 1x1 GEMM belongs to output pixel (s y + i, s x + j)
                    W = np.einsum("ijoc,on->ijnc", F, Kb).reshape(1, k * k * 16, tc_in)
                    part = buf(B, 1, h, w, k * k * 16, dtype=torch.float32, planes=False)
                    self._add(tname + "+heads", src, part, W, np.ones(k * k * 16), np.zeros(k * k * 16), in_d=1, in_h=h,
                              in_w=w, in_c=tc_in, k=(1, 1, 1), stride_d=1, stride_hw=1, pad=(0, 0, 0), out_c=k * k * 16,
                              relu=0, out_dtype=_native.LISEC_F32, out_split=0)
                self._parts.append((part, s))
            elif s == 1:  # 'same' k3 s1: out[y] = sum_ky in[y + 1 - ky] F[ky] = a 3x3 convolution with the flipped kernel
                W = F[::-1, ::-1].reshape(9, 256, tc_in)
                self._add(tname, src, self.concat, W, np.ones(256), bias, in_d=1, in_h=h, in_w=w, in_c=tc_in,
                          k=(1, 3, 3), stride_d=1, stride_hw=1, pad=(0, 1, 1), out_c=256, relu=0, out_pitch=768,
                          out_ch_off=256 * bi)
            else:  # kernel == stride: out[s y + i] = in[y] F[i]
                W = F.reshape(1, k * k * 256, tc_in)
                self._add(tname, src, self.concat, W, np.ones(256), bias, in_d=1, in_h=h, in_w=w, in_c=tc_in,
                          k=(1, 1, 1), stride_d=1, stride_hw=1, pad=(0, 0, 0), out_c=256, relu=0, out_pitch=768,
                          out_ch_off=256 * bi, n_tiles=k * k, shuffle=s)
        # ---- heads (:253-254): 2 + 14 columns of one 1x1 GEMM ----
        if not self.fuse_heads:
            self._add("heads", self.concat, self.heads, Kh.T.reshape(1, 16, 768), np.ones(16), bh, in_d=1,
                      in_h=nx // 2, in_w=ny // 2, in_c=768, k=(1, 1, 1), stride_d=1, stride_hw=1, pad=(0, 0, 0), out_c=16,
                      relu=0, out_dtype=_native.LISEC_F32, out_split=0)
        self._graph: Optional[torch.cuda.CUDAGraph] = None

    def _add(self, name, src, dst, W, scale, shift, *, in_d, in_h, in_w, in_c, k, stride_d, stride_hw, pad, out_c, relu,
             out_pitch=None, out_ch_off=0, n_tiles=1, shuffle=1, out_dtype=None, out_split=None):
        if out_dtype is None:
            out_dtype = _native.LISEC_F32 if self.f32 else _native.LISEC_BF16
        if out_split is None:
            out_split = 1 if self.f32 else 0
        if out_pitch is None:
            out_pitch = n_tiles * out_c if shuffle == 1 else out_c
        if self.f32 and out_c > 128:  # float32 plans keep a row's running sums in registers: N-tiles of 128 columns
            n_tiles, out_c = n_tiles * (out_c // 128), 128
        od = (in_d + 2 * pad[0] - k[0]) // stride_d + 1
        oh = (in_h + 2 * pad[1] - k[1]) // stride_hw + 1
        ow = (in_w + 2 * pad[2] - k[2]) // stride_hw + 1
        dev = self.device
        sc = torch.from_numpy(np.ascontiguousarray(scale, dtype=np.float32)).to(dev)
        sh = torch.from_numpy(np.ascontiguousarray(shift, dtype=np.float32)).to(dev)
        cands = self._schedule(k, stride_hw, in_c, out_c, n_tiles)
        if isinstance(cands, tuple):
            cands = [cands]
        if self._auto_schedule and cands and tuple(cands[0]) == (2, 2):
            # wave quantisation: a halo CTA tile is 16 x 16 positions with two M-tiles, 8 x 16 with one. On the small RPN
            # maps the two-tile grid is a little over one wave of the SMs (50 x 100 x 8 sweeps: 224 tiles on 148 SMs), and
            # one-M-tile CTAs — 0.62 of the time each (measured on the Conv3D blocks) — finish sooner: 25 -> 21 us a layer
            rounds = lambda tiles: -(-tiles // self._sm_count)  # noqa: E731
            t2 = self.batch * od * -(-oh // 16) * -(-ow // 16) * n_tiles
            t1 = self.batch * od * -(-oh // 16) * -(-ow // 8) * n_tiles
            if rounds(t1) * 0.62 < rounds(t2) * 1.0:
                cands = [(1, 2)] + list(cands)
        layer, st = None, _native.LISEC_OK
        for m_tiles, group_kh in cands:
            tw, th = (8, 16) if group_kh == 2 else best_tile(oh, ow, m_tiles, 8 if (m_tiles > 1 or group_kh) else 1)
            Wk = np.asarray(W)
            if group_kh == 1:  # the kernel wants the kh taps of one (kd, kw) adjacent: [kd][kw][kh][N][C]
                Wk = Wk.reshape(k[0], k[1], k[2], -1, in_c).transpose(0, 2, 1, 3, 4).reshape(k[0] * k[1] * k[2], -1, in_c)
            desc = _native.lisec_conv_desc(
                batch=self.batch, in_d=in_d, in_h=in_h, in_w=in_w, in_c=in_c, kd=k[0], kh=k[1], kw=k[2],
                stride_d=stride_d, stride_hw=stride_hw, pad_d=pad[0], pad_h=pad[1], pad_w=pad[2], out_c=out_c,
                n_tiles=n_tiles, shuffle=shuffle,
                out_pitch=out_pitch,
                out_ch_off=out_ch_off, relu=relu, out_dtype=out_dtype, tile_w=tw, tile_h=th, m_tiles=m_tiles,
                in_dtype=_native.LISEC_F32 if self.f32 else _native.LISEC_BF16, out_split=out_split,
                group_kh=group_kh, reserved=0)
            if self.f32:  # hi / lo planes of the float64-folded weights
                hi = tf32_round(Wk.astype(np.float32))
                lo = (Wk.astype(np.float64) - hi.astype(np.float64)).astype(np.float32)
                w = torch.from_numpy(np.stack([hi, lo])).to(dev).contiguous()
            else:
                w = torch.from_numpy(np.ascontiguousarray(Wk, dtype=np.float32)).to(dev).to(torch.bfloat16).contiguous()
            layer = _Layer(name, desc, w, sc, sh, src, dst)
            with torch.cuda.device(dev):
                st = self._lib.lisec_conv_plan_create(C.byref(desc), C.c_void_p(src.data_ptr()),
                                                      C.c_void_p(w.data_ptr()), C.c_void_p(sc.data_ptr()),
                                                      C.c_void_p(sh.data_ptr()), C.c_void_p(dst.data_ptr()),
                                                      C.byref(layer.plan))
            if st != _native.LISEC_ERR_BAD_CONFIG:
                break
        if st != _native.LISEC_OK:
            raise _native.LisecError(st, "%s: %s" % (name, self._lib.lisec_conv_last_error().decode()))
        shape = (C.c_int32 * 3)()
        self._lib.lisec_conv_plan_output_shape(layer.plan, shape)
        want = tuple(dst.shape[-4:-1])
        if tuple(shape) != want:
            raise RuntimeError("%s: plan output %s does not match its buffer %s" % (name, tuple(shape), want))
        self.layers.append(layer)

    @property
    def flops(self) -> float:
        """Multiply-add count x 2 of one forward pass over the batch (algorithmic, un-padded)."""
        total = 0.0
        for L in self.layers:
            d = L.desc
            od = (d.in_d + 2 * d.pad_d - d.kd) // d.stride_d + 1
            oh = (d.in_h + 2 * d.pad_h - d.kh) // d.stride_hw + 1
            ow = (d.in_w + 2 * d.pad_w - d.kw) // d.stride_hw + 1
            total += 2.0 * d.batch * od * oh * ow * d.kd * d.kh * d.kw * d.in_c * d.out_c * d.n_tiles
        return total

    def run_layers(self, first: int = 0, last: Optional[int] = None) -> None:
        """Enqueue the plans [first, last) on the current torch stream."""
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            for L in self.layers[first:last]:
                st = self._lib.lisec_conv_plan_run(L.plan, stream)
                if st != _native.LISEC_OK:
                    raise _native.LisecError(st, "%s: %s" % (L.name, self._lib.lisec_conv_last_error().decode()))

    def combine_heads(self) -> None:
        """fuse_heads: heads = the three blocks' folded contributions summed (one launch); otherwise nothing to do."""
        if not self.fuse_heads:
            return
        (c1, s1), (c2, s2), (c3, s3) = self._parts
        assert s1 == 1
        with torch.cuda.device(self.device):
            st = self._lib.lisec_heads_combine(C.c_void_p(c1.data_ptr()), C.c_void_p(c2.data_ptr()), s2,
                                               C.c_void_p(c3.data_ptr()), s3, C.c_void_p(self.heads.data_ptr()),
                                               self.batch, self.nx // 2, self.ny // 2, 16,
                                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if st != _native.LISEC_OK:
            raise _native.LisecError(st, self._lib.lisec_conv_last_error().decode())

    def attach_frontend(self, fe) -> None:
        """the first Conv3D gathers its input straight from `fe`'s sparse output (occupancy map of the
        last voxelize, the float32 voxel rows in self.voxel_feat, c_empty): the dense grid is never written or read. Then
        forward_sparse(points, offsets) replaces fe.forward(..., out=net.grid) + net.forward()."""
        if self.f32:
            raise ValueError("the gather source is a bf16 plan")
        if self.arch.c3 != 64 or fe.c3 != 64:
            raise ValueError("the gather source reads 64-channel voxel rows (createModel as it stands)")
        cv, ce, mv = C.c_void_p(), C.c_void_p(), C.c_int64()
        fe._check(self._lib.lisec_workspace_pointers(fe._h, C.byref(cv), C.byref(ce), C.byref(mv)))
        self.voxel_feat = torch.empty((int(mv.value), 64), dtype=torch.float32, device=self.device)
        st = self._lib.lisec_conv_plan_set_gather(self.layers[0].plan, cv, C.c_void_p(self.voxel_feat.data_ptr()), ce)
        if st != _native.LISEC_OK:
            raise _native.LisecError(st, self._lib.lisec_conv_last_error().decode())
        self._fe = fe

    def forward_sparse(self, points, sweep_offsets) -> Tuple[torch.Tensor, torch.Tensor]:
        """voxelize + VFE rows (no grid) + the network, the first Conv3D reading the sparse rows (attach_frontend)."""
        fe = self._fe
        fe.voxelize(points, sweep_offsets)
        fe.vfe(out=self.voxel_feat)
        self.run_layers()
        self.combine_heads()
        return self.heads[:, 0, :, :, :2], self.heads[:, 0, :, :, 2:]

    @property
    def launches_per_forward(self) -> int:
        return len(self.layers) + (1 if self.fuse_heads else 0) + (1 if self.f32 else 0)

    def forward(self, grid: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """prob [B, nx/2, ny/2, 2], regress [B, nx/2, ny/2, 14] (float32 views of one buffer) from self.grid."""
        if grid is not None and grid.data_ptr() != self.grid.data_ptr():
            self.grid.copy_(grid)
        if self.f32:  # the front end's float32 grid -> the first plan's hi / lo operand planes
            with torch.cuda.device(self.device):
                st = self._lib.lisec_split_tf32(C.c_void_p(self.grid.data_ptr()), C.c_void_p(self.grid_planes[0].data_ptr()),
                                                C.c_void_p(self.grid_planes[1].data_ptr()), self.grid.numel(),
                                                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            if st != _native.LISEC_OK:
                raise _native.LisecError(st, self._lib.lisec_conv_last_error().decode())
        self.run_layers()
        self.combine_heads()
        return self.heads[:, 0, :, :, :2], self.heads[:, 0, :, :, 2:]

    def forward_to_host(self, out_pinned: torch.Tensor) -> None:
        """forward() + the device-to-host copy of the fused head tensor [B,1,nx/2,ny/2,16] (float32: prob = channels 0-1,
        regress = 2-15) into pinned memory, with the copy on a side stream: it runs under the NEXT call's kernels, and
        the next call's head write waits for it. The caller synchronises (self.host_copy_done.synchronize()) before it
        reads out_pinned."""
        if not out_pinned.is_pinned() or out_pinned.dtype != torch.float32 or out_pinned.numel() != self.heads.numel():
            raise ValueError("out_pinned must be pinned float32 with %d elements" % self.heads.numel())
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self.host_copy_done = torch.cuda.Event()
            self._heads_ready = torch.cuda.Event()
        else:
            main.wait_event(self.host_copy_done)  # the previous copy still reads self.heads
        self.forward()
        self._heads_ready.record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._heads_ready)
            out_pinned.view(self.heads.shape).copy_(self.heads, non_blocking=True)
            self.host_copy_done.record(self._copy_stream)

    def close(self) -> None:
        for L in self.layers:
            if L.plan:
                self._lib.lisec_conv_plan_destroy(L.plan)
                L.plan = C.c_void_p()
        self.layers = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

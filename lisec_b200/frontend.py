"""Host-side owner of one liblisec_b200 handle: the VoxelNet front end on one B200.

torch is used for what it is good at here — device buffers, streams, pinned memory — and nothing else: every
computation is a call through the C ABI (lisec_b200/_native.py -> liblisec_b200.so). There is no CPU path.

Reference call sites this object stands behind (see lisec_b200/compat.py for the drop-in signatures):
    VFE_preprocessing(...)                      model_training.py:112   -> Frontend.voxelize()
    sparse.to_dense(...) + model.predict(...)   Predict.py:29-38        -> Frontend.forward()
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _native as N
from . import constants as K
from .weights import Architecture, validate_vfe_pack, vfe_layers

ArrayLike = Union[np.ndarray, torch.Tensor]


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _offsets(sweep_offsets: Sequence[int]) -> np.ndarray:
    off = np.ascontiguousarray(np.asarray(sweep_offsets, dtype=np.int64))
    if off.ndim != 1 or off.size < 2:
        raise ValueError("sweep_offsets must hold n_sweeps+1 entries")
    return off


@dataclass
class VoxelSet:
    """The grouping of one lisec_voxelize() call, in the reference's terms (model_training.py:113-142)."""

    coords: torch.Tensor  # int32 [V,4]  (sweep, z, x, y)
    counts: torch.Tensor  # int32 [V]    len(clusteredPoints[voxel]) before the T cap
    point_idx: torch.Tensor  # int32 [V,T]  kept indices into the sweep's points, ascending, -1 padded
    features: Optional[torch.Tensor]  # float32 [V,T,6]
    n_voxels_per_sweep: np.ndarray
    n_points_in_range: int
    n_dropped_out_of_range: int
    n_dropped_nonfinite: int


class Frontend:
    def __init__(
        self,
        device: int = 0,
        max_points: int = 1_000_000,
        max_sweeps: int = 8,
        grid_dtype: str = "f32",
        voxel_size=(K.voxelx, K.voxely, K.voxelz),
        sample_size: int = K.maxPoints,
        max_voxel=(K.nx // 2, K.ny // 2, K.nz),
        widths=K.vfe_widths,
        post_dense: bool = False,
    ):
        """widths / post_dense: the graph the weights belong to (lisec_b200.weights.Architecture) — (16, 32, 64),
        False is createModel() as it stands (model_training.py:229-235); (16, 64, 128), True is the graph model.png shows."""
        if not torch.cuda.is_available():
            raise RuntimeError("lisec_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = N.load()
        self.device = torch.device("cuda", device)
        self.grid_torch_dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[grid_dtype]
        cfg = N.lisec_config(
            voxel_x=voxel_size[0], voxel_y=voxel_size[1], voxel_z=voxel_size[2],
            sample_size=sample_size,
            max_voxel_x=max_voxel[0], max_voxel_y=max_voxel[1], max_voxel_z=max_voxel[2],
            c1=widths[0], c2=widths[1], c3=widths[2],
            grid_dtype=N.LISEC_F32 if grid_dtype == "f32" else N.LISEC_BF16,
            max_sweeps=max_sweeps, max_points=max_points, device=device, fcn_post_dense=int(bool(post_dense)),
        )
        self.arch = Architecture(int(widths[0]), int(widths[1]), int(widths[2]), bool(post_dense))
        self.cfg = cfg
        self.T = sample_size
        self.c3 = widths[2]
        self.grid_shape = (max_voxel[2], 2 * max_voxel[0], 2 * max_voxel[1])  # (nz, nx, ny)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.init()
            st = self._lib.lisec_create(C.byref(cfg), C.byref(self._h))
        if st != N.LISEC_OK:
            msg = self._lib.lisec_last_error(self._h).decode() if self._h else ""
            self._lib.lisec_destroy(self._h)
            self._h = C.c_void_p()
            raise N.LisecError(st, msg)
        self._keep = None  # device tensors the last voxelize() refers to
        self._n_sweeps = 0
        self._graphs = {}  # forward(graph=True): (points ptr, dtype, offsets, grid ptr) -> [calls seen, CUDAGraph, tensors]

    # ---- lifetime ----------------------------------------------------------------------------------------
    def close(self) -> None:
        self._graphs = {}
        if getattr(self, "_h", None) and self._h:
            self._lib.lisec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status: int) -> None:
        N.check(self._lib, self._h, status)

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.lisec_workspace_bytes(self._h))

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.lisec_last_launch_count(self._h))

    # ---- weights -----------------------------------------------------------------------------------------
    def set_weights(self, pack: dict, bn_epsilon: float = K.bn_epsilon) -> None:
        """pack: Keras-named arrays (lisec_b200.weights); what load_model()/createModel() would hold for the first
        23 layers (model_training.py:229-235) — 26 with the FCNs' second Dense (Architecture.post_dense)."""
        p = validate_vfe_pack(pack, self.arch)
        w = N.lisec_vfe_weights()
        fp = C.POINTER(C.c_float)
        for i, (d, b, post, _, _) in enumerate(vfe_layers(self.arch)):
            w.dense_kernel[i] = p[d + "/kernel"].ctypes.data_as(fp)
            if post:
                w.post_dense_kernel[i] = p[post + "/kernel"].ctypes.data_as(fp)
            w.bn_gamma[i] = p[b + "/gamma"].ctypes.data_as(fp)
            w.bn_beta[i] = p[b + "/beta"].ctypes.data_as(fp)
            w.bn_mean[i] = p[b + "/moving_mean"].ctypes.data_as(fp)
            w.bn_var[i] = p[b + "/moving_variance"].ctypes.data_as(fp)
        w.bn_epsilon = bn_epsilon
        with torch.cuda.device(self.device):
            self._check(self._lib.lisec_set_vfe_weights(self._h, C.byref(w), self._stream()))

    def c_empty(self) -> np.ndarray:
        out = np.empty(self.c3, dtype=np.float32)
        self._check(self._lib.lisec_get_c_empty(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    # ---- inputs ------------------------------------------------------------------------------------------
    def _device_points(self, points: ArrayLike) -> torch.Tensor:
        if isinstance(points, np.ndarray):
            if points.dtype not in (np.float32, np.float64):
                points = points.astype(np.float64)
            points = torch.from_numpy(np.ascontiguousarray(points)).to(self.device)
        if points.device != self.device:
            points = points.to(self.device)
        if points.dtype not in (torch.float32, torch.float64):
            points = points.to(torch.float64)
        if points.dim() != 2 or points.shape[1] != 3:
            raise ValueError("points must be (n,3), got %s" % (tuple(points.shape),))
        return points.contiguous()

    @staticmethod
    def _dtype_code(t) -> int:
        return N.LISEC_F32 if t in (torch.float32, np.dtype("float32")) else N.LISEC_F64

    # ---- a1-a3 -------------------------------------------------------------------------------------------
    def voxelize(self, points: ArrayLike, sweep_offsets: Optional[Sequence[int]] = None) -> None:
        pts = self._device_points(points)
        off = _offsets([0, pts.shape[0]] if sweep_offsets is None else sweep_offsets)
        if off[-1] != pts.shape[0]:
            raise ValueError("sweep_offsets[-1] = %d but %d points were given" % (off[-1], pts.shape[0]))
        with torch.cuda.device(self.device):
            self._check(
                self._lib.lisec_voxelize(self._h, _ptr(pts), self._dtype_code(pts.dtype),
                                         off.ctypes.data_as(C.POINTER(C.c_int64)), len(off) - 1, self._stream())
            )
        self._keep = pts
        self._n_sweeps = len(off) - 1

    def counts(self):
        per = np.zeros(self._n_sweeps, dtype=np.int32)
        nv, nin, noor, nnf = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        with torch.cuda.device(self.device):
            self._check(
                self._lib.lisec_voxel_counts(self._h, per.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nv),
                                             C.byref(nin), C.byref(noor), C.byref(nnf), self._stream())
            )
        return per, nv.value, nin.value, noor.value, nnf.value

    COUNTS_BYTES = 8 * 8 + 4 * (N.LISEC_MAX_SWEEPS + 1)

    def counts_async(self, pinned: torch.Tensor) -> None:
        """Enqueue the totals read-back into a pinned uint8 tensor (>= COUNTS_BYTES); decode with decode_counts()
        once the stream (or an event recorded after this call) has completed."""
        if not pinned.is_pinned() or pinned.dtype != torch.uint8 or pinned.numel() < self.COUNTS_BYTES:
            raise ValueError("counts_async() wants a pinned uint8 tensor of at least %d bytes" % self.COUNTS_BYTES)
        with torch.cuda.device(self.device):
            self._check(self._lib.lisec_voxel_counts_async(self._h, C.c_void_p(pinned.data_ptr()), pinned.numel(),
                                                           self._stream()))

    def decode_counts(self, pinned: torch.Tensor, n_sweeps: Optional[int] = None):
        n = self._n_sweeps if n_sweeps is None else n_sweeps
        raw = pinned.numpy()
        tot = raw[:64].view(np.int64)
        svs = raw[64:64 + 4 * (n + 1)].view(np.int32)
        return np.diff(svs), int(tot[0]), int(tot[1]), int(tot[5]), int(tot[4])

    def export(self, features: bool = True) -> VoxelSet:
        per, V, nin, noor, nnf = self.counts()
        dev = self.device
        coords = torch.empty((V, 4), dtype=torch.int32, device=dev)
        cnt = torch.empty((V,), dtype=torch.int32, device=dev)
        pidx = torch.empty((V, self.T), dtype=torch.int32, device=dev)
        feat = torch.empty((V, self.T, 6), dtype=torch.float32, device=dev) if features else None
        with torch.cuda.device(dev):
            self._check(
                self._lib.lisec_voxels_export(self._h, _ptr(coords), _ptr(cnt), _ptr(pidx), _ptr(feat),
                                              self._stream())
            )
        return VoxelSet(coords, cnt, pidx, feat, per, nin, noor, nnf)

    def emit_dense_input(self) -> torch.Tensor:
        """[n_sweeps,nz,nx,ny,T,6] float32 — the reference's model input (tests and tiny grids only)."""
        nz, nx, ny = self.grid_shape
        dense = torch.empty((self._n_sweeps, nz, nx, ny, self.T, 6), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.lisec_emit_dense_input(self._h, _ptr(dense), self._stream()))
        return dense

    # ---- a6-a11 ------------------------------------------------------------------------------------------
    def vfe(self, n_voxels: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            if n_voxels is None:
                n_voxels = self.counts()[1]
            out = torch.empty((n_voxels, self.c3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.lisec_vfe_forward(self._h, _ptr(out), self._stream()))
        return out

    def new_grid(self, n_sweeps: int) -> torch.Tensor:
        nz, nx, ny = self.grid_shape
        return torch.empty((n_sweeps, nz, nx, ny, self.c3), dtype=self.grid_torch_dtype, device=self.device)

    def scatter(self, voxel_feat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = self.new_grid(self._n_sweeps)
        with torch.cuda.device(self.device):
            self._check(self._lib.lisec_scatter_dense(self._h, _ptr(voxel_feat), _ptr(out), self._stream()))
        return out

    @property
    def last_fused_kernel_ms(self) -> float:
        """Device time of the fused VFE + grid kernel inside the last fused call (CUDA events on its stream)."""
        ms = C.c_float(0.0)
        with torch.cuda.device(self.device):
            self._check(self._lib.lisec_last_fused_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def vfe_scatter_fused(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """VFE + dense grid in one kernel, on the grouping of the last voxelize()."""
        if out is None:
            out = self.new_grid(self._n_sweeps)
        with torch.cuda.device(self.device):
            self._check(self._lib.lisec_vfe_scatter_fused(self._h, _ptr(out), self._stream()))
        return out

    def forward(self, points: ArrayLike, sweep_offsets: Optional[Sequence[int]] = None,
                out: Optional[torch.Tensor] = None, graph: bool = False) -> torch.Tensor:
        """points already on the device -> dense grid [n_sweeps,nz,nx,ny,c3]; no host round trip.

        graph=True: for callers that hand in the SAME device buffers again and again (a staging buffer, a ring of them):
        the second call with one (points buffer, offsets, grid buffer) captures the step's six kernels into a CUDA graph,
        later calls replay it — the same kernels on the same buffers, with the GPU's launch work done once (a step of 8 x
        100 k points: 0.338 -> 0.327 ms, tools/graph_probe.py). The kernels keep no host-side state between calls (every
        counter is put back by the kernels themselves), which is what makes the replay legal
        (tests/test_gpu_call_state.py). The buffers' CONTENTS may change between calls, their addresses and the offsets may not."""
        pts = self._device_points(points)
        off = _offsets([0, pts.shape[0]] if sweep_offsets is None else sweep_offsets)
        n = len(off) - 1
        if out is None:
            out = self.new_grid(n)
        if graph:
            key = (pts.data_ptr(), pts.dtype, off.tobytes(), out.data_ptr())
            entry = self._graphs.setdefault(key, [0, None, (pts, out)])
            entry[0] += 1
            if entry[1] is None and entry[0] == 2:
                g = torch.cuda.CUDAGraph()
                # (thread-local capture mode: other threads of the process — an NCCL watchdog — may call CUDA meanwhile)
                with torch.cuda.device(self.device), torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self.forward(pts, off, out=out)
                entry[1] = g
            if entry[1] is not None:
                with torch.cuda.device(self.device):
                    entry[1].replay()
                self._keep = pts
                self._n_sweeps = n
                return out
        with torch.cuda.device(self.device):
            self._check(
                self._lib.lisec_frontend_forward(self._h, _ptr(pts), self._dtype_code(pts.dtype),
                                                 off.ctypes.data_as(C.POINTER(C.c_int64)), n, _ptr(out),
                                                 self._stream())
            )
        self._keep = pts
        self._n_sweeps = n
        return out

    def forward_host(self, points_host: ArrayLike, sweep_offsets: Sequence[int],
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """points in HOST memory (numpy array or pinned CPU tensor): the library copies them in on the stream."""
        if isinstance(points_host, torch.Tensor):
            if points_host.is_cuda:
                raise ValueError("forward_host() wants host memory")
            code = self._dtype_code(points_host.dtype)
            ptr, n_pts, keep = points_host.data_ptr(), points_host.shape[0], points_host
            if not points_host.is_contiguous() or points_host.dtype not in (torch.float32, torch.float64):
                raise ValueError("host points must be contiguous float32/float64 (n,3)")
        else:
            a = np.ascontiguousarray(points_host)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            code = self._dtype_code(a.dtype)
            ptr, n_pts, keep = a.ctypes.data, a.shape[0], a
        off = _offsets(sweep_offsets)
        if off[-1] != n_pts:
            raise ValueError("sweep_offsets[-1] = %d but %d points were given" % (off[-1], n_pts))
        n = len(off) - 1
        if out is None:
            out = self.new_grid(n)
        with torch.cuda.device(self.device):
            self._check(
                self._lib.lisec_frontend_forward_host(self._h, C.c_void_p(ptr), code,
                                                      off.ctypes.data_as(C.POINTER(C.c_int64)), n, _ptr(out),
                                                      self._stream())
            )
        self._keep = keep
        self._n_sweeps = n
        return out

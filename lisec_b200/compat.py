"""Drop-in for the reference's call sites on the front-end path — same names, argument meaning and shapes.

    reference (bot15498/Lisec)                                         here
    ------------------------------------------------------------------ --------------------------------------------
    VFE_preprocessing(points, xSize, ySize, zSize, sampleSize,         VFE_preprocessing(...)  -> SparseVoxelTensor
                      maxVoxelX, maxVoxelY, maxVoxelZ)                     .indices [nnz,5] (z,x,y,i,j)  .values
        model_training.py:112, callers Predict.py:21-28,                   .dense_shape  .shape
        model_training.py:270-277, 314-321
    sparse.reshape(t, (1,) + t.shape)          Predict.py:29           sparse.reshape(t, shape)
    sparse.to_dense(t, default_value=0.,       Predict.py:30,          sparse.to_dense(t, ...)  -> DenseVoxelInput
                    validate_indices=False)    model_training.py:279
    tf.stack(points, axis=0)                   model_training.py:285   stack(list, axis=0)
    createModel(nx, ny, nz, maxPoints)         model_training.py:222   createModel(nx, ny, nz, maxPoints, weights=)
    load_model(path, custom_objects={...})     Predict.py:51-52        load_model(path, custom_objects=None)
    model.predict(x)                           Predict.py:38           model.predict(x) -> [prob, regress]  (bf16 tensor-core
                                                                           middle + RPN, lisec_b200/network.py), and
                                                                       model.predict_voxel_grid(x)  (first 23 layers only)

What is different underneath: nothing is densified. SparseVoxelTensor and DenseVoxelInput are lazy handles on the
sweep's points; the 537.6 MB-per-sweep dense input (README.md:58-64 of the reference) only exists if somebody asks
for it (`.numpy()` on a tiny grid). `predict_voxel_grid` sends all stacked sweeps through ONE
lisec_frontend_forward call and returns the [N, nz, nx, ny, 64] tensor the reference's first Conv3D consumes
(model_training.py:235-236), resident on the GPU.

Sampling: the reference subsamples over-full voxels with the unseeded global RNG (np.random.choice,
model_training.py:132); here the kept rows are the first `sampleSize` in point order.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import constants as K
from .frontend import Frontend
from .weights import CURRENT, Architecture, detect_architecture, load_npz, synthetic_model_pack

_FRONTENDS: dict = {}


def _frontend(cfg_key, n_points: int, n_sweeps: int, device: int = 0, grid_dtype: str = "f32",
              arch: Architecture = CURRENT) -> Frontend:
    """One Frontend per (geometry, device, grid dtype, graph); re-created with doubled capacity when a call outgrows it."""
    key = (cfg_key, device, grid_dtype, arch)
    fe = _FRONTENDS.get(key)
    need_pts, need_sw = max(n_points, 1), max(n_sweeps, 1)
    if fe is None or fe.cfg.max_points < need_pts or fe.cfg.max_sweeps < need_sw:
        xs, ys, zs, T, mx, my, mz = cfg_key
        cap_pts = max(need_pts, 2 * (fe.cfg.max_points if fe else 0), 262_144)
        cap_sw = min(64, max(need_sw, 2 * (fe.cfg.max_sweeps if fe else 0), 8))
        weights = getattr(fe, "_pack", None)
        if fe is not None:
            fe.close()
        fe = Frontend(device=device, max_points=cap_pts, max_sweeps=cap_sw, voxel_size=(xs, ys, zs),
                      sample_size=T, max_voxel=(mx, my, mz), grid_dtype=grid_dtype, widths=arch.widths,
                      post_dense=arch.post_dense)
        if weights is not None:
            fe.set_weights(weights)
            fe._pack = weights
        _FRONTENDS[key] = fe
    return fe


class SparseVoxelTensor:
    """Stands where tf.SparseTensor stood (model_training.py:151-152). Lazy: holds the sweep's points."""

    def __init__(self, points, cfg_key):
        self._points = np.ascontiguousarray(points)
        if self._points.dtype not in (np.float32, np.float64):
            self._points = self._points.astype(np.float64)
        self._cfg = cfg_key
        xs, ys, zs, T, mx, my, mz = cfg_key
        self.dense_shape = [mz, mx * 2, my * 2, T, 6]
        self.shape = tuple(self.dense_shape)
        self._coo = None

    def _materialise(self):
        if self._coo is None:
            fe = _frontend(self._cfg, len(self._points), 1)
            fe.voxelize(self._points, [0, len(self._points)])
            vs = fe.export()
            T = self._cfg[3]
            first = vs.point_idx[:, 0].cpu().numpy()
            order = np.argsort(first, kind="stable")  # the reference's dict order: first appearance (:123-126)
            coords = vs.coords.cpu().numpy()[order][:, 1:]
            feats = vs.features.cpu().numpy()[order]
            V = len(order)
            ii, jj = np.meshgrid(np.arange(T), np.arange(6), indexing="ij")
            ind = np.empty((V, T, 6, 5), dtype=np.int64)
            for k in range(3):
                ind[..., k] = coords[:, k, None, None]
            ind[..., 3], ind[..., 4] = ii[None], jj[None]
            self._coo = (ind.reshape(-1, 5), feats.reshape(-1))
        return self._coo

    @property
    def indices(self):
        """[nnz,5] (z,x,y,i,j): 6*T entries per occupied voxel, zero pad rows included (model_training.py:143-149)."""
        return self._materialise()[0]

    @property
    def values(self):
        """float32: the reference's float64 values after the Keras input cast."""
        return self._materialise()[1]


def VFE_preprocessing(points, xSize, ySize, zSize, sampleSize, maxVoxelX, maxVoxelY, maxVoxelZ) -> SparseVoxelTensor:
    return SparseVoxelTensor(points, (float(xSize), float(ySize), float(zSize), int(sampleSize), int(maxVoxelX),
                                      int(maxVoxelY), int(maxVoxelZ)))


class DenseVoxelInput:
    """Stands where the dense [N, nz, nx, ny, T, 6] tensor stood. Lazy: a list of sweeps."""

    def __init__(self, sweeps: List[SparseVoxelTensor], batched: bool):
        self.sweeps = sweeps
        self.batched = batched
        s = tuple(sweeps[0].shape)
        self.shape = ((len(sweeps),) + s) if batched else s

    def numpy(self) -> np.ndarray:
        """The dense tensor itself (float32) — only sensible on small grids."""
        cfg = self.sweeps[0]._cfg
        pts = np.concatenate([s._points.astype(np.float64) for s in self.sweeps])
        off = np.cumsum([0] + [len(s._points) for s in self.sweeps])
        fe = _frontend(cfg, len(pts), len(self.sweeps))
        fe.voxelize(pts, off)
        d = fe.emit_dense_input().cpu().numpy()
        return d if self.batched else d[0]


class sparse:  # noqa: N801 - mirrors `from tensorflow import sparse`
    @staticmethod
    def reshape(t: SparseVoxelTensor, shape: Sequence[int]) -> "DenseVoxelInput | SparseVoxelTensor":
        if tuple(shape) == tuple(t.shape):
            return t
        if tuple(shape) == (1,) + tuple(t.shape):  # Predict.py:29
            b = _BatchedSparse([t])
            return b
        raise ValueError("only the (1,)+shape reshape of Predict.py:29 is supported, got %s" % (tuple(shape),))

    @staticmethod
    def to_dense(t, default_value=0.0, validate_indices=False) -> DenseVoxelInput:
        if default_value != 0.0:
            raise ValueError("the reference only densifies with default_value=0. (model_training.py:279)")
        if isinstance(t, _BatchedSparse):
            return DenseVoxelInput(t.sweeps, batched=True)
        return DenseVoxelInput([t], batched=False)


class _BatchedSparse:
    def __init__(self, sweeps):
        self.sweeps = sweeps
        self.shape = (len(sweeps),) + tuple(sweeps[0].shape)
        self.dense_shape = list(self.shape)


def stack(values: Sequence[DenseVoxelInput], axis: int = 0) -> DenseVoxelInput:
    """tf.stack(points, axis=0) of per-sample dense tensors (model_training.py:285)."""
    if axis != 0:
        raise ValueError("the reference stacks on axis 0 only")
    sweeps = []
    for v in values:
        if v.batched:
            raise ValueError("stack() takes un-batched per-sample tensors")
        sweeps.extend(v.sweeps)
    return DenseVoxelInput(sweeps, batched=True)


class RepeatLayer:  # model_training.py:32-40 — implicit in the fused kernel; kept so custom_objects={...} resolves
    pass


class MaxPoolingVFELayer:  # model_training.py:44-61
    def __init__(self, combine=False, **kwargs):
        self.combineDim = combine


class VoxelNetFrontEnd:
    """createModel (model_training.py:222-257) with its weights: the first 23 Keras layers (VFE stack, :229-235) as the
    fused front-end kernels, the rest (:236-256) as tensor-core convolution plans. `arch` names the graph the weights
    belong to (lisec_b200.weights.Architecture; None: read it off the kernels' shapes, as load_model() reads it off the
    .h5's own model_config): createModel() as it stands, or the older graph model.png shows."""

    def __init__(self, nx, ny, nz, maxPoints, pack: dict, arch: Optional[Architecture] = None):
        self.grid = (nz, nx, ny)
        self.maxPoints = maxPoints
        self.pack = pack
        self.arch = arch if arch is not None else detect_architecture(pack)
        self._nets: dict = {}

    def _run_frontend(self, x: DenseVoxelInput, device: int, grid_dtype: str, out=None) -> torch.Tensor:
        if not isinstance(x, DenseVoxelInput):
            raise TypeError("expected the output of sparse.to_dense()/stack(); a materialised dense array would be "
                            "the 500 GB path this library exists to avoid")
        cfg = x.sweeps[0]._cfg
        nz, nx, ny = self.grid
        if (cfg[6], 2 * cfg[4], 2 * cfg[5]) != (nz, nx, ny) or cfg[3] != self.maxPoints:
            raise ValueError("input shape %s does not match the model's InputVoxel (%d,%d,%d,%d,6)"
                             % (x.shape, nz, nx, ny, self.maxPoints))
        dts = {s._points.dtype for s in x.sweeps}
        dt = np.float32 if dts == {np.dtype("float32")} else np.float64
        pts = np.concatenate([s._points.astype(dt, copy=False) for s in x.sweeps])
        off = np.cumsum([0] + [len(s._points) for s in x.sweeps]).astype(np.int64)
        fe = _frontend(cfg, len(pts), len(x.sweeps), device, grid_dtype, self.arch)
        if getattr(fe, "_pack", None) is not self.pack:
            fe.set_weights(self.pack)
            fe._pack = self.pack
        return fe.forward_host(pts, off, out=out)

    def predict_voxel_grid(self, x: DenseVoxelInput, device: int = 0) -> torch.Tensor:
        """The float32 [N, nz, nx, ny, C3] tensor the reference's first Conv3D consumes (:235-236), on the GPU."""
        return self._run_frontend(x, device, "f32")

    def predict(self, x: DenseVoxelInput, device: int = 0, dtype: str = "bf16") -> list:
        """model.predict(x) (Predict.py:38): [prob (N, nx/2, ny/2, 2), regress (N, nx/2, ny/2, 14)] float32 numpy arrays.
        The front end writes its grid straight into the dense network's input buffer; the middle Conv3D stack, RPN and
        heads run as tensor-core plans (lisec_b200/network.py): dtype="bf16" (bf16 operands, float32 accumulation; 2e-2
        of the reference) or dtype="f32" (3xTF32 on float32 hi/lo planes; 1e-5 of the reference, ~7x slower)."""
        from .network import DenseNetwork

        n = len(x.sweeps) if isinstance(x, DenseVoxelInput) else 0
        key = (n, device, dtype)
        net = self._nets.get(key)
        if net is None:
            nz, nx, ny = self.grid
            net = self._nets[key] = DenseNetwork(self.pack, batch=n, nx=nx, ny=ny, nz=nz, device=device, dtype=dtype,
                                                 arch=self.arch)
        self._run_frontend(x, device, "f32" if dtype == "f32" else "bf16", out=net.grid)
        prob, reg = net.forward()
        return [prob.contiguous().cpu().numpy(), reg.contiguous().cpu().numpy()]


def predictMain(samples, outPath, level5Data, model, combine_lidar_data=None, dtype: str = "bf16", batch: int = 8,
                dataDir: Optional[str] = None):
    """Predict.predictMain(samples, outPath, level5Data, model) (Predict.py:9-40): for every sample write
    `sample{i}_label.npy` (prob, (1, nx/2, ny/2, 2)) and `sample{i}_regress.npy` ((1, nx/2, ny/2, 14)) into outPath.

    The reference handles one sample per model.predict call; here up to `batch` samples share one pass (sweeps are
    independent, the files are the same). The points come from `combine_lidar_data(sample, dataDir, level5Data)`
    (model_training.py:73-98): by default lisec_b200.ingest's GPU version (the sensor files under `dataDir`, which the
    reference hard-codes at Predict.py:12 / Constants.py:4; `level5Data` only needs the SDK's `.get(table, token)`), or
    any callable of that signature returning the (n, 3) float array."""
    import os

    if combine_lidar_data is None:
        if dataDir is None:
            raise ValueError("dataDir: the Lyft dataset directory (the reference's Constants.lyft_data_dir)")
        from .ingest import combine_lidar_data
    cfg = (K.voxelx, K.voxely, K.voxelz, K.maxPoints, K.nx // 2, K.ny // 2, K.nz)
    for i0 in range(0, len(samples), batch):
        dense = []
        for sample in samples[i0:i0 + batch]:
            pts = combine_lidar_data(sample, dataDir, level5Data)
            t = VFE_preprocessing(pts, *cfg)                        # Predict.py:21-28
            t = sparse.reshape(t, (1,) + tuple(t.shape))            # Predict.py:29
            dense.append(sparse.to_dense(t, default_value=0., validate_indices=False))  # Predict.py:30
        x = DenseVoxelInput([s for d in dense for s in d.sweeps], batched=True)
        prob, regress = model.predict(x, dtype=dtype)               # Predict.py:38
        for j in range(len(dense)):
            np.save(os.path.join(outPath, "sample%d_label.npy" % (i0 + j)), prob[j:j + 1])      # Predict.py:39
            np.save(os.path.join(outPath, "sample%d_regress.npy" % (i0 + j)), regress[j:j + 1])  # Predict.py:40


def save_model(pack: dict, save_path: str) -> None:
    """model.save(save_path) (model_training.py:302, 346): `.h5` in Keras's weight layout (lisec_b200/h5write.py), anything
    else as an .npz weight pack. Either loads back with load_model()."""
    if str(save_path).lower().endswith((".h5", ".hdf5")):
        from .h5write import write_keras_weights

        write_keras_weights(save_path, pack)
    else:
        from .weights import save_npz

        save_npz(save_path, pack)


def _fit(pack, samples, level5Data, save_path, labels_dir, steps_per_epoch, batch_size, dataDir, combine_lidar_data, device,
         labels):
    import os

    from .train import TrainStep

    if combine_lidar_data is None:
        if dataDir is None:
            raise ValueError("dataDir: the Lyft dataset directory (the reference's Constants.lyft_data_dir)")
        from .ingest import combine_lidar_data
    clouds = [np.ascontiguousarray(combine_lidar_data(s, dataDir, level5Data)) for s in samples]  # :266-283
    if labels is None:  # :288-290 (the reference spells the path with a Windows separator)
        def load(name):
            for cand in (os.path.join(labels_dir, name), labels_dir + "\\" + name):
                if os.path.exists(cand):
                    return np.load(cand, allow_pickle=True)
            raise FileNotFoundError(os.path.join(labels_dir, name))
        labels = (load("labelsClass.npy"), load("regressClass.npy"))
    outClass, outRegress = (np.asarray(a, dtype=np.float32) for a in labels)
    if len(outClass) < len(clouds) or len(outRegress) < len(clouds):
        raise ValueError("%d samples but %d / %d label entries" % (len(clouds), len(outClass), len(outRegress)))
    n = len(clouds)
    step = TrainStep(pack, batch=batch_size, max_points=batch_size * max(len(c) for c in clouds), device=device,
                     nx=K.nx, ny=K.ny, nz=K.nz)
    dev = torch.device("cuda", device)
    yc, yr = torch.from_numpy(outClass).to(dev), torch.from_numpy(outRegress).to(dev)
    history = []
    for it in range(steps_per_epoch):  # fit(batch_size=1, epochs=1, steps_per_epoch=180) (:299)
        idx = [(it * batch_size + b) % n for b in range(batch_size)]
        pts = np.concatenate([clouds[i] for i in idx])
        off = np.cumsum([0] + [len(clouds[i]) for i in idx]).tolist()
        loss = step.step(pts, off, yc[idx].contiguous(), yr[idx].contiguous())
        history.append(loss)
    history = [float(l) for l in history]  # one host read-back at the end
    out = step.to_pack()
    step.close()
    save_model(out, save_path)  # :302
    return {"loss": history}


def train(samples, level5Data, save_path, labels_dir="labels3", steps_per_epoch=180, batch_size=1, dataDir=None,
          combine_lidar_data=None, device=0, seed=0, labels=None):
    """model_training.train(samples, level5Data, save_path) (model_training.py:260-302): voxelize every sample, load the
    labels (`labels_dir`/labelsClass.npy, regressClass.npy, or `labels=(cls, reg)`), createModel(), SGD(lr=0.01, decay=1e-6,
    momentum=0.9, nesterov=True), loss=['mse','mse'], fit(batch_size=1, epochs=1, steps_per_epoch=180), model.save.
    Every step runs on the GPU (lisec_b200/train.py: TrainStep): the VFE stack with batch statistics, the dense network
    as bf16 tensor-core plans with float32 master weights, both backward passes, the Keras update. The samples are not
    densified (the reference stacks 1.1 GB per sample at :279-285); the step walks them one per iteration, cyclically.
    Returns history.history (the per-step losses), which the reference prints (:301)."""
    from .weights import keras_default_init_pack

    return _fit(keras_default_init_pack(seed), samples, level5Data, save_path, labels_dir, steps_per_epoch, batch_size,
                dataDir, combine_lidar_data, device, labels)


def train_with_model(samples, level5Data, model_path, save_path, labels_dir="labels3", steps_per_epoch=180, batch_size=1,
                     dataDir=None, combine_lidar_data=None, device=0, labels=None):
    """model_training.train_with_model (model_training.py:305-346): as train(), starting from load_model(model_path)."""
    model = load_model(model_path)
    if model.arch != CURRENT:
        raise ValueError("train_with_model: %s holds the graph %s; the training step is built for the graph train() creates "
                         "(createModel, model_training.py:222-257: %s) — inference (predictMain) runs either"
                         % (model_path, model.arch, CURRENT))
    return _fit(model.pack, samples, level5Data, save_path, labels_dir, steps_per_epoch, batch_size, dataDir,
                combine_lidar_data, device, labels)


def createModel(nx=K.nx, ny=K.ny, nz=K.nz, maxPoints=K.maxPoints, weights: Optional[dict] = None, seed: int = 0,
                arch: Optional[Architecture] = None):
    """model_training.py:222. Keras would random-initialise; `weights` (Keras-named arrays) or a seeded synthetic
    pack stands in. arch: None = the graph createModel() builds today (or, with `weights`, the one they were saved from)."""
    if weights is None:
        weights = synthetic_model_pack(seed, arch or CURRENT)
    return VoxelNetFrontEnd(nx, ny, nz, maxPoints, weights, arch)


def load_model(path: str, custom_objects: Optional[dict] = None, nx=K.nx, ny=K.ny, nz=K.nz, maxPoints=K.maxPoints):
    """Predict.py:51-52 / model_training.py:337-338. `path`: the Keras .h5 file model.save() wrote (read by
    lisec_b200/h5weights.py — a reader of the HDF5 format itself, h5py is not in this image) or a .npz weight pack with
    the same Keras weight names (lisec_b200/weights.py). custom_objects is accepted and ignored: RepeatLayer and
    MaxPoolingVFELayer are part of the fused VFE kernel."""
    if str(path).lower().endswith((".h5", ".hdf5", ".keras.h5")):
        from .h5weights import read_keras_weights

        return VoxelNetFrontEnd(nx, ny, nz, maxPoints, read_keras_weights(path))
    return VoxelNetFrontEnd(nx, ny, nz, maxPoints, load_npz(path))

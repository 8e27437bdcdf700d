"""Both graphs the reference's .h5 may hold: createModel() as it stands and the older graph model.png shows
(the lines commented out at model_training.py:172 and :205 switched on; VFE widths 16 | 64 | 128). The float32 VFE kernel
that serves every graph but the current one (lisec_b200/csrc/vfe_generic.cu) against the float64 oracle at north_star's
float32 bar, against the tensor-core kernel on the graph they share, and the whole model.predict() of the older graph
against the whole CPU oracle. Parity unpinned like every floating-point half of this path (no TensorFlow, no .h5 here)."""
import os

import numpy as np
import pytest
import torch

from lisec_b200 import synth
from lisec_b200.weights import CURRENT, MODEL_PNG, Architecture, synthetic_model_pack, synthetic_vfe_pack
from oracle import lisec_oracle as O
from oracle import network_oracle as NO

pytestmark = pytest.mark.gpu
REF = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)
T = 35


def within(got, ref):
    """max |err| / max(|ref|, rms(ref)): north_star's relative bar with the denominator floored at the tensor's rms (the
    outputs are ReLU outputs: exact zeros and values near zero are common)."""
    ref = np.asarray(ref, dtype=np.float64)
    floor = np.sqrt(np.mean(ref * ref))
    return float((np.abs(np.asarray(got, dtype=np.float64) - ref) / np.maximum(np.abs(ref), floor)).max())


def cloud():
    a = synth.lyft_like_sweep(30_000, seed=7)
    rng = np.random.default_rng(5)
    full = rng.uniform([2.0, 1.0, 0.5], [2.49, 1.24, 0.74], size=(80, 3)).astype(np.float32)  # one voxel far above T
    tiny = np.asarray([[-1e-30, 0.3, 0.6], [1.3, -1e-30, 0.6]], np.float32)                    # the floor's edge cases
    return np.concatenate([a, full, tiny])


def make_frontend(arch, generic_env=False, **kw):
    from lisec_b200 import Frontend

    old = os.environ.get("LISEC_GENERIC_VFE")
    if generic_env:
        os.environ["LISEC_GENERIC_VFE"] = "1"  # read by lisec_create()
    try:
        return Frontend(device=0, widths=arch.widths, post_dense=arch.post_dense, **kw)
    finally:
        if generic_env:
            if old is None:
                del os.environ["LISEC_GENERIC_VFE"]
            else:
                os.environ["LISEC_GENERIC_VFE"] = old


ARCHS = [CURRENT, MODEL_PNG, Architecture(16, 32, 64, True), Architecture(16, 64, 128, False)]


@pytest.mark.parametrize("arch", ARCHS, ids=lambda a: "%d-%d-%d-%s" % (a.c1, a.c2, a.c3, "post" if a.post_dense else "relu"))
@pytest.mark.parametrize("seed", [0, 4])
def test_float32_vfe_kernel_matches_the_oracle_on_every_graph(arch, seed):
    pack = synthetic_vfe_pack(seed, arch)
    pts = cloud()
    vox = O.voxelize_np(pts, **REF)
    ref = O.vfe_forward(vox["features"].astype(np.float32), pack, np.float64, post_dense=arch.post_dense)
    assert ref.shape == (len(vox["counts"]), arch.c3)
    fe = make_frontend(arch, generic_env=True, max_points=len(pts), max_sweeps=1, grid_dtype="f32")
    fe.set_weights(pack)
    want_empty = O.c_empty(pack, T)
    assert within(fe.c_empty(), want_empty) <= 1e-5
    fe.voxelize(pts, [0, len(pts)])
    rows = fe.vfe().cpu().numpy()
    assert rows.shape == ref.shape
    e = within(rows, ref)
    assert e <= 1e-5, e
    # the grid: every cell written once — the voxel's row where the occupancy map says so, c_empty elsewhere
    grid = torch.full((1, 8, 200, 400, arch.c3), float("nan"), dtype=torch.float32, device="cuda")
    fe.forward(pts, [0, len(pts)], out=grid)
    g = grid.cpu().numpy().reshape(-1, arch.c3)
    assert np.array_equal(g[vox["linear"]], rows)
    mask = np.ones(len(g), bool)
    mask[vox["linear"]] = False
    assert (g[mask] == fe.c_empty()).all()
    fe.close()
    # bf16 grid: the same rows rounded once
    fb = make_frontend(arch, generic_env=True, max_points=len(pts), max_sweeps=1, grid_dtype="bf16")
    fb.set_weights(pack)
    gb = fb.forward(pts, [0, len(pts)]).float().cpu().numpy().reshape(-1, arch.c3)
    want = torch.from_numpy(rows).to(torch.bfloat16).float().numpy()
    assert np.array_equal(gb[vox["linear"]], want)
    assert (gb[mask] == torch.from_numpy(fb.c_empty()).to(torch.bfloat16).float().numpy()).all()
    fb.close()


def test_older_graph_on_float64_points_ragged_sweeps_and_another_sample_size():
    """The float32 kernel away from the defaults: float64 points (the reference's own dtype, model_training.py:93-94), three
    sweeps of which one is empty, T = 20 on a reduced grid — voxel rows, c_empty and the grid against the oracle."""
    arch, Tn, mx, my, mz = MODEL_PNG, 20, 16, 24, 4
    ref = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=Tn, maxVoxelX=mx, maxVoxelY=my, maxVoxelZ=mz)
    rng = np.random.default_rng(9)
    a = np.stack([rng.uniform(-8.5, 8.5, 9000), rng.uniform(-6.5, 6.5, 9000), rng.uniform(-0.2, 1.2, 9000)], axis=1)
    a[:3000, :2] *= 0.1  # a dense core: voxels far past T
    b = np.stack([rng.uniform(-3, 3, 700), rng.uniform(-2, 2, 700), rng.uniform(0, 1, 700)], axis=1)
    sweeps = [a, np.zeros((0, 3)), b]
    pts = np.concatenate(sweeps)
    off = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
    pack = synthetic_vfe_pack(7, arch)
    fe = make_frontend(arch, max_points=len(pts), max_sweeps=3, grid_dtype="f32", max_voxel=(mx, my, mz), sample_size=Tn)
    fe.set_weights(pack)
    grid = fe.forward(pts, off).cpu().numpy()
    assert grid.shape == (3, mz, 2 * mx, 2 * my, arch.c3) and grid.dtype == np.float32
    ce = O.c_empty(pack, Tn)
    assert within(fe.c_empty(), ce) <= 1e-5
    for s, p in enumerate(sweeps):
        vox = O.voxelize_np(p, **ref)
        want = O.scatter_dense(vox["coords"], O.vfe_forward(vox["features"].astype(np.float32), pack), ce,
                               (mz, 2 * mx, 2 * my), dtype=np.float64)
        e = within(grid[s], want)
        assert e <= 1e-5, (s, e)
    assert (grid[1] == fe.c_empty()).all()  # the empty sweep: background only
    fe.close()


def test_float32_and_tensor_core_vfe_kernels_agree_on_the_graph_they_share():
    pack = synthetic_vfe_pack(1)
    sweeps = [synth.lyft_like_sweep(40_000, seed=2), synth.lyft_like_sweep(25_000, seed=3)]
    pts = np.concatenate(sweeps)
    off = [0, len(sweeps[0]), len(pts)]
    out = []
    for generic in (False, True):
        fe = make_frontend(CURRENT, generic_env=generic, max_points=len(pts), max_sweeps=2, grid_dtype="f32")
        fe.set_weights(pack)
        fe.voxelize(pts, off)
        out.append((fe.vfe().cpu().numpy(), fe.c_empty(), fe.last_launch_count))
        fe.close()
    (a, ea, _), (b, eb, _) = out
    assert a.shape == b.shape and within(a, b) <= 2e-5 and within(ea, eb) <= 2e-5


def test_training_is_refused_for_the_older_graph():
    from lisec_b200 import _native

    arch = MODEL_PNG
    pts = synth.lyft_like_sweep(2_000, seed=1)
    fe = make_frontend(arch, max_points=len(pts), max_sweeps=1, grid_dtype="f32")
    fe.set_weights(synthetic_vfe_pack(0, arch))
    fe.voxelize(pts, [0, len(pts)])
    import ctypes as C

    p = _native.lisec_vfe_train_params()
    grid = torch.empty((1, 8, 200, 400, 128), dtype=torch.float32, device="cuda")
    st = fe._lib.lisec_vfe_train_forward(fe._h, C.byref(p), C.c_void_p(grid.data_ptr()), None)
    assert st == _native.LISEC_ERR_UNSUPPORTED and b"createModel" in fe._lib.lisec_last_error(fe._h)
    fe.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_model_png_graph_predicts_end_to_end(dtype, tmp_path):
    """Predict.predictMain's sequence (Predict.py:21-38) for a weight set of the OLDER graph, saved to and re-read from a
    Keras-layout .h5 (load_model tells the graph from the kernels' shapes), against the whole CPU oracle of that graph."""
    from lisec_b200 import compat
    from lisec_b200.h5write import write_keras_weights

    arch = MODEL_PNG
    mx, my, mz = 12, 20, 8
    rng = np.random.default_rng(3)
    sweeps = []
    for _ in range(2):
        n = 5000
        pts = np.stack([rng.uniform(-6.5, 6.5, n), rng.uniform(-5.5, 5.5, n), rng.uniform(-0.3, 2.3, n)], axis=1)
        pts[: n // 3, :2] *= 0.15  # a dense core: voxels past the T cap
        sweeps.append(pts.astype(np.float32))
    pack = synthetic_model_pack(5, arch)
    path = str(tmp_path / "older_graph.h5")
    write_keras_weights(path, pack)
    model = compat.load_model(path, nx=2 * mx, ny=2 * my, nz=mz, maxPoints=T)
    assert model.arch == arch
    dense = []
    for pts in sweeps:
        t = compat.VFE_preprocessing(pts, 0.5, 0.25, 0.25, T, mx, my, mz)
        dense.append(compat.sparse.to_dense(t, default_value=0., validate_indices=False))
    prob, reg = model.predict(compat.stack(dense, axis=0), dtype=dtype)
    assert prob.shape == (2, mx, my, 2) and reg.shape == (2, mx, my, 14) and prob.dtype == np.float32

    ref = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=T, maxVoxelX=mx, maxVoxelY=my, maxVoxelZ=mz)
    grids = []
    for pts in sweeps:
        vox = O.voxelize_np(pts, **ref)
        feat = O.vfe_forward(vox["features"].astype(np.float32), pack, post_dense=True)
        grids.append(O.scatter_dense(vox["coords"], feat, O.c_empty(pack, T), (mz, 2 * mx, 2 * my), dtype=np.float64))
    want_p, want_r = NO.network_forward(np.stack(grids), pack, arch=arch)
    for got, want in ((prob, want_p), (reg, want_r)):
        e = within(got, want)
        if dtype == "f32":  # north_star's float32 bar, element-wise
            assert e <= 1e-5, e
        else:  # the bf16 bars of tests/test_gpu_network.py: 2e-2 of the tensor's scale, 1e-2 in L2; element-wise on record
            err = np.abs(got.astype(np.float64) - want)
            emax, el2 = float(err.max() / np.abs(want).max()), float(np.sqrt((err ** 2).sum() / (want ** 2).sum()))
            print("bf16 older graph: max/scale %.3e rel-L2 %.3e element-wise %.3e" % (emax, el2, e))
            assert emax <= 2e-2 and el2 <= 1e-2 and e <= 4e-2, (emax, el2, e)

"""The reference's own call sequence (Predict.py:21-38, model_training.py:266-285) through the drop-in layer."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from lisec_b200 import constants as Constants  # noqa: E402
from lisec_b200 import synth  # noqa: E402
from lisec_b200.weights import load_npz, save_npz, synthetic_vfe_pack  # noqa: E402
from oracle import lisec_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
REF = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)


def rel_err(got, ref, feat):
    floor = np.sqrt(np.mean(feat * feat))
    return float((np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref), floor)).max())


def test_predict_main_call_sequence(tmp_path):
    from lisec_b200.compat import MaxPoolingVFELayer, RepeatLayer, VFE_preprocessing, load_model, sparse

    pack = synthetic_vfe_pack(5)
    path = os.path.join(tmp_path, "15SampleEpoch0.npz")
    save_npz(path, pack)
    model = load_model(path, custom_objects={"RepeatLayer": RepeatLayer, "MaxPoolingVFELayer": MaxPoolingVFELayer})
    assert all(np.array_equal(model.pack[k], pack[k]) for k in pack)

    sampleLidarPoints = synth.lyft_like_sweep(30_000, seed=4).astype(np.float64)  # combine_lidar_data -> float64 (n,3)
    # Predict.py:21-30, verbatim call shapes
    trainVFEPoints = VFE_preprocessing(sampleLidarPoints, Constants.voxelx, Constants.voxely, Constants.voxelz,
                                       Constants.maxPoints, Constants.nx // 2, Constants.ny // 2, Constants.nz)
    assert trainVFEPoints.dense_shape == [8, 200, 400, 35, 6]
    trainVFEPoints = sparse.reshape(trainVFEPoints, (1,) + trainVFEPoints.shape)
    testVFEPointsDense = sparse.to_dense(trainVFEPoints, default_value=0., validate_indices=False)
    assert testVFEPointsDense.shape == (1, 8, 200, 400, 35, 6)
    grid = model.predict_voxel_grid(testVFEPointsDense)  # the first 23 layers of model.predict (Predict.py:38)
    assert tuple(grid.shape) == (1, 8, 200, 400, 64)

    vox = O.voxelize_np(sampleLidarPoints, **REF)
    feat = O.vfe_forward(vox["features"].astype(np.float32), pack)
    want = O.scatter_dense(vox["coords"], feat, O.c_empty(pack, 35), (8, 200, 400), dtype=np.float64)
    assert rel_err(grid[0].cpu().numpy(), want, feat) <= 1e-5


def test_train_call_sequence_stacks_sweeps():
    from lisec_b200.compat import VFE_preprocessing, createModel, sparse, stack

    pack = synthetic_vfe_pack(6)
    points = []
    raw = [synth.lyft_like_sweep(n, seed=10 + i) for i, n in enumerate((9_000, 12_345, 7_001))]
    for sampleLidarPoints in raw:  # model_training.py:266-280
        vfe_points = VFE_preprocessing(sampleLidarPoints, Constants.voxelx, Constants.voxely, Constants.voxelz,
                                       Constants.maxPoints, Constants.nx // 2, Constants.ny // 2, Constants.nz)
        points.append(sparse.to_dense(vfe_points, default_value=0., validate_indices=False))
    trainPoints = stack(points, axis=0)  # model_training.py:285
    assert trainPoints.shape == (3, 8, 200, 400, 35, 6)
    model = createModel(Constants.nx, Constants.ny, Constants.nz, Constants.maxPoints, weights=pack)
    grid = model.predict_voxel_grid(trainPoints).cpu().numpy()
    for s, pts in enumerate(raw):
        vox = O.voxelize_np(pts, **REF)
        feat = O.vfe_forward(vox["features"].astype(np.float32), pack)
        want = O.scatter_dense(vox["coords"], feat, O.c_empty(pack, 35), (8, 200, 400), dtype=np.float64)
        assert rel_err(grid[s], want, feat) <= 1e-5


def test_sparse_tensor_fields_match_golden():
    """.indices / .values of the returned object == what the reference's VFE_preprocessing returned (golden vectors
    from its own source lines), after the Keras float32 input cast."""
    from lisec_b200.compat import VFE_preprocessing

    with np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_first.npz")) as z:
        g = {k: z[k] for k in z.files}
    t = VFE_preprocessing(g["points"], 0.5, 0.25, 0.25, 35, 100, 200, 8)
    assert t.dense_shape == g["dense_shape"].tolist()
    assert np.array_equal(t.indices, g["indices"])
    assert t.values.tobytes() == g["values"].astype(np.float32).tobytes()


def test_dense_input_materialises_on_small_grid():
    from lisec_b200.compat import VFE_preprocessing, sparse

    rng = np.random.default_rng(0)
    pts = rng.normal([0, 0, 0.5], [1.0, 0.8, 0.3], size=(3000, 3))
    t = VFE_preprocessing(pts, 0.5, 0.25, 0.25, 35, 5, 8, 4)
    dense = sparse.to_dense(sparse.reshape(t, (1,) + t.shape)).numpy()
    vox = O.voxelize_np(pts, xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=5, maxVoxelY=8, maxVoxelZ=4)
    ind, val = O.coo_from_voxels(vox, 35)
    assert dense[0].tobytes() == O.to_dense(ind, val, [4, 10, 16, 35, 6]).astype(np.float32).tobytes()

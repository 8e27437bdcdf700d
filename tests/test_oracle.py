"""The oracle against the golden vectors minted from the reference's own source (tests/golden/make_golden.py), and
against itself (loops == vectorised, dense == sparse). CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from lisec_b200 import synth
from lisec_b200.weights import synthetic_vfe_pack
from oracle import lisec_oracle as O
from oracle import literal_reference as lit


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(golden_dir, name):
    with np.load(os.path.join(golden_dir, name)) as z:
        return {k: z[k] for k in z.files}


def groups_of(g):
    return [g["groups_flat"][g["groups_off"][i]:g["groups_off"][i + 1]].tolist() for i in range(len(g["groups_off"]) - 1)]


def test_loops_oracle_reproduces_seeded_reference_run(golden_dir, ref_args):
    """np.random.seed(3) + the reference's np.random.choice call sequence: bit-identical COO (model_training.py:132)."""
    g = load(golden_dir, "tiny_rng.npz")
    np.random.seed(int(g["seed"]))
    st, clustered, _, _ = O.vfe_preprocessing_loops(g["points"], sampler="numpy_rng", return_groups=True, **ref_args)
    assert np.array_equal(np.asarray(st.indices), g["indices"])
    assert np.asarray(st.values, dtype=np.float64).tobytes() == g["values"].tobytes()
    assert list(st.dense_shape) == g["dense_shape"].tolist() == [8, 200, 400, 35, 6]
    assert [v for v in clustered.values()] == groups_of(g)


@pytest.mark.parametrize("name", ["tiny_first.npz", "adversarial.npz"])
def test_loops_and_vectorised_oracle_match_first_T_golden(golden_dir, ref_args, name):
    g = load(golden_dir, name)
    st, clustered, _, _ = O.vfe_preprocessing_loops(g["points"], sampler="first_T", return_groups=True, **ref_args)
    assert np.array_equal(np.asarray(st.indices), g["indices"])
    assert np.asarray(st.values, dtype=np.float64).tobytes() == g["values"].tobytes()  # incl. the sign of zeros
    assert [v for v in clustered.values()] == groups_of(g)

    vox = O.voxelize_np(g["points"], **ref_args)
    ind, val = O.coo_from_voxels(vox, ref_args["sampleSize"])
    assert np.array_equal(ind, g["indices"])
    assert val.tobytes() == g["values"].tobytes()
    # counts and ordered lists, in dict (first appearance) order
    order = np.argsort(vox["first_idx"], kind="stable")
    gr = groups_of(g)
    assert vox["counts"][order].tolist() == [len(x) for x in gr]
    for row, lst in zip(vox["point_idx"][order], gr):
        kept = lst[:35]
        assert row[:len(kept)].tolist() == kept and (row[len(kept):] == -1).all()


def test_golden_covers_the_edge_cases(golden_dir):
    g = load(golden_dir, "adversarial.npz")
    counts = np.diff(g["groups_off"])
    assert (counts > 35).any() and (counts == 35).any() and (counts == 1).any()
    ind = g["indices"]
    # plane 0 of every axis is never occupied (strict range test); the top planes are reachable
    assert ind[:, 0].min() >= 1 and ind[:, 1].min() >= 1 and ind[:, 2].min() >= 1
    assert ind[:, 0].max() == 7 and ind[:, 1].max() == 199 and ind[:, 2].max() == 399
    assert np.signbit(g["values"][g["values"] == 0]).any()  # negative zeros survive


def test_vectorised_oracle_matches_100k_sweep_digests(golden_dir, ref_args):
    with open(os.path.join(golden_dir, "sweep100k.json")) as f:
        meta = json.load(f)
    pts = synth.lyft_like_sweep(100_000, seed=0)
    assert sha(pts) == meta["points_sha256"], "the synthetic sweep generator drifted"
    vox = O.voxelize_np(pts, **ref_args)
    assert len(vox["counts"]) == meta["n_voxels"]
    assert int(vox["counts"].sum()) == meta["n_in_range"]
    ind, val = O.coo_from_voxels(vox, ref_args["sampleSize"])
    assert len(val) == meta["nnz"]
    assert sha(ind.astype(np.int32)) == meta["indices_int32_sha256"]
    assert sha(val) == meta["values_float64_sha256"]
    assert sha(val.astype(np.float32)) == meta["values_float32_sha256"]


@pytest.mark.skipif(not lit.available(), reason="/root/reference is only present in the build container")
def test_oracle_matches_reference_source_live(ref_args):
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform([-6, -3, 0], [6, 3, 2.2], size=(500, 3)),
                          np.tile([[1.26, 0.3, 0.9]], (41, 1))]).astype(np.float32)
    st, groups = lit.run(pts, sampler="first_T", **ref_args)
    vox = O.voxelize_np(pts, **ref_args)
    ind, val = O.coo_from_voxels(vox, 35)
    assert np.array_equal(ind, np.asarray(st.indices))
    assert val.tobytes() == np.asarray(st.values, dtype=np.float64).tobytes()
    st2, groups2 = lit.run(pts, sampler="numpy_rng", seed=9, **ref_args)
    np.random.seed(9)
    st3 = O.vfe_preprocessing_loops(pts, sampler="numpy_rng", **ref_args)
    assert np.asarray(st3.values).tobytes() == np.asarray(st2.values, dtype=np.float64).tobytes()
    assert groups == groups2


def test_empty_and_all_dropped_clouds(ref_args):
    for pts in (np.zeros((0, 3)), np.asarray([[1e3, 0, 1.0], [0, 0, 0.1], [np.nan, 0, 1], [0, np.inf, 1]])):
        vox = O.voxelize_np(pts, **ref_args)
        assert len(vox["counts"]) == 0 and vox["features"].shape == (0, 35, 6)
    assert vox["n_nonfinite"] == 2 and vox["n_out_of_range"] == 2


def test_dense_forward_equals_sparse_forward_on_toy_grid():
    """The unmasked network on the dense input == per-voxel forward scattered over a c_empty background (§2.3-7)."""
    args = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=5, maxVoxelX=3, maxVoxelY=4, maxVoxelZ=4)
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform([-1.4, -0.9, 0.0], [1.4, 0.9, 1.0], size=(60, 3)),
                          rng.uniform([0.5, 0.25, 0.5], [1.0, 0.5, 0.75], size=(9, 3))])  # one voxel over the cap
    pack = synthetic_vfe_pack(1)
    vox = O.voxelize_np(pts, **args)
    assert (vox["counts"] > 5).any() and (vox["counts"] == 1).any()
    ind, val = O.coo_from_voxels(vox, 5)
    dense = O.to_dense(ind, val, [4, 6, 8, 5, 6])
    full = O.vfe_forward(dense, pack)
    sparse = O.scatter_dense(vox["coords"], O.vfe_forward(vox["features"], pack), O.c_empty(pack, 5), (4, 6, 8),
                             dtype=np.float64)
    assert np.abs(full - sparse).max() < 1e-12
    ce = O.c_empty(pack, 5)
    assert np.abs(ce).max() > 1e-3, "synthetic BN statistics must make the empty-voxel vector non-zero"
    # zero-filling instead of c_empty is NOT parity
    assert np.abs(full - O.scatter_dense(vox["coords"], O.vfe_forward(vox["features"], pack), 0 * ce, (4, 6, 8),
                                         dtype=np.float64)).max() > 1e-3


def test_virtual_pad_row_equivalence():
    """All pad rows of a voxel are identical, so one zero row reproduces T - s of them; a full voxel has none."""
    pack = synthetic_vfe_pack(2)
    rng = np.random.default_rng(1)
    for s in (1, 3, 34, 35):
        f = np.zeros((1, 35, 6))
        f[0, :s] = rng.normal(size=(s, 6))
        ref = O.vfe_forward(f, pack)
        rows = f[:, :min(s + 1, 35)]  # kept rows + one pad row (none when s == T)
        assert np.abs(O.vfe_forward(rows, pack) - ref).max() == 0.0


def test_float32_forward_noise_floor():
    pack = synthetic_vfe_pack(0)
    vox = O.voxelize_np(synth.lyft_like_sweep(5000, seed=3), xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35,
                        maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)
    f64 = O.vfe_forward(vox["features"], pack, np.float64)
    f32 = O.vfe_forward(vox["features"].astype(np.float32), pack, np.float32)
    scale = np.abs(f64).max()
    assert np.abs(f32 - f64).max() <= 1e-5 * scale

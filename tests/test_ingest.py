"""Lidar ingest: rotate_points / combine_lidar_data (model_training.py:65-98).

CPU: the oracle against the golden vectors minted from the reference's own lines (tests/golden/make_golden_ingest.py),
the exact FMA-chain restatement against np.dot, the product's quaternion -> matrix host step against the oracle's.
GPU: ingest_kernel through the C ABI, bit for bit against both."""
import os

import numpy as np
import pytest

from oracle import ingest_oracle as IO


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "ingest.npz"))


def test_oracle_matches_the_reference_lines(golden):
    for p in "ab":
        got = IO.transform_segments(golden[p + "_records"], golden[p + "_offsets"], golden[p + "_quats"], golden[p + "_trans"])
        assert got.dtype == np.float64 and got.tobytes() == golden[p + "_points"].tobytes()
    n0 = int(golden["a_offsets"][1])
    inv = IO.rotate_points(golden["a_records"][:n0, :3], golden["a_quats"][0], inverse=True)
    assert inv.tobytes() == golden["a_inverse_first"].tobytes()


@pytest.mark.skipif(not IO.literal_available(), reason="needs /root/reference (build container)")
def test_literal_reference_lines_reproduce_the_golden_file(golden, tmp_path):
    from lisec_b200 import synth

    _, combine = IO.literal_functions()
    sample, tables = synth.synthetic_lyft_sample(str(tmp_path), n_points=2500, seed=5)
    assert combine(sample, str(tmp_path), tables).tobytes() == golden["a_points"].tobytes()
    assert IO.combine_lidar_data(sample, str(tmp_path), tables).tobytes() == golden["a_points"].tobytes()


def test_fma_chain_is_what_numpy_dot_computes(golden):
    """The kernel's rounding sequence, restated exactly, equals np.dot on this host (and therefore the golden file)."""
    rec, off = golden["a_records"], golden["a_offsets"]
    for s in range(3):
        a, b = int(off[s]), min(int(off[s + 1]), int(off[s]) + 400)
        m = IO.Quaternion(golden["a_quats"][s]).rotation_matrix
        chain = IO.rotate_points_fma_chain(rec[a:b, :3], m, golden["a_trans"][s])
        assert chain.tobytes() == golden["a_points"][a:b].tobytes()


def test_quaternion_matrix_host_step():
    from lisec_b200.ingest import quaternion_rotation_matrix

    rng = np.random.default_rng(2)
    for k in range(50):
        q = rng.normal(size=4) * (1.0 if k % 2 else 3.0)
        for inverse in (False, True):
            want = (IO.Quaternion(q).inverse if inverse else IO.Quaternion(q)).rotation_matrix
            got = quaternion_rotation_matrix(q, inverse)
            assert got.tobytes() == np.ascontiguousarray(want).tobytes()
        # textbook closed form of the unit quaternion's matrix
        w, x, y, z = q / np.linalg.norm(q)
        ref = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(quaternion_rotation_matrix(q) - ref).max() < 1e-15
    m = quaternion_rotation_matrix([1.0, 0.0, 0.0, 0.0])
    assert np.array_equal(m, np.eye(3))


# ---- GPU -----------------------------------------------------------------------------------------------------------
def check_points(got, rec, off, quats, trans, exact_rows=250):
    """The kernel's output against the oracle. Bit for bit against the exact FMA-chain restatement on the first rows of
    every segment (the golden file pins that this chain IS the reference's np.dot on the build container); against the
    oracle's np.dot everywhere within 4 eps (|x|+|y|+|z|+|t|): np.dot's rounding sequence belongs to the host's BLAS
    kernel and is not the same on every CPU (seen on the GPU box: last-bit differences at n = 100 k)."""
    got = np.asarray(got)
    with np.errstate(invalid="ignore"):
        want = IO.transform_segments(rec, off, quats, trans)
    fin = np.isfinite(want).all(axis=1)
    bound = 4 * np.finfo(np.float64).eps * (np.abs(rec[:, :3].astype(np.float64)).sum(axis=1) + 4.0)
    assert got.shape == want.shape and got.dtype == np.float64
    assert (np.abs(got[fin] - want[fin]).max(axis=1) <= bound[fin]).all()
    for s in range(len(off) - 1):
        a, b = int(off[s]), min(int(off[s + 1]), int(off[s]) + exact_rows)
        rows = np.arange(a, b)[fin[a:b]]
        if len(rows):
            m = IO.Quaternion(quats[s]).rotation_matrix
            chain = IO.rotate_points_fma_chain(rec[rows, :3], m, trans[s])
            assert got[rows].tobytes() == chain.tobytes(), s
    return want, fin


@pytest.mark.gpu
def test_gpu_ingest_matches_golden_bit_for_bit(golden):
    from lisec_b200.ingest import LidarIngest, rotate_points

    ing = LidarIngest()
    for p in "ab":
        out = ing.transform(golden[p + "_records"], golden[p + "_offsets"], golden[p + "_quats"], golden[p + "_trans"])
        assert out.dtype.is_floating_point and out.shape == golden[p + "_points"].shape
        assert out.cpu().numpy().tobytes() == golden[p + "_points"].tobytes()
        assert ing.last_launch_count == 1
    n0 = int(golden["a_offsets"][1])
    inv = rotate_points(golden["a_records"][:n0, :3], golden["a_quats"][0], inverse=True)
    assert inv.tobytes() == golden["a_inverse_first"].tobytes()


@pytest.mark.gpu
def test_gpu_combine_lidar_data_drop_in(tmp_path):
    from lisec_b200 import synth
    from lisec_b200.ingest import combine_lidar_data, combine_lidar_data_device

    for seed, sensors, n in ((1, ("LIDAR_TOP", "LIDAR_FRONT_RIGHT", "LIDAR_FRONT_LEFT"), 100_003),
                             (2, ("LIDAR_TOP",), 257), (3, ("LIDAR_FRONT_LEFT", "LIDAR_TOP"), 1)):
        d = str(tmp_path / ("s%d" % seed))
        sample, tables = synth.synthetic_lyft_sample(d, n_points=n, seed=seed, sensors=sensors)
        got = combine_lidar_data(sample, d, tables)
        assert got.dtype == np.float64 and got.shape == (n, 3)
        files = [tables.get("sample_data", sample["data"][s]) for s in IO.SENSOR_TYPES if s in sample["data"]]
        recs = [np.fromfile(os.path.join(d, f["filename"]), dtype=np.float32).reshape(-1, 5) for f in files]
        cs = [tables.get("calibrated_sensor", f["calibrated_sensor_token"]) for f in files]
        check_points(got, np.concatenate(recs), np.cumsum([0] + [len(r) for r in recs]),
                     [c["rotation"] for c in cs], [c["translation"] for c in cs])
        want = IO.combine_lidar_data(sample, d, tables)
        assert want.shape == got.shape and np.abs(want - got).max() < 1e-12
        assert combine_lidar_data_device(sample, d, tables).is_cuda


@pytest.mark.gpu
def test_gpu_ingest_many_segments_ragged_and_special_values():
    """More segments than one launch carries (24), empty segments, signed zeros, non-finite records passed through."""
    import torch

    from lisec_b200.ingest import LidarIngest

    rng = np.random.default_rng(7)
    sizes = [0, 1, 255, 256, 257, 0, 1000] + [int(v) for v in rng.integers(0, 600, size=55)]
    off = np.cumsum([0] + sizes)
    rec = (rng.normal(size=(off[-1], 5)) * 40).astype(np.float32)
    rec[3] = [0.0, -0.0, 0.0, 1, 2]
    rec[4] = [np.inf, 1.0, -2.0, 0, 0]
    rec[5] = [np.nan, 1.0, -2.0, 0, 0]
    quats = rng.normal(size=(len(sizes), 4))
    trans = rng.normal(size=(len(sizes), 3)) * 2
    ing = LidarIngest()
    out = ing.transform(rec, off, quats, trans).cpu().numpy()
    assert ing.last_launch_count == 3  # 62 segments, 24 per launch
    want, fin = check_points(out, rec, off, quats, trans, exact_rows=40)
    assert np.array_equal(np.isnan(out), np.isnan(want)) and np.array_equal(np.isinf(out[~fin]), np.isinf(want[~fin]))
    # no segments / no points
    assert ing.transform(np.zeros((0, 5), np.float32), [0], [], []).shape == (0, 3)
    assert ing.transform(np.zeros((0, 5), np.float32), [0, 0], [quats[0]], [trans[0]]).shape == (0, 3)
    # records already on the device, caller-owned output
    buf = torch.empty((off[-1] + 10, 3), dtype=torch.float64, device="cuda")
    out2 = ing.transform(torch.from_numpy(rec).cuda(), off, quats, trans, out=buf)
    assert out2.data_ptr() == buf.data_ptr() and out2.cpu().numpy()[fin].tobytes() == out[fin].tobytes()


@pytest.mark.gpu
def test_gpu_ingest_feeds_the_voxelizer(ref_args):
    """records -> ingest_kernel -> float64 points on the device -> lisec_voxelize: same grouping as the CPU chain."""
    from lisec_b200 import Frontend
    from lisec_b200.ingest import LidarIngest
    from oracle import lisec_oracle as O

    rng = np.random.default_rng(3)
    n = 30_000
    r, az = rng.gamma(2.0, 9.0, size=n), rng.uniform(0, 2 * np.pi, size=n)
    rec = np.stack([r * np.cos(az), r * np.sin(az), rng.normal(-1.0, 0.7, size=n), np.zeros(n), np.zeros(n)], 1).astype(np.float32)
    off, quats = [0, 12_000, 21_000, n], [[0.9999, 0.002, -0.004, 0.008], [0.92, 0.01, 0.0, -0.38], [0.92, -0.01, 0.0, 0.38]]
    trans = [[1.2, 0.0, 1.83], [2.04, -0.61, 0.63], [2.04, 0.61, 0.62]]
    pts_dev = LidarIngest().transform(rec, off, quats, trans)
    want_pts = pts_dev.cpu().numpy()  # checked against the oracle here, then the oracle voxelizes the same float64 points
    check_points(want_pts, rec, off, quats, trans)
    fe = Frontend(max_points=n, max_sweeps=1)
    fe.voxelize(pts_dev, [0, n])
    vs = fe.export()
    vox = O.voxelize_np(want_pts, **ref_args)
    assert np.array_equal(vs.coords.cpu().numpy()[:, 1:], vox["coords"])
    assert np.array_equal(vs.point_idx.cpu().numpy(), vox["point_idx"])
    assert vs.features.cpu().numpy().tobytes() == vox["features"].astype(np.float32).tobytes()
    fe.close()

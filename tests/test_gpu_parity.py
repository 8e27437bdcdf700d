"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the committed golden vectors.

Bars: bit-exact for voxel coordinates, point->voxel assignment, counts, ordered lists and the float32 feature rows;
VFE / grid outputs within REL_TOL = 1e-5 of the float64 oracle, measured as |gpu - ref| <= REL_TOL * max(|ref|, rms(ref))
(elementwise relative error with the tensor's RMS magnitude as the floor for near-zero ReLU outputs)."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from lisec_b200 import synth  # noqa: E402
from lisec_b200._native import LisecError  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402
from oracle import lisec_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
REF = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)


def within(gpu, ref, tol=REL_TOL, floor=None):
    ref = np.asarray(ref, dtype=np.float64)
    if floor is None:
        floor = np.sqrt(np.mean(ref * ref)) if ref.size else 1.0
    err = np.abs(np.asarray(gpu, dtype=np.float64) - ref) / np.maximum(np.abs(ref), floor)
    return float(err.max()) if err.size else 0.0


@pytest.fixture(scope="module")
def fe():
    from lisec_b200 import Frontend

    f = Frontend(device=0, max_points=1_100_000, max_sweeps=8)
    f.set_weights(synthetic_vfe_pack(0))
    yield f
    f.close()


def oracle_batch(points, offsets):
    """Per-sweep oracle results concatenated in the product's voxel order (sweep, then linear cell id)."""
    vs = [O.voxelize_np(points[offsets[s]:offsets[s + 1]], **REF) for s in range(len(offsets) - 1)]
    coords = np.concatenate([np.concatenate([np.full((len(v["counts"]), 1), s), v["coords"]], axis=1)
                             for s, v in enumerate(vs)]) if vs else np.zeros((0, 4), np.int64)
    return vs, {
        "coords": coords,
        "counts": np.concatenate([v["counts"] for v in vs]),
        "point_idx": np.concatenate([v["point_idx"] for v in vs]),
        "features": np.concatenate([v["features"] for v in vs]),
        "per_sweep": np.asarray([len(v["counts"]) for v in vs]),
        "n_nonfinite": sum(v["n_nonfinite"] for v in vs),
        "n_out_of_range": sum(v["n_out_of_range"] for v in vs),
    }


def check_grouping(fe, points, offsets):
    fe.voxelize(points, offsets)
    vs = fe.export()
    _, ref = oracle_batch(np.asarray(points), offsets)
    assert vs.n_voxels_per_sweep.tolist() == ref["per_sweep"].tolist()
    assert vs.n_dropped_nonfinite == ref["n_nonfinite"]
    assert vs.n_dropped_out_of_range == ref["n_out_of_range"]
    assert vs.n_points_in_range == int(ref["counts"].sum())
    assert np.array_equal(vs.coords.cpu().numpy(), ref["coords"])          # bit-exact voxel coordinates
    assert np.array_equal(vs.counts.cpu().numpy(), ref["counts"])          # bit-exact counts
    assert np.array_equal(vs.point_idx.cpu().numpy(), ref["point_idx"])    # bit-exact ordered first-T lists
    got = vs.features.cpu().numpy()
    want = ref["features"].astype(np.float32)                              # the Keras input cast
    assert got.tobytes() == want.tobytes()                                 # bit-exact rows, signed zeros included
    return vs, ref


def test_grouping_bit_exact_100k_sweep(fe):
    pts = synth.lyft_like_sweep(100_000, seed=0)
    vs, ref = check_grouping(fe, pts, [0, len(pts)])
    with open(os.path.join(os.path.dirname(__file__), "golden", "sweep100k.json")) as f:
        meta = json.load(f)
    assert vs.coords.shape[0] == meta["n_voxels"] and vs.n_points_in_range == meta["n_in_range"]


def test_grouping_bit_exact_adversarial_with_nonfinite(fe):
    adv = synth.adversarial_tail()
    bad = np.asarray([[np.nan, 0, 1], [0, np.inf, 1], [1, 1, -np.inf], [np.nan, np.nan, np.nan]], dtype=np.float32)
    pts = np.concatenate([adv[:300], bad, adv[300:]])
    vs, ref = check_grouping(fe, pts, [0, len(pts)])
    assert vs.n_dropped_nonfinite == 4
    c = vs.counts.cpu().numpy()
    assert (c > 35).any() and (c == 35).any()


@pytest.mark.parametrize("name", ["tiny_first.npz", "adversarial.npz"])
def test_against_golden_from_reference_source(fe, name):
    """GPU output rebuilt into the reference's COO (dict order, (z,x,y,i,j), 210 entries per voxel) == the golden
    vectors produced by the reference's own lines model_training.py:103-152."""
    with np.load(os.path.join(os.path.dirname(__file__), "golden", name)) as z:
        g = {k: z[k] for k in z.files}
    fe.voxelize(g["points"], [0, len(g["points"])])
    vs = fe.export()
    first = vs.point_idx[:, 0].cpu().numpy()
    order = np.argsort(first, kind="stable")  # dict order = first appearance (model_training.py:123-126)
    coords = vs.coords.cpu().numpy()[order]
    feats = vs.features.cpu().numpy()[order]
    V = len(order)
    ii, jj = np.meshgrid(np.arange(35), np.arange(6), indexing="ij")
    ind = np.empty((V, 35, 6, 5), dtype=np.int64)
    for k in range(3):
        ind[..., k] = coords[:, k + 1, None, None]
    ind[..., 3], ind[..., 4] = ii[None], jj[None]
    assert np.array_equal(ind.reshape(-1, 5), g["indices"])
    assert feats.reshape(-1).tobytes() == g["values"].astype(np.float32).tobytes()
    counts = np.diff(g["groups_off"])
    assert np.array_equal(vs.counts.cpu().numpy()[order], counts)


def test_grouping_multi_sweep_ragged_and_empty_sweeps(fe):
    a = synth.lyft_like_sweep(20_001, seed=1)
    b = synth.lyft_like_sweep(3, seed=2, tail=False)
    c = synth.lyft_like_sweep(7_777, seed=3)
    pts = np.concatenate([a, b, c])
    offsets = [0, len(a), len(a), len(a) + len(b), len(pts), len(pts)]  # sweeps 1 and 4 are empty
    check_grouping(fe, pts, offsets)


def test_grouping_float64_points(fe):
    rng = np.random.default_rng(4)
    pts = rng.uniform([-52, -52, -0.3], [52, 52, 2.3], size=(50_000, 3))  # genuine float64 mantissas
    pts[:1000] = np.round(pts[:1000] * 4) / 4  # and exact voxel faces
    check_grouping(fe, pts, [0, len(pts)])


def test_grouping_saturated_cloud(fe):
    """BASELINE config 4 shape (scaled to run in seconds on the oracle side): near-field voxels far above T."""
    pts = synth.saturated_cloud(300_000, n_sweeps=3, theta=2.0)
    vs, ref = check_grouping(fe, pts, [0, len(pts)])
    assert (vs.counts.cpu().numpy() > 35).sum() > 500


def test_grouping_is_deterministic(fe):
    pts = synth.saturated_cloud(200_000, n_sweeps=2, theta=2.5)
    outs = []
    for _ in range(3):
        fe.voxelize(pts, [0, len(pts)])
        vs = fe.export()
        outs.append((vs.coords.cpu().numpy().tobytes(), vs.point_idx.cpu().numpy().tobytes(),
                     vs.features.cpu().numpy().tobytes()))
    assert outs[0] == outs[1] == outs[2]


def test_empty_cloud_and_all_dropped(fe):
    for pts in (np.zeros((0, 3), np.float32), np.asarray([[1e3, 0, 1.0], [0, 0, 0.1]], np.float32)):
        fe.voxelize(pts, [0, len(pts)])
        vs = fe.export()
        assert vs.coords.shape[0] == 0
        grid = fe.forward(pts, [0, len(pts)])
        ce = fe.c_empty()
        assert torch.equal(grid.reshape(-1, 64).cpu(), torch.from_numpy(ce).expand(640_000, 64))


def test_c_empty_matches_oracle(fe):
    ce = fe.c_empty()
    ref = O.c_empty(synthetic_vfe_pack(0), 35)
    assert np.abs(ce).max() > 1e-3
    assert within(ce, ref) <= REL_TOL


@pytest.mark.parametrize("seed", [0, 1])
def test_vfe_forward_within_tolerance(fe, seed):
    pack = synthetic_vfe_pack(seed)
    fe.set_weights(pack)
    pts = synth.lyft_like_sweep(100_000, seed=seed)
    fe.voxelize(pts, [0, len(pts)])
    got = fe.vfe().cpu().numpy()
    vox = O.voxelize_np(pts, **REF)
    ref = O.vfe_forward(vox["features"].astype(np.float32), pack, np.float64)  # float32 input cast, float64 math
    assert got.shape == ref.shape
    assert within(got, ref) <= REL_TOL
    fe.set_weights(synthetic_vfe_pack(0))


def test_vfe_forward_saturated_and_float64(fe):
    pack = synthetic_vfe_pack(0)
    pts = synth.saturated_cloud(150_000, n_sweeps=3, theta=2.0).astype(np.float64)
    fe.voxelize(pts, [0, len(pts)])
    got = fe.vfe().cpu().numpy()
    vox = O.voxelize_np(pts, **REF)
    ref = O.vfe_forward(vox["features"].astype(np.float32), pack, np.float64)
    assert within(got, ref) <= REL_TOL


def test_scatter_writes_every_cell_once(fe):
    pts, off = synth.sweep_batch(2, 60_000, seed0=5)
    fe.voxelize(pts, off)
    vs = fe.export(features=False)
    feat = fe.vfe()
    grid = fe.scatter(feat)
    assert grid.shape == (2, 8, 200, 400, 64) and grid.dtype == torch.float32
    ce = torch.from_numpy(fe.c_empty()).cuda()
    expect = ce.expand(2, 8, 200, 400, 64).clone()
    c = vs.coords.long()
    expect[c[:, 0], c[:, 1], c[:, 2], c[:, 3]] = feat
    assert torch.equal(grid, expect)  # bit-exact placement: occupied rows and the c_empty background


def test_fused_forward_equals_modular_and_host_entry(fe):
    pts, off = synth.sweep_batch(3, 40_000, seed0=9)
    fe.voxelize(pts, off)
    modular = fe.scatter(fe.vfe())
    poison = lambda: torch.full_like(modular, float("nan"))  # a cell nobody writes stays NaN and fails the compare
    fused = fe.forward(torch.from_numpy(pts).cuda(), off, out=poison())
    assert torch.equal(fused, modular)
    fe.voxelize(pts, off)
    assert torch.equal(fe.vfe_scatter_fused(out=poison()), modular)  # the fused stage alone, on an existing grouping
    pinned = torch.from_numpy(pts).pin_memory()
    host = fe.forward_host(pinned, off, out=poison())
    torch.cuda.synchronize()
    assert torch.equal(host, modular)
    assert fe.last_fused_kernel_ms > 0.0
    assert fe.last_launch_count == 6  # point pass, 2 scan kernels, fill (+ the tile plan in extra blocks), order (+ row tables), fused VFE + grid


def test_grid_against_dense_reference_forward_small_grid():
    """End to end on a grid small enough to run the reference's dense formulation on the CPU: emit the dense
    [N,nz,nx,ny,T,6] input, push it through the layer-by-layer oracle, compare with the sparse CUDA path."""
    from lisec_b200 import Frontend

    args = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=6, maxVoxelY=10, maxVoxelZ=4)
    f = Frontend(max_points=20_000, max_sweeps=2, max_voxel=(6, 10, 4))
    pack = synthetic_vfe_pack(3)
    f.set_weights(pack)
    rng = np.random.default_rng(2)
    pts = rng.normal([0, 0, 0.5], [1.5, 1.2, 0.4], size=(6000, 3)).astype(np.float32)
    off = [0, 2500, 6000]
    grid = f.forward(pts, off).cpu().numpy()
    dense = f.emit_dense_input().cpu().numpy()
    assert dense.shape == (2, 4, 12, 20, 35, 6)
    for s in range(2):
        vox = O.voxelize_np(pts[off[s]:off[s + 1]], **args)
        ind, val = O.coo_from_voxels(vox, 35)
        want = O.to_dense(ind, val, [4, 12, 20, 35, 6]).astype(np.float32)
        assert dense[s].tobytes() == want.tobytes()  # the reference's model input, bit for bit
        ref = O.vfe_forward(dense[s], pack, np.float64)  # the reference's dense formulation
        occ = ref[tuple(vox["coords"].T)]  # floor = RMS over the occupied voxels, as in the sparse tests
        assert within(grid[s], ref, floor=np.sqrt(np.mean(occ * occ))) <= REL_TOL
    f.close()


def test_bf16_grid():
    from lisec_b200 import Frontend

    f = Frontend(max_points=120_000, max_sweeps=1, grid_dtype="bf16")
    pack = synthetic_vfe_pack(0)
    f.set_weights(pack)
    pts = synth.lyft_like_sweep(100_000, seed=0)
    grid = f.forward(pts, [0, len(pts)])
    assert grid.dtype == torch.bfloat16
    feat = f.vfe()
    vs = f.export(features=False)
    c = vs.coords.long()
    assert torch.equal(grid[c[:, 0], c[:, 1], c[:, 2], c[:, 3]], feat.to(torch.bfloat16))  # one rounding
    vox = O.voxelize_np(pts, **REF)
    ref = O.scatter_dense(vox["coords"], O.vfe_forward(vox["features"].astype(np.float32), pack),
                          O.c_empty(pack, 35), (8, 200, 400), dtype=np.float64)
    assert within(grid[0].float().cpu().numpy(), ref, 2e-2) <= 2e-2
    f.close()


def test_full_size_properties_config2(fe):
    """BASELINE config 2 (8 sweeps x 100 k points): size-independent properties of the 1.31 GB grid."""
    pts, off = synth.sweep_batch(8, 100_000, seed0=0)
    grid = fe.forward(pts, off)
    per, V, nin, noor, nnf = fe.counts()
    assert nin + noor + nnf == len(pts)
    ce = torch.from_numpy(fe.c_empty()).cuda()
    flat = grid.view(-1, 64)
    occupied = (flat != ce).any(dim=1)
    vs = fe.export(features=False)
    lin = ((vs.coords[:, 0].long() * 8 + vs.coords[:, 1]) * 200 + vs.coords[:, 2]) * 400 + vs.coords[:, 3]
    assert torch.all(lin[1:] > lin[:-1])  # voxel rows ascend in (sweep, z, x, y)
    is_voxel = torch.zeros(flat.shape[0], dtype=torch.bool, device="cuda")
    is_voxel[lin] = True
    assert int(is_voxel.sum()) == V and not bool((occupied & ~is_voxel).any())  # nothing outside the voxel set differs
    assert torch.equal(flat[lin], fe.vfe())
    # sweep independence: sweep 3 alone gives the same slab
    alone = fe.forward(pts[off[3]:off[4]], [0, off[4] - off[3]])
    assert torch.equal(alone[0], grid[3])
    # plane 0 of every axis is pure background (strict range test)
    assert torch.equal(grid[:, 0], ce.expand_as(grid[:, 0]))
    assert torch.equal(grid[:, :, 0], ce.expand_as(grid[:, :, 0]))
    assert torch.equal(grid[:, :, :, 0], ce.expand_as(grid[:, :, :, 0]))


def test_error_behaviour(fe):
    with pytest.raises(LisecError) as e:
        fe.voxelize(np.zeros((1_100_001, 3), np.float32), [0, 1_100_001])
    assert e.value.status == -3
    with pytest.raises(LisecError) as e:
        fe.voxelize(np.zeros((10, 3), np.float32), [0] + [10] * 9)  # 9 sweeps > max_sweeps = 8
    assert e.value.status == -3
    with pytest.raises(LisecError) as e:
        fe.voxelize(np.zeros((10, 3), np.float32), [0, 7, 5, 10])
    assert e.value.status == -1
    from lisec_b200 import Frontend

    f = Frontend(max_points=1000, max_sweeps=1)
    with pytest.raises(LisecError) as e:
        f.forward(np.zeros((10, 3), np.float32), [0, 10])  # weights not set
    assert e.value.status == -5
    f.close()


def test_config4_million_point_cloud(fe):
    """BASELINE config 4: one aggregated ~1 M-point cloud, a large share of voxels at the T cap (no pad row, the
    first-35 rule decides which points count)."""
    pack = synthetic_vfe_pack(0)
    pts = synth.saturated_cloud(1_000_000, n_sweeps=10, theta=3.5)
    vs, ref = check_grouping(fe, pts, [0, len(pts)])
    assert int((ref["counts"] >= 35).sum()) > 1000  # the cap really is exercised
    got = fe.vfe().cpu().numpy()
    want = O.vfe_forward(ref["features"].astype(np.float32), pack, np.float64)
    assert within(got, want) <= REL_TOL


def test_full_voxels_and_tile_boundaries(fe):
    """Voxels with exactly T, T-1 and T+1 points back to back (with / without the virtual pad row), enough of them
    that many 256-row tiles end on a full voxel; plus single-point voxels so that tiles also end on 2-row voxels."""
    rng = np.random.default_rng(11)
    pts = []
    k = 0
    for ix in range(-60, 60):
        for iy in range(-40, 40, 2):
            n = (34, 35, 36, 1, 70)[k % 5]
            k += 1
            base = np.array([ix * 0.5 + 0.05, iy * 0.25 + 0.02, 0.55], np.float32)
            pts.append(base + rng.uniform(0, [0.4, 0.2, 0.15], size=(n, 3)).astype(np.float32))
    pts = np.concatenate(pts).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    pack = synthetic_vfe_pack(2)
    fe.set_weights(pack)
    vs, ref = check_grouping(fe, pts, [0, len(pts)])
    assert set(np.unique(ref["counts"])) >= {1, 34, 35, 36, 70}
    got = fe.vfe().cpu().numpy()
    want = O.vfe_forward(ref["features"].astype(np.float32), pack, np.float64)
    assert within(got, want) <= REL_TOL
    grid = fe.forward(pts, [0, len(pts)], out=torch.full((1, 8, 200, 400, 64), float("nan"), device="cuda"))
    c = vs.coords.long()
    assert torch.equal(grid[c[:, 0], c[:, 1], c[:, 2], c[:, 3]], torch.from_numpy(got).cuda())
    assert not bool(torch.isnan(grid).any())
    fe.set_weights(synthetic_vfe_pack(0))


@pytest.mark.parametrize("T", [2, 8, 64])
def test_other_sample_sizes_small_grid(T):
    """sampleSize is a parameter of VFE_preprocessing (model_training.py:112): the tile packing (rows per tile =
    256 - T + 1), the T cap and the pad-row rule must hold for any supported T."""
    from lisec_b200 import Frontend

    args = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=T, maxVoxelX=8, maxVoxelY=12, maxVoxelZ=4)
    f = Frontend(max_points=40_000, max_sweeps=2, max_voxel=(8, 12, 4), sample_size=T)
    pack = synthetic_vfe_pack(4)
    f.set_weights(pack)
    rng = np.random.default_rng(T)
    pts = rng.normal([0, 0, 0.5], [0.9, 0.7, 0.3], size=(30_000, 3)).astype(np.float32)  # dense core, sparse rim
    off = [0, 12_000, 30_000]
    grid = f.forward(pts, off, out=torch.full((2, 4, 16, 24, 64), float("nan"), device="cuda")).cpu().numpy()
    ce = O.c_empty(pack, T)
    for s in range(2):
        vox = O.voxelize_np(pts[off[s]:off[s + 1]], **args)
        assert vox["counts"].max() > T  # the cap is exercised
        feat = O.vfe_forward(vox["features"].astype(np.float32), pack, np.float64)
        want = O.scatter_dense(vox["coords"], feat, ce, (4, 16, 24), dtype=np.float64)
        assert within(grid[s], want, floor=np.sqrt(np.mean(feat * feat))) <= REL_TOL
    f.close()


@pytest.mark.parametrize("size", [(0.3, 0.2, 0.25), (0.7, 0.1, 0.3)])
def test_voxel_sizes_that_are_not_powers_of_two(size):
    """get_voxel divides by the voxel size (model_training.py:103-107). Constants.py's sizes are powers of two, where
    x * (1 / size) is the same float64 as x / size; for any other size the kernel must really divide. Points ON cell
    borders (k * size in float64, and its float64 neighbours) and on the range limits are where a reciprocal would differ:
    coordinates, counts, ordered lists and feature rows bit-exact against the oracle."""
    from lisec_b200 import Frontend

    xs, ys, zs = size
    mx, my, mz, T = 20, 30, 6, 35
    args = dict(xSize=xs, ySize=ys, zSize=zs, sampleSize=T, maxVoxelX=mx, maxVoxelY=my, maxVoxelZ=mz)
    rng = np.random.default_rng(17)
    n = 20_000
    pts = np.stack([rng.uniform(-mx * xs * 1.05, mx * xs * 1.05, n), rng.uniform(-my * ys * 1.05, my * ys * 1.05, n),
                    rng.uniform(-0.1 * zs, mz * zs * 1.05, n)], axis=1)
    kx, ky, kz = rng.integers(-mx - 1, mx + 2, 3000), rng.integers(-my - 1, my + 2, 3000), rng.integers(-1, mz + 2, 3000)
    border = np.stack([kx * xs, ky * ys, kz * zs], axis=1)  # exactly on cell borders, the range limits included
    up, down = np.nextafter(border, np.inf), np.nextafter(border, -np.inf)
    pts = np.concatenate([pts, border, up, down])
    rng.shuffle(pts)
    f = Frontend(max_points=len(pts), max_sweeps=1, voxel_size=size, max_voxel=(mx, my, mz), sample_size=T)
    f.voxelize(pts, [0, len(pts)])
    vs = f.export()
    ref = O.voxelize_np(pts, **args)
    assert ref["n_out_of_range"] > 0 and len(ref["counts"]) > 1000
    assert vs.n_dropped_out_of_range == ref["n_out_of_range"]
    assert np.array_equal(vs.coords.cpu().numpy()[:, 1:], ref["coords"])
    assert np.array_equal(vs.counts.cpu().numpy(), ref["counts"])
    assert np.array_equal(vs.point_idx.cpu().numpy(), ref["point_idx"])
    assert vs.features.cpu().numpy().tobytes() == ref["features"].astype(np.float32).tobytes()
    f.close()


def test_bf16_fused_grid_every_cell_written():
    from lisec_b200 import Frontend

    f = Frontend(max_points=250_000, max_sweeps=2, grid_dtype="bf16")
    f.set_weights(synthetic_vfe_pack(1))
    pts, off = synth.sweep_batch(2, 100_000, seed0=40)
    grid = f.forward(pts, off, out=torch.full((2, 8, 200, 400, 64), float("nan"), dtype=torch.bfloat16, device="cuda"))
    assert not bool(torch.isnan(grid).any())
    modular = f.scatter(f.vfe())
    assert torch.equal(grid, modular)
    f.close()

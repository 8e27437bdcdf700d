"""The VFE stack in training mode on the GPU (lisec_vfe_train_forward / _backward, lisec_b200/csrc/vfe_train.cu) against
the float64 oracle on rows with multiplicities (oracle/train_oracle.py: forward_train_rows — itself proven equal to the
dense graph of model_training.py:229-235 in tests/test_train_oracle.py). Parity unpinned like every floating-point half
of this path (no TensorFlow here); the bars below are float32-vs-float64 bars."""
import numpy as np
import pytest
import torch

from lisec_b200 import synth
from lisec_b200.weights import VFE_BN, VFE_DENSE, synthetic_vfe_pack
from oracle import lisec_oracle as O
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu
REF = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)
T, CELLS = 35, 8 * 200 * 400


def rel_l2(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.linalg.norm(np.asarray(got, np.float64) - ref) / max(np.linalg.norm(ref), 1e-30))


def clouds():
    rng = np.random.default_rng(11)
    a = synth.lyft_like_sweep(12_000, seed=5)
    dup = np.tile(np.asarray([[1.3, 0.3, 0.6]], np.float32), (4, 1))            # exact ties between kept rows
    full = rng.uniform([2.0, 1.0, 0.5], [2.49, 1.24, 0.74], size=(60, 3)).astype(np.float32)  # one voxel far above T
    b = synth.lyft_like_sweep(7_000, seed=6)
    return [np.concatenate([a, dup, full]), b]


@pytest.mark.parametrize("seed", [0, 3])
def test_vfe_training_forward_and_backward_match_the_row_oracle(seed):
    from lisec_b200 import Frontend
    from lisec_b200.train import VfeTrainer

    pack = synthetic_vfe_pack(seed)
    sweeps = clouds()
    pts = np.concatenate(sweeps)
    off = np.concatenate([[0], np.cumsum([len(s) for s in sweeps])]).tolist()
    fe = Frontend(device=0, max_points=len(pts), max_sweeps=2, grid_dtype="f32")
    dev = torch.device("cuda", 0)
    params = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(dev) for k, v in pack.items()}
    grads = {k: torch.zeros_like(v) for k, v in params.items() if "moving_" not in k}
    tr = VfeTrainer(fe, params, grads)
    grid = tr.forward(pts, off)
    g = torch.Generator(device="cpu").manual_seed(seed)
    # an upstream gradient with structure: dense noise plus a few large entries
    dgrid = torch.randn((2, 8, 200, 400, 64), generator=g, dtype=torch.float32) * 1e-3
    dgrid_dev = dgrid.to(dev)
    tr.backward(dgrid_dev)
    torch.cuda.synchronize()

    # ---- oracle on rows with multiplicities, float64 ----
    vox = [O.voxelize_np(s, **REF) for s in sweeps]
    kept = np.concatenate([np.minimum(v["counts"], T) for v in vox])
    rows = np.concatenate([v["features"][i, :k].astype(np.float32).astype(np.float64)  # the Keras float32 input cast
                           for v in vox for i, k in enumerate(np.minimum(v["counts"], T))])
    row_voxel = np.repeat(np.arange(len(kept)), kept)
    cells = np.concatenate([s * CELLS + v["linear"] for s, v in enumerate(vox)])
    V = len(kept)
    p = TO.to_params(pack)
    vout, eout, stats = TO.forward_train_rows(torch.from_numpy(rows), torch.from_numpy(row_voxel), torch.from_numpy(kept),
                                              2 * CELLS, T, p)
    gw = dgrid.double().reshape(-1, 64)
    g_occ = gw[torch.from_numpy(cells)]
    g_empty = gw.sum(0) - g_occ.sum(0)
    names = [k for k in pack if "moving_" not in k]
    ref_grads = torch.autograd.grad((vout * g_occ).sum() + (eout * g_empty).sum(), [p[k] for k in names])

    # forward: per-voxel rows, the empty voxels' row, the grid, the batch statistics and the moving statistics
    got_rows, mean3, var3 = tr.read_layer(2, V)
    assert rel_l2(got_rows[:V], vout.detach().numpy()) <= 2e-5
    assert rel_l2(got_rows[V], eout.detach().numpy()) <= 2e-5
    gridh = grid.cpu().numpy().reshape(-1, 64)
    assert np.array_equal(gridh[cells], got_rows[:V])
    mask = np.ones(len(gridh), bool)
    mask[cells] = False
    assert (gridh[mask] == got_rows[V]).all()
    for layer, bn in enumerate(VFE_BN):
        _, mean, var = tr.read_layer(layer, V)
        m_ref, v_ref = (t.numpy() for t in stats[bn])
        assert rel_l2(mean, m_ref) <= 2e-5 and rel_l2(var, v_ref) <= 1e-4, bn
        mm = params[bn + "/moving_mean"].cpu().numpy()
        assert rel_l2(mm, pack[bn + "/moving_mean"] * 0.99 + m_ref * 0.01) <= 1e-6
        mv = params[bn + "/moving_variance"].cpu().numpy()
        assert rel_l2(mv, pack[bn + "/moving_variance"] * 0.99 + v_ref * 0.01) <= 1e-6
    # backward: every parameter gradient
    for k, gr in zip(names, ref_grads):
        e = rel_l2(grads[k].cpu().numpy(), gr.numpy())
        assert e <= 2e-3, (k, e)
    fe.close()

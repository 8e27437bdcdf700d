"""The whole training step on the GPU (lisec_b200/train.py: TrainStep; compat.train = model_training.train, :260-302) on
a reduced grid: the loss of the first step against the float64 oracle of the whole graph, descent over a few steps, every
parameter moved, moving statistics updated, model.save / load_model round trip."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def small_clouds(n, seed):
    rng = np.random.default_rng(seed)
    return [rng.uniform([-5.9, -4.9, 0.26], [5.9, 4.9, 1.99], size=(1500 + 100 * i, 3)).astype(np.float32) for i in range(n)]


def test_train_step_loss_matches_the_oracle_and_descends():
    from lisec_b200.train import TrainStep
    from lisec_b200.weights import synthetic_model_pack
    from oracle import lisec_oracle as O
    from oracle import train_oracle as TO

    nx, ny, nz, B = 24, 40, 8, 2
    pack = {k: np.asarray(v, np.float32) for k, v in synthetic_model_pack(2).items()}
    clouds = small_clouds(B, 1)
    pts = np.concatenate(clouds)
    off = np.cumsum([0] + [len(c) for c in clouds]).tolist()
    g = torch.Generator(device="cpu").manual_seed(5)
    yc = torch.randint(0, 3, (B, nx // 2, ny // 2, 2), generator=g).float()
    yr = torch.randn((B, nx // 2, ny // 2, 14), generator=g) * 0.5
    LR = 0.002  # (the reference's 0.01 diverges on this synthetic pack — in the float64 oracle too: 1.9, 3.7, 19.8, 4.1)
    step = TrainStep(pack, batch=B, max_points=len(pts), nx=nx, ny=ny, nz=nz, lr=LR)
    losses = [float(step.step(pts, off, yc.cuda(), yr.cuda())) for _ in range(4)]
    after = step.to_pack()
    step.close()

    # the first step's loss against the float64 oracle of the WHOLE graph on the dense input (model_training.py:229-257)
    ref = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=nx // 2, maxVoxelY=ny // 2, maxVoxelZ=nz)
    dense = []
    for c in clouds:
        vox = O.voxelize_np(c, **ref)
        ind, val = O.coo_from_voxels(vox, 35)
        dense.append(O.to_dense(ind, val, [nz, nx, ny, 35, 6]))
    x = np.stack(dense).astype(np.float32).astype(np.float64)
    opack, accum, want = {k: np.asarray(v, np.float64) for k, v in pack.items()}, {}, []
    for it in range(3):  # three whole fit() steps in float64: forward, autograd, the Keras SGD-Nesterov update
        loss, opack, accum, _ = TO.train_step(opack, accum, it, x, yc.numpy().astype(np.float64), yr.numpy().astype(np.float64), lr=LR)
        want.append(loss)
    print("losses", losses, "oracle", want)
    # step 0 checks the forward; steps 1 and 2 check gradients + update: a wrong gradient would not track the oracle's descent
    assert abs(losses[0] - want[0]) <= 2e-2 * want[0]
    assert abs(losses[1] - want[1]) <= 5e-2 * want[1] and abs(losses[2] - want[2]) <= 8e-2 * want[2]
    assert losses[-1] < losses[0]                      # SGD on a fixed batch descends
    assert set(after) == set(pack)
    moved = [k for k in pack if np.abs(after[k] - pack[k]).max() > 0]
    # every weight moves except the biases in front of a training-mode BatchNormalization (zero gradient up to rounding)
    assert len(moved) >= len(pack) - 25, sorted(set(pack) - set(moved))
    for k in pack:
        assert after[k].shape == pack[k].shape and np.isfinite(after[k]).all(), k


def test_graph_replay_is_the_eager_step():
    """TrainStep(use_graph=True) captures the dense network's forward + loss + backward and the operand refresh after one
    eager step and replays them: the same kernels on the same buffers — the weights after five steps must be IDENTICAL to
    the eager run's (every reduction that feeds a gradient has a fixed order); the reported loss is a float64 sum built
    with atomics, equal to rounding."""
    from lisec_b200.train import TrainStep
    from lisec_b200.weights import synthetic_model_pack

    nx, ny, nz, B = 24, 40, 8, 2
    pack = {k: np.asarray(v, np.float32) for k, v in synthetic_model_pack(4).items()}
    batches = []
    for i in range(5):  # a different batch every step: replays must read the new inputs
        clouds = small_clouds(B, 10 + i)
        g = torch.Generator(device="cpu").manual_seed(i)
        batches.append((np.concatenate(clouds), np.cumsum([0] + [len(c) for c in clouds]).tolist(),
                        torch.randint(0, 3, (B, nx // 2, ny // 2, 2), generator=g).float().cuda(),
                        (torch.randn((B, nx // 2, ny // 2, 14), generator=g) * 0.5).cuda()))
    out = {}
    for mode in (False, True):
        step = TrainStep(pack, batch=B, max_points=max(len(b[0]) for b in batches), nx=nx, ny=ny, nz=nz, lr=0.002,
                         use_graph=mode)
        losses = [float(step.step(*b)) for b in batches]
        assert (step._graph_dense is not None) == mode
        out[mode] = (losses, step.to_pack())
        step.close()
    assert np.allclose(out[False][0], out[True][0], rtol=1e-12, atol=0), (out[False][0], out[True][0])
    for k in pack:
        assert np.array_equal(out[False][1][k], out[True][1][k]), k


def test_compat_train_runs_and_saves_a_loadable_model(tmp_path):
    from lisec_b200 import compat
    from lisec_b200 import constants as K

    n = 2
    rng = np.random.default_rng(0)
    clouds = {i: rng.uniform([-40, -40, 0.3], [40, 40, 1.9], size=(3000, 3)) for i in range(n)}
    labels = (rng.integers(0, 3, size=(n, K.nx // 2, K.ny // 2, 2)).astype(np.float32),
              rng.normal(size=(n, K.nx // 2, K.ny // 2, 14)).astype(np.float32))
    for ext in ("npz", "h5"):
        path = os.path.join(tmp_path, "model." + ext)
        hist = compat.train(list(range(n)), None, path, steps_per_epoch=3, labels=labels,
                            combine_lidar_data=lambda s, d, l: clouds[s])
        assert len(hist["loss"]) == 3 and all(np.isfinite(hist["loss"]))
        model = compat.load_model(path)
        assert len(model.pack) == 142
    hist2 = compat.train_with_model(list(range(n)), None, path, os.path.join(tmp_path, "m2.npz"), steps_per_epoch=2,
                                    labels=labels, combine_lidar_data=lambda s, d, l: clouds[s])
    assert len(hist2["loss"]) == 2

"""Measured parity error of the CUDA VFE against the float64 oracle over several weight packs and clouds (the metric
of tests/test_gpu_parity.py: |gpu - ref| / max(|ref|, rms(ref))). Prints one line per case and the worst."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402
from oracle import lisec_oracle as O  # noqa: E402

REF = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)


def within(gpu, ref):
    ref = np.asarray(ref, dtype=np.float64)
    floor = np.sqrt(np.mean(ref * ref))
    return float((np.abs(np.asarray(gpu, dtype=np.float64) - ref) / np.maximum(np.abs(ref), floor)).max())


fe = Frontend(device=0, max_points=400_000, max_sweeps=4)
clouds = {
    "lyft 100k": synth.lyft_like_sweep(100_000, seed=3),
    "lyft 20k": synth.lyft_like_sweep(20_000, seed=0),
    "saturated 150k": synth.saturated_cloud(150_000, n_sweeps=3, theta=2.0),
}
vox = {k: O.voxelize_np(p, **REF) for k, p in clouds.items()}
worst = 0.0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    pack = synthetic_vfe_pack(seed)
    fe.set_weights(pack)
    for name, pts in clouds.items():
        fe.voxelize(pts, [0, len(pts)])
        got = fe.vfe().cpu().numpy()
        ref = O.vfe_forward(vox[name]["features"].astype(np.float32), pack, np.float64)
        e = within(got, ref)
        worst = max(worst, e)
        print("seed %d  %-15s  %6d voxels  err %.2e" % (seed, name, len(got), e))
print("worst %.2e (bar 1e-5)" % worst)

"""Small, fixed command line for ncu: a few passes of the 8-sweep front end (configs[1]) on cuda:0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.weights import synthetic_vfe_pack

n_sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pts, off = synth.sweep_batch(n_sweeps, 100_000, seed0=0)
fe = Frontend(max_points=len(pts), max_sweeps=n_sweeps)
fe.set_weights(synthetic_vfe_pack(0))
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(n_sweeps)
for _ in range(iters):
    fe.forward(dev, off, out=grid)
torch.cuda.synchronize()
print("ok", fe.counts()[1], "voxels")

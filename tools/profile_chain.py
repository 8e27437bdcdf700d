"""ncu target: the grouping chain + row features on the 8-sweep batch (two warm passes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.weights import synthetic_vfe_pack

pts, off = synth.sweep_batch(8, 100_000, seed0=0)
fe = Frontend(max_points=len(pts), max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(8)
for _ in range(3):
    fe.forward(dev, off, out=grid)
torch.cuda.synchronize()
print("ok")

"""ncu target: one rows-only VFE launch (MODE 0) and one fused VFE + grid launch (MODE 1) on the 8-sweep batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.weights import synthetic_vfe_pack

pts, off = synth.sweep_batch(8, 100_000, seed0=0)
fe = Frontend(max_points=len(pts), max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))          # vfe_kernel launch #0 (c_empty)
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(8)
fe.voxelize(dev, off)
feat = fe.vfe()                                # launch #1: MODE 0
fe.forward(dev, off, out=grid)                 # launch #2: MODE 1
feat = fe.vfe(out=feat)                        # launch #3: MODE 0 (warm)
fe.forward(dev, off, out=grid)                 # launch #4: MODE 1 (warm)
torch.cuda.synchronize()
print("ok")

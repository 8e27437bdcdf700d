"""Device-time breakdown of the front end (CUDA events on the launching stream), 8 sweeps x 100k points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.weights import synthetic_vfe_pack

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

pts, off = synth.sweep_batch(8, 100_000, seed0=0)
fe = Frontend(max_points=len(pts), max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(8)
far = torch.full((len(pts), 3), 1000.0, device="cuda")  # every point out of range: the fused kernel only writes background
print("forward (real)        %.3f ms" % timeit(lambda: fe.forward(dev, off, out=grid)))
print("forward (all dropped) %.3f ms" % timeit(lambda: fe.forward(far, off, out=grid)))
print("voxelize (real)       %.3f ms" % timeit(lambda: fe.voxelize(dev, off)))
print("voxelize (dropped)    %.3f ms" % timeit(lambda: fe.voxelize(far, off)))
fe.voxelize(dev, off)
feat = fe.vfe()
print("centroids + vfe rows  %.3f ms" % timeit(lambda: fe.vfe(out=feat)))
print("grid_write            %.3f ms" % timeit(lambda: fe.scatter(feat, out=grid)))

"""Full inference with the first Conv3D reading the dense bf16 grid vs gathering from the sparse front-end output."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.network import DenseNetwork
from lisec_b200.weights import synthetic_network_pack, synthetic_vfe_pack


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pts, off = synth.sweep_batch(B, 100_000, seed0=0)
dev = torch.from_numpy(pts).cuda()
fe = Frontend(max_points=len(pts), max_sweeps=B, grid_dtype="bf16")
fe.set_weights(synthetic_vfe_pack(0))
pack = synthetic_network_pack(0)
dense = DenseNetwork(pack, batch=B)
sparse = DenseNetwork(pack, batch=B)
sparse.attach_frontend(fe)

def step_dense():
    fe.forward(dev, off, out=dense.grid)
    dense.forward()

print("%d sweeps: dense grid path   %.3f ms" % (B, timeit(step_dense)))
print("%d sweeps: sparse gather path %.3f ms" % (B, timeit(lambda: sparse.forward_sparse(dev, off))))
print("   first conv alone: dense %.3f ms, gather %.3f ms" % (timeit(lambda: dense.run_layers(0, 1)), timeit(lambda: sparse.run_layers(0, 1))))
print("   front end: fused bf16 grid %.3f ms, voxelize + VFE rows %.3f ms" % (
    timeit(lambda: fe.forward(dev, off, out=dense.grid)), timeit(lambda: (fe.voxelize(dev, off), fe.vfe(out=sparse.voxel_feat)))))

"""ncu target: the first Conv3D alone, reading the dense bf16 grid (conv_halo_kernel<false>) and gathering from the sparse
front-end output (conv_halo_kernel<true>), 8 sweeps; three launches each, dense first."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.network import DenseNetwork  # noqa: E402
from lisec_b200.weights import synthetic_network_pack, synthetic_vfe_pack  # noqa: E402

B = 8
pts, off = synth.sweep_batch(B, 100_000, seed0=0)
dev = torch.from_numpy(pts).cuda()
fe = Frontend(max_points=len(pts), max_sweeps=B, grid_dtype="bf16")
fe.set_weights(synthetic_vfe_pack(0))
pack = synthetic_network_pack(0)
dense = DenseNetwork(pack, batch=B)
sparse = DenseNetwork(pack, batch=B)
sparse.attach_frontend(fe)
fe.forward(dev, off, out=dense.grid)
fe.voxelize(dev, off)
fe.vfe(out=sparse.voxel_feat)
torch.cuda.synchronize()
for _ in range(3):
    dense.run_layers(0, 1)
torch.cuda.synchronize()
for _ in range(3):
    sparse.run_layers(0, 1)
torch.cuda.synchronize()
print("ok")

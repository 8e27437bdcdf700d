"""Front-end step and fused-kernel time (8 sweeps x 100 k points, bf16 grid), for whichever library LISEC_LIB_PATH names.
    python tools/fused_probe.py [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
batches = []
for b in range(4):
    sw = [synth.lyft_like_sweep(100_000, seed=8 * b + s) for s in range(8)]
    batches.append(torch.from_numpy(np.concatenate(sw)).cuda())
off = [100_000 * i for i in range(9)]
fe = Frontend(device=0, max_points=800_000, max_sweeps=8, grid_dtype="bf16")
fe.set_weights(synthetic_vfe_pack(0))
grid = fe.new_grid(8)
for i in range(10):
    fe.forward(batches[i % 4], off, out=grid)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
kms = []
a.record()
for i in range(steps):
    fe.forward(batches[i % 4], off, out=grid)
b.record()
torch.cuda.synchronize()
for i in range(20):
    fe.forward(batches[i % 4], off, out=grid)
    ms = C.c_float()
    fe._check(fe._lib.lisec_last_fused_kernel_ms(fe._h, C.byref(ms)))
    kms.append(ms.value)
print("%s: step %.4f ms, fused kernel %.4f ms (min %.4f)" % (os.environ.get("LISEC_LIB_PATH", "default"),
                                                              a.elapsed_time(b) / steps, float(np.mean(kms)), min(kms)))

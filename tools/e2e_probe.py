"""End-to-end front-end step (pinned host points in, counts out; bench.py's `e2e` loop) for the current library settings.
    LISEC_H2D_PIECE_BYTES=... python tools/e2e_probe.py [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
host = []
for b in range(4):
    sw = [synth.lyft_like_sweep(100_000, seed=8 * b + s) for s in range(8)]
    host.append(torch.from_numpy(np.concatenate(sw)).pin_memory())
off = [100_000 * i for i in range(9)]
fe = Frontend(device=0, max_points=800_000, max_sweeps=8, grid_dtype="f32")
fe.set_weights(synthetic_vfe_pack(0))
grid = fe.new_grid(8)
DEPTH = 2
pinned = [torch.empty(fe.COUNTS_BYTES, dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]
ready = [torch.cuda.Event() for _ in range(DEPTH)]


def step(i):
    fe.forward_host(host[i % 4], off, out=grid)
    fe.counts_async(pinned[i % DEPTH])
    ready[i % DEPTH].record()
    if i >= DEPTH - 1:
        ready[(i - DEPTH + 1) % DEPTH].synchronize()


for i in range(10):
    step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    step(i)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print("piece %s: e2e step %.4f ms = %.0f sweeps/s" % (os.environ.get("LISEC_H2D_PIECE_BYTES", "default"), ms, 8e3 / ms))

"""configs[4] training step (2 sweeps x 100 k points, full grid) eager against CUDA-graph replay of its dense region:
wall-clock per step over a synchronised loop = what a user sees.    python tools/train_step_probe.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import synth  # noqa: E402
from lisec_b200.train import TrainStep  # noqa: E402
from lisec_b200.weights import keras_default_init_pack  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = 2
sweeps = [synth.lyft_like_sweep(100_000, seed=s) for s in range(B)]
pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
off = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
g = torch.Generator().manual_seed(0)
yc = torch.randint(0, 2, (B, 100, 200, 2), generator=g).float().cuda()
yr = (torch.randn((B, 100, 200, 14), generator=g) * 0.3).cuda()
modes = (False, True) if len(sys.argv) <= 2 else (sys.argv[2] == "graph",)
for mode in modes:
    ts = TrainStep(keras_default_init_pack(0), batch=B, max_points=len(pts), use_graph=mode)
    for _ in range(3):
        ts.step(pts, off, yc, yr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = ts.step(pts, off, yc, yr)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print("use_graph=%s: %.3f ms per step (%.1f sweeps/s), loss %.5f" % (mode, ms, B * 1e3 / ms, float(loss)))
    ts.close()
    del ts
    torch.cuda.empty_cache()

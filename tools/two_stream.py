"""Experiment: two Frontend handles on two streams, batches alternating — does the grouping chain of batch i+1 find room
beside the fused VFE kernel of batch i? LISEC_VFE_CTAS=n leaves 148 - n SMs without a persistent VFE CTA."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.weights import synthetic_vfe_pack

batches = [synth.sweep_batch(8, 100_000, seed0=8 * b) for b in range(4)]
dev = [torch.from_numpy(p).cuda() for p, _ in batches]
off = batches[0][1]
pack = synthetic_vfe_pack(0)
fes = [Frontend(max_points=800_000, max_sweeps=8) for _ in range(2)]
for fe in fes:
    fe.set_weights(pack)
grids = [fes[0].new_grid(8), fes[1].new_grid(8)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(n, two):
    for i in range(n):
        k = i % 2 if two else 0
        with torch.cuda.stream(streams[k]):
            fes[k].forward(dev[i % 4], off, out=grids[k])


for two in (False, True, False, True):
    run(10, two)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    run(n, two)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s: %.4f ms per step, %.0f sweeps/s" % ("two streams" if two else "one stream ", dt / n * 1e3, 8 * n / dt))

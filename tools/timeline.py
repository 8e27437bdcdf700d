"""In-step timeline of the kernel chain (LISEC_TRACE=1): every kernel stamps the moment its predecessor completed
(the return of its griddepcontrol.wait), so consecutive differences are the kernels' in-step durations, warm caches
and launch overlap included — which an ncu launch list (serialised, cold) cannot show."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LISEC_TRACE"] = "1"
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402

NAMES = ["point_pass", "scan_reduce", "scan_down", "fill_pass", "order_pass", "(unused)", "vfe_kernel<1>", "end"]
batches = [synth.sweep_batch(8, 100_000, seed0=8 * b) for b in range(3)]
fe = Frontend(device=0, max_points=800_000, max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
dev = [torch.from_numpy(p).cuda() for p, _ in batches]
off = batches[0][1]
grid = fe.new_grid(8)
buf = np.zeros(256 * 16, dtype=np.int64)


def read():
    fe._check(fe._lib.lisec_debug_trace(fe._h, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size))
    t = buf.reshape(256, 16)[200:208].astype(np.uint64)
    return t[:, 0].astype(np.float64), t[:, 1].astype(np.float64)


for i in range(4):
    fe.forward(dev[i % 3], off, out=grid)
torch.cuda.synchronize()
read()  # re-arm
acc = np.zeros(8)
first = np.zeros(2)  # the EARLIEST CTA's writer / VFE-pipeline end
n = 10
for i in range(n):
    for j in range(3):  # three back-to-back steps; the stamps keep the min/max, so measure them one at a time
        fe.forward(dev[(i + j) % 3], off, out=grid)
    torch.cuda.synchronize()
    read()
    fe.forward(dev[i % 3], off, out=grid)  # the measured step follows 3 others immediately (warm, queued)
    torch.cuda.synchronize()
    lo, hi = read()
    start = lo.copy()
    start[7] = hi[7]
    start[5] = hi[5]  # the writers' end (latest CTA)
    acc += start - start[0]
    first += np.array([lo[5], lo[7]]) - start[0]
acc /= n
first /= n
print("in-step timeline, us after the point pass started (mean of %d steps):" % n)
for k in (0, 1, 2, 3):
    print("  %-14s starts %7.1f   runs %6.1f" % (NAMES[k], acc[k] / 1e3, (acc[k + 1] - acc[k]) / 1e3))
print("  %-14s starts %7.1f   runs %6.1f" % (NAMES[4], acc[4] / 1e3, (acc[6] - acc[4]) / 1e3))
print("  %-14s starts %7.1f" % (NAMES[6], acc[6] / 1e3))
print("  background writers done (first / last CTA) %7.1f / %7.1f" % (first[0] / 1e3, acc[5] / 1e3))
print("  VFE pipeline done (first / last CTA)       %7.1f / %7.1f" % (first[1] / 1e3, acc[7] / 1e3))

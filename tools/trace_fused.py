"""Cycle counters of the VFE kernel's pipeline stages (LISEC_TRACE=1): where the front / back warps spend their time.
Usage: python tools/trace_fused.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LISEC_TRACE"] = "1"
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402

NAMES = ["front total", "VFE-1", "pool1", "Q2", "dense_1", "wait X free", "X pointwise", "pool2", "X pooled",
         "wait acc free", "info + MMA issue", "back total", "back wait full", "back scan", "tiles", "-"]
pts, off = synth.sweep_batch(8, 100_000, seed0=0)
fe = Frontend(device=0, max_points=len(pts), max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(8)


def dump(tag):
    buf = np.zeros(256 * 16, dtype=np.int64)
    fe._check(fe._lib.lisec_debug_trace(fe._h, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size))
    t = buf.reshape(256, 16)[:148].astype(np.float64)
    tiles = t[:, 14].mean()
    print("== %s: %.1f tiles per CTA; cycles per tile (mean over CTAs)" % (tag, tiles))
    for i, n in enumerate(NAMES[:14]):
        print("   %-18s %9.0f" % (n, (t[:, i] / np.maximum(t[:, 14], 1)).mean()))


for i in range(3):
    fe.forward(dev, off, out=grid)
torch.cuda.synchronize()
dump("fused forward")
fe.voxelize(dev, off)
feat = fe.vfe()
fe.vfe(out=feat)
torch.cuda.synchronize()
dump("VFE rows only (MODE 0)")

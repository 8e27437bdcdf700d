"""Per-CTA timeline of the two concurrent kernels of the fused path (LISEC_TRACE=1): where the background writer's and
the VFE kernel's CTAs ran and when. Usage: LISEC_TRACE=1 python tools/trace_fused.py [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LISEC_TRACE"] = "1"
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pts, off = synth.sweep_batch(8, 100_000, seed0=0)
fe = Frontend(device=0, max_points=len(pts), max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(8)


def dump(tag):
    buf = np.zeros(2 * 256 * 4, dtype=np.uint64)
    fe._check(fe._lib.lisec_debug_trace(fe._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size))
    t = buf.reshape(2, 256, 4)[:, :148].astype(np.int64)
    t0 = min(t[0, :, 1].min(), t[1, :, 1].min())
    print("== %s  (us relative to the first CTA start)" % tag)
    for k, name in enumerate(["background", "vfe"]):
        sm, a, b = t[k, :, 0], (t[k, :, 1] - t0) / 1e3, (t[k, :, 2] - t0) / 1e3
        per_sm = np.bincount(sm, minlength=148)
        print("  %-10s start %7.1f..%7.1f  end %7.1f..%7.1f  dur mean %7.1f max %7.1f | SMs used %d, max CTAs on one SM %d"
              % (name, a.min(), a.max(), b.min(), b.max(), (b - a).mean(), (b - a).max(), int((per_sm > 0).sum()),
                 int(per_sm.max())))


for i in range(steps):
    fe.forward(dev, off, out=grid)
    torch.cuda.synchronize()
    dump("forward #%d (synchronised before and after)" % i)
for i in range(4):
    fe.forward(dev, off, out=grid)
torch.cuda.synchronize()
dump("last of 4 back-to-back forwards")
fe.voxelize(dev, off)
torch.cuda.synchronize()
fe.vfe_scatter_fused(out=grid)
torch.cuda.synchronize()
dump("vfe_scatter_fused alone")
print("last_background_ms", fe.last_background_ms)

"""Front end and whole inference step on both graphs the .h5 may hold, device-timed with CUDA events:
the tensor-core VFE kernel (current graph), the float32 kernel on the same graph (LISEC_GENERIC_VFE=1) and on the older
graph model.png shows. 8 sweeps x 100 k points, bf16 grid.

    python tools/arch_bench.py [steps]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.network import DenseNetwork  # noqa: E402
from lisec_b200.weights import CURRENT, MODEL_PNG, synthetic_model_pack  # noqa: E402


def timed(fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    only = sys.argv[2] if len(sys.argv) > 2 else ""  # substring of the case name
    sweeps = [synth.lyft_like_sweep(100_000, seed=s) for s in range(8)]
    pts = torch.from_numpy(np.concatenate(sweeps)).cuda()
    off = np.cumsum([0] + [len(s) for s in sweeps]).tolist()
    out = {}
    for name, arch, env in (("current graph, tensor-core kernel", CURRENT, False),
                            ("current graph, float32 kernel", CURRENT, True),
                            ("older graph (16|64|128, Dense-BN-Dense), float32 kernel", MODEL_PNG, True)):
        if only not in name:
            continue
        if env:
            os.environ["LISEC_GENERIC_VFE"] = "1"
        else:
            os.environ.pop("LISEC_GENERIC_VFE", None)
        pack = synthetic_model_pack(0, arch)
        fe = Frontend(device=0, max_points=len(pts), max_sweeps=8, grid_dtype="bf16", widths=arch.widths,
                      post_dense=arch.post_dense)
        fe.set_weights(pack)
        net = DenseNetwork(pack, batch=8, arch=arch)
        fe.voxelize(pts, off)
        rows = torch.empty((fe.counts()[1], arch.c3), dtype=torch.float32, device="cuda")
        r = {
            "frontend_ms": timed(lambda: fe.forward(pts, off, out=net.grid), steps),
            "vfe_rows_ms": timed(lambda: fe.vfe(out=rows), steps),
            "network_ms": timed(lambda: net.forward(), steps),
            "grid_MB": net.grid.numel() * 2 / 1e6,
            "network_gflop": net.flops / 1e9,
        }
        r["step_ms"] = timed(lambda: (fe.forward(pts, off, out=net.grid), net.forward()), steps)
        r["sweeps_per_s"] = 8e3 / r["step_ms"]
        out[name] = r
        print("%-60s front end %.3f ms (VFE rows %.3f)  network %.3f ms  step %.3f ms = %.0f sweeps/s" %
              (name, r["frontend_ms"], r["vfe_rows_ms"], r["network_ms"], r["step_ms"], r["sweeps_per_s"]))
        net.close()
        fe.close()
        del net, fe
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/arch_bench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

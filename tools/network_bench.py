"""Timing and error of the dense network (middle Conv3D + RPN + heads) on one GPU: per-layer CUDA-event times at the
full 8 x 200 x 400 grid, TFLOP/s, and the bf16 error against the float32/float64 torch-CPU oracle on one sweep."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200.network import DenseNetwork  # noqa: E402
from lisec_b200.weights import synthetic_network_pack  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
check = len(sys.argv) > 2 and sys.argv[2] == "check"
pack = synthetic_network_pack(0)
sched = os.environ.get("NET_SCHED")  # e.g. "1,1" = (m_tiles, group_kh) for every layer that accepts it
if sched == "halo":
    from lisec_b200.network import halo_schedule

    net = DenseNetwork(pack, batch=batch, schedule=halo_schedule)
elif sched:
    from lisec_b200.network import default_schedule

    want = tuple(int(x) for x in sched.split(","))
    net = DenseNetwork(pack, batch=batch, schedule=lambda *a: [want] + default_schedule(*a) + [(1, 0)])
else:
    net = DenseNetwork(pack, batch=batch)
g = torch.Generator(device="cpu").manual_seed(3)
grid = torch.rand((1, 8, 200, 400, 64), generator=g).to(torch.bfloat16)
for b in range(batch):
    net.grid[b].copy_(grid[0])
for _ in range(2):
    net.forward()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(net.layers) + 1)]
reps = 5
acc = np.zeros(len(net.layers))
for _ in range(reps):
    ev[0].record()
    for i in range(len(net.layers)):
        net.run_layers(i, i + 1)
        ev[i + 1].record()
    torch.cuda.synchronize()
    acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(len(net.layers))]
acc /= reps
tot = 0.0
for L, ms in zip(net.layers, acc):
    d = L.desc
    od = (d.in_d + 2 * d.pad_d - d.kd) // d.stride_d + 1
    oh = (d.in_h + 2 * d.pad_h - d.kh) // d.stride_hw + 1
    ow = (d.in_w + 2 * d.pad_w - d.kw) // d.stride_hw + 1
    fl = 2.0 * d.batch * od * oh * ow * d.kd * d.kh * d.kw * d.in_c * d.out_c * d.n_tiles
    tot += ms
    print("%-20s %8.3f ms  %7.1f TFLOP/s  tile %dx%d x%d%s" % (L.name, ms, fl / ms / 1e9, d.tile_w, d.tile_h, d.m_tiles,
                                                               (" kh-halo" if d.group_kh == 1 else " halo" if d.group_kh == 2 else "")))
t0 = torch.cuda.Event(enable_timing=True)
t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    net.forward()
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
print("network forward, batch %d: %.3f ms (%.3f ms/sweep), %.1f TFLOP/s algorithmic; sum of layers %.3f ms"
      % (batch, ms, ms / batch, net.flops / ms / 1e9, tot))
if check:
    from oracle import network_oracle as NO

    prob, reg = net.forward()
    torch.cuda.synchronize()
    t = time.time()
    wp, wr = NO.network_forward(grid.float().numpy(), pack, dtype=torch.float32)
    print("oracle (torch CPU float32, %d threads): %.1f s per sweep" % (torch.get_num_threads(), time.time() - t))
    for name, got, want in (("prob", prob[0].cpu().numpy(), wp[0]), ("regress", reg[0].cpu().numpy(), wr[0])):
        err = np.abs(got.astype(np.float64) - want)
        rms = np.sqrt(np.mean(want.astype(np.float64) ** 2))
        print("%-8s max|err| %.3e  / max|ref| = %.3e   / rms = %.3e   rel-L2 %.3e   elementwise(max(|ref|,rms)) %.3e"
              % (name, err.max(), err.max() / np.abs(want).max(), err.max() / rms,
                 np.sqrt((err ** 2).sum() / (want.astype(np.float64) ** 2).sum()),
                 (err / np.maximum(np.abs(want), rms)).max()))

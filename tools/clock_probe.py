"""SM clock and board power while one leg of the front end runs back to back for ~2 s (NVML, 2 ms polling)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pynvml
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.weights import synthetic_vfe_pack

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
pts, off = synth.sweep_batch(8, 100_000, seed0=0)
fe = Frontend(max_points=len(pts), max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(8)
far = torch.full((len(pts), 3), 1000.0, device="cuda")
fe.voxelize(dev, off)
feat = fe.vfe()


def run(name, fn, seconds=2.0):
    clocks, power, stop = [], [], threading.Event()

    def poll():
        while not stop.is_set():
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.002)

    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t = threading.Thread(target=poll)
    t.start()
    n, t0 = 0, time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    stop.set()
    t.join()
    print("%-28s %.3f ms/iter   clock median %d min %d MHz   power median %.0f max %.0f W" %
          (name, a.elapsed_time(b) / n, np.median(clocks), min(clocks), np.median(power), max(power)))


run("forward (fused)", lambda: fe.forward(dev, off, out=grid))
run("forward (all dropped)", lambda: fe.forward(far, off, out=grid))
run("vfe rows only (MODE 0)", lambda: fe.vfe(out=feat))
run("grid_write alone", lambda: fe.scatter(feat, out=grid))
run("voxelize", lambda: fe.voxelize(dev, off))

"""Front-end step through eager launches against a CUDA-graph replay of the same six kernels (8 sweeps x 100 k points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lisec_b200 import Frontend, synth
from lisec_b200.weights import synthetic_vfe_pack

pts, off = synth.sweep_batch(8, 100_000, seed0=0)
fe = Frontend(max_points=len(pts), max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
dev = torch.from_numpy(pts).cuda()
grid = fe.new_grid(8)
side = torch.cuda.Stream()


def timed(fn, n=300):
    for _ in range(20):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.current_stream().synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n


with torch.cuda.stream(side):
    eager = timed(lambda: fe.forward(dev, off, out=grid))
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fe.forward(dev, off, out=grid)
    replay = timed(g.replay)
    g4 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g4, stream=side):
        for _ in range(4):
            fe.forward(dev, off, out=grid)
    replay4 = timed(g4.replay, 75) / 4
print("eager %.4f ms per step, graph replay %.4f ms, graph of 4 steps %.4f ms per step" % (eager, replay, replay4))

"""Worker of tests/test_gpu_pdl.py: a history of calls through every kernel chain that is launched with programmatic
dependent launch, printing digests of what each call produced. Run once with LISEC_NO_PDL=1 (plain stream order) and once
without: the digests must be identical."""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.network import DenseNetwork  # noqa: E402
from lisec_b200.train import TrainStep  # noqa: E402
from lisec_b200.weights import synthetic_model_pack, synthetic_network_pack, synthetic_vfe_pack  # noqa: E402


def digest(*tensors):
    h = hashlib.sha1()
    for t in tensors:
        a = t.detach().cpu().contiguous()
        h.update(a.view(torch.uint8).numpy().tobytes() if a.dtype == torch.bfloat16 else a.numpy().tobytes())
    return h.hexdigest()


out = {}
# ---- front end: device points and host points (alternating staging buffers), the same grid buffer every call ----
batches = []
for b in range(3):
    sw = [synth.lyft_like_sweep(30_000 + 5_000 * b, seed=10 * b + s) for s in range(2)]
    batches.append((np.concatenate(sw), [0, len(sw[0]), len(sw[0]) + len(sw[1])]))
fe = Frontend(device=0, max_points=80_000, max_sweeps=2, grid_dtype="f32")
fe.set_weights(synthetic_vfe_pack(3))
grid = fe.new_grid(2)
dev = [torch.from_numpy(p).cuda() for p, _ in batches]
host = [torch.from_numpy(p).pin_memory() for p, _ in batches]
d = []
for i in range(7):
    fe.forward(dev[i % 3], batches[i % 3][1], out=grid)
    vs = fe.export()
    d.append(digest(grid, vs.coords, vs.counts, vs.point_idx))
out["frontend_device"] = d
d = []
for i in range(7):
    fe.forward_host(host[(2 * i) % 3], batches[(2 * i) % 3][1], out=grid)
    d.append(digest(grid))
out["frontend_host"] = d
fe.close()

# ---- dense inference network: two different grids through the same plans and buffers ----
nx, ny = 24, 40
net = DenseNetwork(synthetic_network_pack(1), batch=2, nx=nx, ny=ny)
g = torch.Generator().manual_seed(0)
grids = [torch.rand((2, 8, nx, ny, 64), generator=g).to(torch.bfloat16).cuda() for _ in range(2)]
d = []
for i in range(5):
    p, r = net.forward(grids[i % 2])
    d.append(digest(p.contiguous(), r.contiguous()))
out["network"] = d
net.close()

# ---- the training step: different batches step after step, every buffer reused ----
pack = {k: np.asarray(v, np.float32) for k, v in synthetic_model_pack(4).items()}
rng = np.random.default_rng(2)
tb = []
for i in range(4):
    clouds = [rng.uniform([-5.9, -4.9, 0.26], [5.9, 4.9, 1.99], size=(1500 + 200 * i + 100 * s, 3)).astype(np.float32) for s in range(2)]
    gl = torch.Generator().manual_seed(i)
    tb.append((np.concatenate(clouds), np.cumsum([0] + [len(c) for c in clouds]).tolist(),
               torch.randint(0, 3, (2, nx // 2, ny // 2, 2), generator=gl).float().cuda(),
               (torch.randn((2, nx // 2, ny // 2, 14), generator=gl) * 0.5).cuda()))
step = TrainStep(pack, batch=2, max_points=max(len(b[0]) for b in tb), nx=nx, ny=ny, nz=8, lr=0.002)
d = []
for i in range(6):
    step.step(*tb[i % 4])
    d.append(digest(step.store.var))
out["train"] = d
step.close()
print("PDL_WORKER " + json.dumps(out))

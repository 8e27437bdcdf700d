"""ncu target: rpnToRegion (decode_kernel + nms_kernel) on 8 synthetic head tensors, and the lidar ingest kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lisec_b200 import synth
from lisec_b200.decode import RegionDecoder
from lisec_b200.ingest import LidarIngest

dec = RegionDecoder()
heads = torch.from_numpy(np.stack([np.concatenate(synth.synthetic_rpn_output(100 + s), axis=-1) for s in range(8)])).cuda()
for _ in range(3):
    picks, n_picks, boxes, probs = dec.regions(heads[..., :2], heads[..., 2:])
torch.cuda.synchronize()
print("picks per sample", n_picks.cpu().tolist())
ing = LidarIngest()
n = 800_000
rec = torch.randn((n, 5), device="cuda")
off = np.linspace(0, n, 25).astype(np.int64)
poses = ing.make_poses([[0.99, 0.01, -0.02, 0.1]] * 24, [[1.0, 0.2, 1.8]] * 24)
out = torch.empty((n, 3), dtype=torch.float64, device="cuda")
for _ in range(3):
    ing.transform(rec, off, out=out, poses=poses)
torch.cuda.synchronize()
print("ok")

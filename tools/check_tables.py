"""Replays the voxelize history of tests/test_gpu_parity.py and checks the tile / row tables on the host."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from lisec_b200 import Frontend, synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402

SLOTS = 12


def table(fe, which, n):
    out = np.zeros(n, np.int32)
    fe._check(fe._lib.lisec_debug_table(fe._h, which, out.ctypes.data_as(C.POINTER(C.c_int32)), n))
    return out


def check(fe, tag):
    per, V, nin, noor, nnf = fe.counts()
    rs = table(fe, 0, V + 1)
    rows = int(rs[V])
    rv = table(fe, 1, rows) & ~(1 << 30)
    assert (np.diff(rs) >= 2).all() or V == 0
    want = np.repeat(np.arange(V), np.diff(rs))
    bad_rv = int((rv != want).sum())
    n_chunks = (rows - 1) // 478 + 1 if V else 0  # (may name one trailing chunk no voxel starts in: zero tiles)
    nt = table(fe, 4, max(n_chunks, 1))[:n_chunks]
    tf = table(fe, 2, max(n_chunks, 1) * SLOTS).reshape(-1, SLOTS)[:n_chunks]
    tr = table(fe, 3, max(n_chunks, 1) * SLOTS).reshape(-1, SLOTS)[:n_chunks]
    bad = 0
    prev_end = 0
    for c in range(n_chunks):
        n = int(nt[c])
        if not (0 <= n <= SLOTS - 1) or tf[c, 0] != prev_end or (n == 0 and c != n_chunks - 1):
            bad += 1
            continue
        for j in range(n + 1):
            v = int(tf[c, j])
            if not (0 <= v <= V) or tr[c, j] != rs[v]:
                bad += 1
        for j in range(n):
            if not (0 < tr[c, j + 1] - tr[c, j] <= 128):
                bad += 1
        prev_end = int(tf[c, n])
    if n_chunks and prev_end != V:
        bad += 1
    print("%-28s V %7d rows %7d chunks %5d tiles %5d  bad row_voxel %d  bad tile entries %d" %
          (tag, V, rows, n_chunks, int(nt.sum()) if n_chunks else 0, bad_rv, bad))
    return bad + bad_rv


fe = Frontend(device=0, max_points=1_100_000, max_sweeps=8)
fe.set_weights(synthetic_vfe_pack(0))
total = 0
clouds = [("lyft100k", synth.lyft_like_sweep(100_000, seed=0)), ("adversarial", synth.adversarial_tail()),
          ("ragged", synth.lyft_like_sweep(20_001, seed=1)),
          ("f64", np.random.default_rng(4).uniform([-52, -52, -0.3], [52, 52, 2.3], size=(50_000, 3))),
          ("saturated300k", synth.saturated_cloud(300_000, n_sweeps=3, theta=2.0)),
          ("saturated200k", synth.saturated_cloud(200_000, n_sweeps=2, theta=2.5)),
          ("saturated200k", synth.saturated_cloud(200_000, n_sweeps=2, theta=2.5)),
          ("dropped", np.asarray([[1e3, 0, 1.0], [0, 0, 0.1]], np.float32)),
          ("lyft100k again", synth.lyft_like_sweep(100_000, seed=0))]
for rep in range(2):
    for tag, pts in clouds:
        fe.voxelize(pts, [0, len(pts)])
        total += check(fe, tag)
        if tag.startswith("lyft"):
            got = fe.vfe()
            torch.cuda.synchronize()
            print("   vfe ok, finite:", bool(torch.isfinite(got).all()))
print("TOTAL BAD", total)

"""The N > 1 path on CPU: world_size-2 gloo processes shard a batch of sweeps with no data-path collective; only the
scalar max/sum reductions bench.py uses go through torch.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lisec_b200.sharding import max_over_ranks, shard_offsets, shard_range, sum_over_ranks


def test_shard_ranges_partition_the_sweeps():
    for n in (0, 1, 7, 8, 16, 33):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [e - b for b, e in got]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import lisec_oracle as O  # the checker: each rank groups ITS sweeps, nothing is exchanged

    rng = np.random.default_rng(0)  # same batch on every rank
    sizes = [300, 500, 200, 400, 100]
    pts = np.concatenate([rng.uniform([-6, -5, 0], [6, 5, 2], size=(n, 3)) for n in sizes])
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    p0, p1, off = shard_offsets(offsets, world, rank)
    ref = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=12, maxVoxelY=20, maxVoxelZ=8)
    n_vox = 0
    for a, b in zip(off[:-1], off[1:]):
        n_vox += len(O.voxelize_np(pts[p0 + a:p0 + b], **ref)["counts"])
    total = sum_over_ranks(float(n_vox))
    slow = max_over_ranks(10.0 + rank)
    dist.barrier()
    q.put((rank, p0, p1, n_vox, total, slow))
    dist.destroy_process_group()


def test_two_gloo_ranks_shard_a_batch_without_exchanging_data():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, v0, t0, s0), (r1, a1, b1, v1, t1, s1) = out
    assert (a0, b0, a1, b1) == (0, 1000, 1000, 1500)  # 3 + 2 sweeps, contiguous, every point owned once
    assert t0 == t1 == v0 + v1  # the only communication: a scalar sum ...
    assert s0 == s1 == 11.0     # ... and the max over ranks of a time
    # the shards' voxel counts equal the single-process counts of the same sweeps
    from oracle import lisec_oracle as O

    rng = np.random.default_rng(0)
    sizes = [300, 500, 200, 400, 100]
    ref = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=12, maxVoxelY=20, maxVoxelZ=8)
    alone = [len(O.voxelize_np(rng.uniform([-6, -5, 0], [6, 5, 2], size=(n, 3)), **ref)["counts"]) for n in sizes]
    assert v0 == sum(alone[:3]) and v1 == sum(alone[3:])

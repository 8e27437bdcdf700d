"""Multi-GPU plumbing of the path: sweeps are independent units, so ranks shard them with NO data-path
collective. One process per GPU; torch.distributed is used only for the timing barrier and the max / sum over ranks of
scalars (NCCL on the GPUs, gloo in the CPU tests). The reference has no counterpart: it loops over samples serially
(Predict.py:17, model_training.py:266)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_units` sweeps for `rank`: sizes differ by at most one, earlier ranks take the extra."""
    if not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard request: %d units, rank %d of %d" % (n_units, rank, world))
    q, r = divmod(n_units, world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def shard_offsets(sweep_offsets: Sequence[int], world: int, rank: int) -> Tuple[int, int, List[int]]:
    """(first point, last point, rebased sweep_offsets) of this rank's sweeps inside a concatenated batch."""
    n = len(sweep_offsets) - 1
    b, e = shard_range(n, world, rank)
    p0 = int(sweep_offsets[b])
    return p0, int(sweep_offsets[e]), [int(o) - p0 for o in sweep_offsets[b:e + 1]]


def _reduce(x: float, op, device) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(x: float, device="cpu") -> float:
    """Device times are reported as the max over ranks (never wall clock)."""
    return _reduce(x, dist.ReduceOp.MAX, device)


def sum_over_ranks(x: float, device="cpu") -> float:
    return _reduce(x, dist.ReduceOp.SUM, device)

"""Multi-GPU plan for the MLT-CNN path: the work shards over INDEPENDENT units (frames of a throughput run, or
whole encodes = sequence x QP of a CTC sweep) with no data-path collective -- each rank owns one
GPU, one mlt_ctx and a disjoint set of units; results are only gathered at the end (host side, any backend).

The reference authors ran exactly this by hand (`CUDA_VISIBLE_DEVICES=1 ./X_enc.sh`, script_128/archive)."""
from __future__ import annotations

from typing import Sequence


def shard_range(n_units: int, world: int, rank: int) -> range:
    """Contiguous, balanced split of units 0..n-1 (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def assign_encodes(costs: Sequence[float], world: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of independent encodes (cost ~ width*height*frames) to `world`
    GPUs; returns the job indices per rank.  Deterministic (ties broken by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return out


def gather_rows(local_rows, world: int, rank: int):
    """Gather per-rank numpy result rows on every rank via torch.distributed (gloo on CPU, nccl on GPU boxes).
    Only used after the timed region; the hot path itself never communicates."""
    import numpy as np
    import torch
    import torch.distributed as dist

    if world == 1:
        return np.asarray(local_rows)
    objs = [None] * world
    dist.all_gather_object(objs, np.asarray(local_rows))
    return np.concatenate(objs, 0)

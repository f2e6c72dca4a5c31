"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs; the
same host logic runs under gloo in the CPU tests).

* walks shard by contiguous ranges of GLOBAL walk ids -- the partitioning main_link.py:263-264
  applies across its process pool -- with no communication; Philox counters are keyed by the
  global id, so the corpus does not depend on the number of ranks;
* SGNS shards the walk corpus; vocabulary counts are summed once (all_reduce of int64[N]); the
  replicated syn0 / syn1neg tables are averaged every K steps (all_reduce(sum) * 1/world).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int, world_size: int):
    """Contiguous chunk `rank` of `total` items, chunk = ceil(total / world) as
    main_link.py:263-264 (`range(0, len, ceil(len/num_pool))`)."""
    per = int(math.ceil(float(total) / world_size)) if total else 0
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def step_walk_ids(step: int, rank: int, world_size: int, batch: int):
    """Global id of the first walk rank `rank` simulates in step `step` (weak scaling: every rank
    takes `batch` consecutive ids of the step's world*batch block)."""
    return (step * world_size + rank) * batch


def sum_counts(counts: torch.Tensor) -> torch.Tensor:
    if world()[1] > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def average_tables(*tables: torch.Tensor):
    """Parameter averaging of replicated tables, in place."""
    _, w = world()
    if w == 1:
        return
    for t in tables:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.mul_(1.0 / w)

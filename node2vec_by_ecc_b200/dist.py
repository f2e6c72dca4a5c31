"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs; the
same host logic runs under gloo in the CPU tests).

* walks shard by contiguous ranges of GLOBAL walk ids -- the partitioning main_link.py:263-264
  applies across its process pool -- with no communication; Philox counters are keyed by the
  global id, so the corpus does not depend on the number of ranks;
* SGNS shards the walk corpus; vocabulary counts are summed once (all_reduce of int64[N]); the
  replicated syn0 / syn1neg tables are combined by DELTA-SUM every few thousand walks:
  new = base + sum over ranks of (replica - base). Plain parameter averaging (all_reduce * 1/world,
  the scheme SURVEY.md 8e sketches) divides every row's update by `world` because a row is usually
  touched by one replica per interval; measured on a 1 M-node planted graph it loses 0.04 AUC at 2
  replicas and collapses at 8, whatever the interval, while delta-sum stays within +-0.005 of the
  single-replica run when the interval obeys total pairs per sync <= ~100 * V / world
  (scripts/auc_multi_replica_large.py, profiles/r01_o_multi_replica_auc.txt).
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


def pool_plan(total: int, world_size: int, pool_walks: int):
    """How a corpus of `total` walks, cut into `world_size` contiguous shares (shard_range), is trained by
    the block-partitioned trainer: -> (per, pools) with per = walks per rank after padding every share to
    one length (empty walks add no pairs) and pools = [(first walk of the share, walks per rank)], the same
    on every rank -- all ranks must run the same number of pools of the same size, because every pool ends
    in collectives. Pool p covers the global example range [first * world, (first + n) * world)."""
    per = int(math.ceil(float(total) / world_size)) if total else 0
    pool = int(max(1, min(per, pool_walks))) if per else 1
    return per, [(p0, min(pool, per - p0)) for p0 in range(0, per, pool)]


def sum_counts(counts: torch.Tensor) -> torch.Tensor:
    if world()[1] > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def sync_walks_per_rank(vocab_size: int, world_size: int, pairs_per_walk: float, factor: float = 100.0) -> int:
    """Walks one rank may train on between two delta-sum syncs while the replicas stay within the
    +-0.005 AUC band: total pairs per sync <= factor * V / world."""
    if world_size <= 1:
        return 1 << 62
    return max(256, int(factor * vocab_size / (world_size * world_size) / max(pairs_per_walk, 1.0)))


class ReplicaSync:
    """Delta-sum synchronisation of replicated tables: after sync() every rank holds
    base + sum_r (replica_r - base), computed as all_reduce(sum of replicas) - (world-1) * base."""

    def __init__(self, *tables: torch.Tensor):
        self.tables = tables
        self.world = world()[1]
        self.bases = [t.clone() for t in tables] if self.world > 1 else []

    def sync(self):
        if self.world == 1:
            return
        for t, b in zip(self.tables, self.bases):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.add_(b, alpha=-(self.world - 1))
            b.copy_(t)


def average_tables(*tables: torch.Tensor):
    """Plain parameter averaging, in place (kept for comparison; see the module docstring for why
    ReplicaSync is what the trainers use)."""
    _, w = world()
    if w == 1:
        return
    for t in tables:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.mul_(1.0 / w)


# ---- one logical table pair across the GPUs of a node (peer memory over NVLink) -------------------
def _ipc_export(t: torch.Tensor):
    """(ipc handle bytes, byte offset of t inside its cudaMalloc block)"""
    from cuda.bindings import driver as cu, runtime as cudart
    err, h = cudart.cudaIpcGetMemHandle(t.data_ptr())
    if int(err) != 0:
        raise RuntimeError("cudaIpcGetMemHandle failed: %s" % err)
    err, base, _size = cu.cuMemGetAddressRange(t.data_ptr())
    if int(err) != 0:
        raise RuntimeError("cuMemGetAddressRange failed: %s" % err)
    return bytes(bytearray(h.reserved)), int(t.data_ptr()) - int(base)


def _ipc_import(handle: bytes, offset: int) -> int:
    from cuda.bindings import runtime as cudart
    h = cudart.cudaIpcMemHandle_t()
    h.reserved = list(handle)
    err, ptr = cudart.cudaIpcOpenMemHandle(h, cudart.cudaIpcMemLazyEnablePeerAccess)
    if int(err) != 0:
        raise RuntimeError("cudaIpcOpenMemHandle failed: %s" % err)
    return int(ptr) + offset


def exchange_peer_pointers(local: torch.Tensor):
    """Every rank contributes one device tensor; returns the list of world device addresses at
    which THIS rank can read/write rank r's tensor (its own for r == rank, NVLink peer mappings
    otherwise). The tensors must stay alive until every rank is done with them."""
    rank, w = world()
    if w == 1:
        return [int(local.data_ptr())]
    mine = _ipc_export(local)
    everyone = [None] * w
    dist.all_gather_object(everyone, mine)
    return [int(local.data_ptr()) if r == rank else _ipc_import(*everyone[r]) for r in range(w)]


# ---- block-partitioned SGNS (word2vec.BlockSgnsTrainer) --------------------------------------------
def bucket_of(rank: int, sub_step: int, world_size: int) -> int:
    """the syn0 part rank `rank` holds -- and the pair bucket it trains -- in sub-step `sub_step`"""
    return (rank + sub_step) % world_size


def gather_pool(local_walks: torch.Tensor) -> torch.Tensor:
    """every rank's [n, stride] walk buffer -> the pool [world * n, stride] in rank order"""
    rank, w = world()
    if w == 1:
        return local_walks
    local_walks = local_walks.contiguous()
    pool = torch.empty((w * local_walks.shape[0],) + tuple(local_walks.shape[1:]), dtype=local_walks.dtype,
                       device=local_walks.device)
    dist.all_gather_into_tensor(pool, local_walks)
    return pool


def ring_pass(held: torch.Tensor, spare: torch.Tensor):
    """hand `held` to rank - 1 and receive rank + 1's into `spare`; -> (new held, new spare).
    After world_size passes every part is back where it started."""
    rank, w = world()
    if w == 1:
        return held, spare
    ops = [dist.P2POp(dist.isend, held, (rank - 1) % w), dist.P2POp(dist.irecv, spare, (rank + 1) % w)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return spare, held

"""Multi-process check of the block-partitioned SGNS (run under torchrun, one rank per GPU):
 1. with one warp per launch the N-rank run must reproduce, bit for bit, the single-process run with
    all N parts on one device (same buckets, same order inside every bucket);
 2. link-prediction AUC on C2 at full width (main_link.py:519-565 protocol, as scripts/auc_c2.py).
   torchrun --nproc-per-node N scripts/dist_block_check.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import build_neg_samples, roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import BlockSgnsTrainer, DeviceGraph, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, R, L = 10000, 5, 40
lo, hi = synth.planted_edges(n, 333000, seed=42, device=dev)
edges = np.stack([lo.cpu().numpy(), hi.cpu().numpy()], 1).astype(np.int64)
tr_e, te = split_edges(edges)
dg = DeviceGraph.from_coo(tr_e[:, 0], tr_e[:, 1], None, n, undirected=True)
t = dg.build_alias_tables(0.25, 4.0)
starts = torch.arange(n, dtype=torch.int32, device=dev).repeat(R)
neg = build_neg_samples(n, edges, len(te), seed=1)
out = {"world": world}


def train(trainer, walks, per_rank, gw, local_run):
    total = walks.shape[0]
    pool = per_rank * world
    for p0 in range(0, total - pool + 1, pool):
        mine = walks[p0: p0 + pool] if local_run else walks[p0 + rank * per_rank: p0 + (rank + 1) * per_rank]
        trainer.train(mine.contiguous(), None, mine.shape[0], L, total_examples=total, example_base=p0, sent_id_base=p0, sent_per_job=10000 // L, grid_warps=gw)
    trainer.check_overflow()
    return trainer.gather()


# 1. exactness, one warp
walks, _ = dg.walk_alias(t, starts[:4096], L, seed=9)
counts = torch.bincount(walks[walks >= 0].to(torch.int64), minlength=n)
a0, a1 = train(BlockSgnsTrainer(counts, dim=128, seed=3), walks, 512, 1, False)
if rank == 0:
    b0, b1 = train(BlockSgnsTrainer(counts, dim=128, seed=3, local_parts=world), walks, 512, 1, True)
    out["max_abs_diff_vs_one_device"] = [float((a0 - b0).abs().max()), float((a1 - b1).abs().max())]
    out["moved"] = float(b1.abs().max())
# 2. AUC at full width
aucs = []
for seed in (1, 2, 3):
    walks, _ = dg.walk_alias(t, starts, L, seed=seed)
    counts = torch.bincount(walks[walks >= 0].to(torch.int64), minlength=n)
    trn = BlockSgnsTrainer(counts, dim=128, seed=seed)
    s0, _ = train(trn, walks, 2048, None, False)
    emb = np.zeros((n, 128), np.float32); emb[trn.order.cpu().numpy()] = s0.cpu().numpy()
    aucs.append(roc_auc_cosine(emb, te, neg))
out["auc_mean"], out["auc_runs"] = float(np.mean(aucs)), aucs
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

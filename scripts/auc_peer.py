"""Link-prediction AUC of the peer-memory SGNS (one table pair sharded over N GPUs, trained by all
of them over NVLink) on C2, main_link.main protocol (main_link.py:519-565) -- same graph, split and
walks as scripts/auc_c2.py. Run: torchrun --nproc-per-node N scripts/auc_peer.py  (N = 1, 2, 4, 8).
Rank r trains chunk r of every round of N chunks (global walk ids, global alpha progress), the way
gensim's worker threads take jobs from one queue."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import build_neg_samples, roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import DeviceGraph, PeerSgnsTrainer, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
CHUNK = int(os.environ.get("CHUNK", "2048"))
n, R, L = 10000, 5, 40
lo, hi = synth.planted_edges(n, 333000, seed=42, device=dev)
edges = np.stack([lo.cpu().numpy(), hi.cpu().numpy()], 1).astype(np.int64)
tr, te = split_edges(edges)
dg = DeviceGraph.from_coo(tr[:, 0], tr[:, 1], None, n, undirected=True)
t = dg.build_alias_tables(0.25, 4.0)
starts = torch.arange(n, dtype=torch.int32, device=dev).repeat(R)
neg = build_neg_samples(n, edges, len(te), seed=1)
aucs = []
for seed in (1, 2, 3, 4, 5):
    walks, lens = dg.walk_alias(t, starts, L, seed=seed)          # every rank: the whole corpus (tiny here)
    counts = torch.bincount(walks[walks >= 0].to(torch.int64), minlength=n)
    trn = PeerSgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=seed)
    if world > 1:
        dist.barrier()
    total = walks.shape[0]
    for c0 in range(rank * CHUNK, total, world * CHUNK):
        c1 = min(total, c0 + CHUNK)
        trn.train(walks[c0:c1], None, c1 - c0, L, total_examples=total, example_base=c0, sent_id_base=c0, sent_per_job=250)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s0, _ = trn.gather()
    emb = np.zeros((n, 128), np.float32); emb[trn.order.cpu().numpy()] = s0.cpu().numpy()
    aucs.append(roc_auc_cosine(emb, te, neg))
    if rank == 0:
        print("seed", seed, round(aucs[-1], 4), flush=True)
    del trn
    if world > 1:
        dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, "chunk": CHUNK, "auc_mean": float(np.mean(aucs)), "auc_std": float(np.std(aucs)), "runs": aucs}))
if world > 1:
    dist.destroy_process_group()

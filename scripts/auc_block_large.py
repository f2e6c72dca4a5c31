"""Block-partitioned SGNS quality at scale, all parts on one GPU (= what n GPUs compute): planted-partition graph with 1 M nodes / 20 M edges, main_link protocol (50 % of the edges
held out, walks R=5 L=40 p=0.25 q=4 on the rest, d=128, window 10), parts x neg_group x pool size (walks per pool).
   GRID="parts,run,pool;..." python scripts/auc_block_large.py AUC on 1 M held-out edges vs 1 M non-edges."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch
from sklearn.metrics import roc_auc_score
from node2vec_by_ecc_b200 import BlockSgnsTrainer, DeviceGraph, SgnsTrainer, synth
from node2vec_by_ecc_b200._lib import check, lib, ptr, stream

dev = torch.device("cuda", 0)
n, m = int(float(os.environ.get("N", 1e6))), int(float(os.environ.get("M", 2e7)))
lo, hi = synth.planted_edges(n, m, seed=42, max_deg=4000, communities=max(39, n // 1000), device=dev)
g = torch.Generator(device=dev); g.manual_seed(123)
perm = torch.randperm(lo.numel(), device=dev, generator=g)
half = lo.numel() // 2
tr_i, te_i = perm[:half], perm[half:half + 1_000_000]
dg = DeviceGraph.from_coo(lo[tr_i], hi[tr_i], None, n, undirected=True)
pos_a, pos_b = lo[te_i].clone(), hi[te_i].clone()
# non-edges: random pairs not in the edge set
keys = torch.sort(lo.to(torch.int64) * n + hi.to(torch.int64)).values
ra = torch.randint(0, n, (1_300_000,), device=dev, generator=g); rb = torch.randint(0, n, (1_300_000,), device=dev, generator=g)
a_, b_ = torch.minimum(ra, rb), torch.maximum(ra, rb)
k_ = a_ * n + b_
ok = (a_ != b_) & (keys[torch.searchsorted(keys, k_).clamp_(max=keys.numel() - 1)] != k_)
neg_a, neg_b = a_[ok][:1_000_000].to(torch.int32), b_[ok][:1_000_000].to(torch.int32)
R, L = 5, 40
total = R * n
B = 1 << 19
walks = torch.empty((B, L), dtype=torch.int32, device=dev); lens = torch.empty(B, dtype=torch.int32, device=dev)

def walk(g0, nb):
    st = ((g0 + torch.arange(nb, device=dev)) % n).to(torch.int32)
    dg.walk_reject(0.25, 4.0, st, L, 1, g0, out=(walks[:nb], lens[:nb]))

GW = int(os.environ.get("GRID_WARPS", "0")) or None
counts = torch.zeros(n, dtype=torch.int64, device=dev)
for g0 in range(0, total, B):
    nb = min(B, total - g0); walk(g0, nb)
    check(lib().n2v_vocab_count(ptr(walks), C.c_int64(nb * L), C.c_int32(n), ptr(counts), stream()))

def auc_of(tr):
    v = tr.vocab_of_id
    out = torch.empty(2_000_000, dtype=torch.float32, device=dev)
    ia = torch.cat([v[pos_a.long()], v[neg_a.long()]]).contiguous(); ib = torch.cat([v[pos_b.long()], v[neg_b.long()]]).contiguous()
    check(lib().n2v_cosine_pairs(ptr(tr.syn0), C.c_int32(128), ptr(ia), ptr(ib), C.c_int64(ia.numel()), ptr(out), stream()))
    s = out.cpu().numpy()
    y = np.concatenate([np.ones(pos_a.numel()), np.zeros(neg_a.numel())])
    return float(roc_auc_score(y, s))


class _View:
    pass

def run(parts, neg_group, pool):
    tr = BlockSgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=1, local_parts=parts, neg_group=neg_group)
    for p0 in range(0, total, pool):
        p1 = min(total, p0 + pool)
        for g0 in range(p0, p1, B):                 # the pool is trained in walk batches of B (pair buffers stay small)
            nb = min(B, p1 - g0); walk(g0, nb)
            tr.train(walks[:nb], None, nb, L, total_examples=total, example_base=g0, sent_id_base=g0, sent_per_job=10000 // L)
    tr.check_overflow()
    v = _View(); v.vocab_of_id = tr.vocab_of_id; v.syn0, _ = tr.gather()
    return auc_of(v)

grid = os.environ.get("GRID", "1,1,524288;2,1,524288;4,1,524288;8,1,524288;8,8,524288")
for cfg in grid.split(";"):
    parts, rp, pool = (int(x) for x in cfg.split(","))
    t0 = time.time()
    a = run(parts, rp, pool)
    print(json.dumps({"parts": parts, "neg_group": rp, "pool_walks": pool, "auc": a, "seconds": round(time.time() - t0, 1)}), flush=True)

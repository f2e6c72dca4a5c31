"""Multi-GPU SGNS quality at scale, emulated exactly on one GPU (replicas are independent between
syncs): planted-partition graph with 1 M nodes / 20 M edges, main_link protocol (50 % of the edges
held out, walks R=5 L=40 p=0.25 q=4 on the rest, d=128, window 10), W replicas x combine rule x
sync interval (walks per replica between syncs). AUC on 1 M held-out edges vs 1 M non-edges."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch
from sklearn.metrics import roc_auc_score
from node2vec_by_ecc_b200 import DeviceGraph, SgnsTrainer, synth
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

def run(W, combine, interval):
    global GW
    tr = SgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=1)
    step_walks = interval * W
    for s0 in range(0, total, step_walks):
        s1 = min(total, s0 + step_walks)
        if W == 1:
            for g0 in range(s0, s1, B):
                nb = min(B, s1 - g0); walk(g0, nb)
                tr.train(walks[:nb], None, nb, L, total_examples=total, example_base=g0, sent_id_base=g0, sent_per_job=250, negative_sharing=int(os.environ.get("SHARED", "1")), grid_warps=GW)
            continue
        base0, base1 = tr.syn0.clone(), tr.syn1neg.clone()
        acc0, acc1 = torch.zeros_like(base0), torch.zeros_like(base1)
        sq0 = torch.zeros(base0.shape[0], 1, device=dev); sq1 = torch.zeros_like(sq0)
        for r in range(W):
            tr.syn0.copy_(base0); tr.syn1neg.copy_(base1)
            ra_, rb_ = s0 + r * interval, min(s1, s0 + (r + 1) * interval)
            for g0 in range(ra_, rb_, B):
                nb = min(B, rb_ - g0); walk(g0, nb)
                tr.train(walks[:nb], None, nb, L, total_examples=total, example_base=g0, sent_id_base=g0, sent_per_job=250, negative_sharing=int(os.environ.get("SHARED", "1")), grid_warps=GW)
            d0, d1 = tr.syn0 - base0, tr.syn1neg - base1
            if os.environ.get("BF16_DELTAS"):          # what a bf16 all-reduce of the deltas would carry
                d0, d1 = d0.bfloat16().float(), d1.bfloat16().float()
            acc0 += d0; acc1 += d1
            sq0 += (d0 * d0).sum(1, keepdim=True); sq1 += (d1 * d1).sum(1, keepdim=True)
            if os.environ.get("BF16_DELTAS"):
                acc0, acc1 = acc0.bfloat16().float(), acc1.bfloat16().float()
        if combine == "adapt":     # per-row redundancy c = |sum d|^2 / sum |d|^2 in [0, W]: identical deltas -> average, independent -> sum
            c0 = ((acc0 * acc0).sum(1, keepdim=True) / sq0.clamp_min(1e-30)).clamp_(min=1.0)
            c1 = ((acc1 * acc1).sum(1, keepdim=True) / sq1.clamp_min(1e-30)).clamp_(min=1.0)
            tr.syn0.copy_(base0 + acc0 / c0); tr.syn1neg.copy_(base1 + acc1 / c1)
        else:
            sc = 1.0 / W if combine == "avg" else 1.0
            tr.syn0.copy_(base0 + sc * acc0); tr.syn1neg.copy_(base1 + sc * acc1)
        del base0, base1, acc0, acc1
    return auc_of(tr)

print(json.dumps({"n": n, "edges": int(lo.numel()), "walks": total, "W": 1, "grid_warps": GW or "default", "auc": run(1, "avg", B)}), flush=True)
if os.environ.get("ONLY_W1"):
    sys.exit(0)
grid = os.environ.get("GRID")
cases = ([tuple(int(x) for x in c.split(":")) for c in grid.split(",")] if grid
         else [(W, iv) for W in (2, 8) for iv in (1 << 19, 1 << 16, 1 << 13)])
for W, interval in cases:
    for combine in (tuple(os.environ.get("COMBINE", "sum").split(",")) if grid else ("avg", "sum")):
        for _ in (0,):
            t0 = time.time()
            print(json.dumps({"W": W, "combine": combine, "walks_per_replica_per_sync": interval, "auc": run(W, combine, interval),
                              "s": round(time.time() - t0, 1)}), flush=True)

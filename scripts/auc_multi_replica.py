"""Does multi-GPU SGNS (replicated tables, periodic all-reduce) keep the link-prediction AUC?
Emulated exactly on ONE GPU: W replicas of (syn0, syn1neg) each train on their shard of every
batch, then are combined -- "avg": parameter averaging, "sum": base + sum of the
replicas' deltas. C2 graph, main_link protocol, vs the single-replica run and the CPU oracle."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_neg_samples, roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import DeviceGraph, SgnsTrainer, synth

n = 10000
lo, hi = synth.planted_edges(n, 333000, seed=42, device="cuda")
edges = np.stack([lo.cpu().numpy(), hi.cpu().numpy()], 1).astype(np.int64)
tr_e, te = split_edges(edges)
dg = DeviceGraph.from_coo(tr_e[:, 0], tr_e[:, 1], None, n, undirected=True)
t = dg.build_alias_tables(0.25, 4.0)
R, L = 5, 40
starts = torch.arange(n, dtype=torch.int32).repeat(R)
neg = build_neg_samples(n, edges, len(te), seed=1)
n_syncs = int(os.environ.get("N_SYNCS", "10"))

def auc_of(tr):
    emb = np.zeros((n, 128), np.float32); emb[tr.order.cpu().numpy()] = tr.syn0.cpu().numpy()
    return roc_auc_cosine(emb, te, neg)

res = {}
for seed in (1,):
    walks, lens = dg.walk_alias(t, starts, L, seed=seed)
    counts = torch.bincount(walks.reshape(-1).to(torch.int64), minlength=n)
    total = walks.shape[0]
    for W in (1, 2, 4, 8):
        for combine in (("avg", "sum", "tw") if W > 1 else ("avg",)):
            tr = SgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=seed)
            per_sync = total // n_syncs
            for s in range(n_syncs):
                a, b = s * per_sync, (total if s == n_syncs - 1 else (s + 1) * per_sync)
                base0, base1 = tr.syn0.clone(), tr.syn1neg.clone()
                acc0, acc1 = torch.zeros_like(base0), torch.zeros_like(base1)
                cnt0 = torch.zeros(base0.shape[0], 1, device=base0.device); cnt1 = torch.zeros_like(cnt0)
                shard = (b - a + W - 1) // W
                for r in range(W):                       # replica r trains on its contiguous shard
                    tr.syn0.copy_(base0); tr.syn1neg.copy_(base1)
                    ra, rb = a + r * shard, min(b, a + (r + 1) * shard)
                    if rb > ra:
                        tr.train(walks[ra:rb], None, rb - ra, L, total_examples=total, example_base=ra,
                                 sent_id_base=ra, sent_per_job=250, negative_sharing=1)
                    d0, d1 = tr.syn0 - base0, tr.syn1neg - base1
                    acc0 += d0; acc1 += d1
                    cnt0 += (d0 != 0).any(dim=1, keepdim=True).float(); cnt1 += (d1 != 0).any(dim=1, keepdim=True).float()
                if combine == "tw":      # average over the replicas that touched the row
                    tr.syn0.copy_(base0 + acc0 / cnt0.clamp_min(1.0)); tr.syn1neg.copy_(base1 + acc1 / cnt1.clamp_min(1.0))
                else:
                    scale = 1.0 / W if combine == "avg" else 1.0
                    tr.syn0.copy_(base0 + scale * acc0); tr.syn1neg.copy_(base1 + scale * acc1)
            res.setdefault(f"W{W}_{combine}", []).append(auc_of(tr))
    print("seed", seed, {k: round(v[-1], 4) for k, v in res.items()}, flush=True)
print(json.dumps({k: {"mean": float(np.mean(v)), "runs": v} for k, v in res.items()}))

"""Link-prediction AUC of the block-partitioned SGNS (BlockSgnsTrainer, all parts on one device =
exactly what n GPUs compute) on C2, main_link.main protocol (main_link.py:519-565), same graph /
split / walks as scripts/auc_c2.py. Sweeps parts x neg_group x pool size.
   GRID="parts,neg_group,pool[,warps];..." SEEDS=3 python scripts/auc_block.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import build_neg_samples, roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import BlockSgnsTrainer, DeviceGraph, SgnsTrainer, synth

n, R, L = 10000, 5, 40
lo, hi = synth.planted_edges(n, 333000, seed=42, device="cuda")
edges = np.stack([lo.cpu().numpy(), hi.cpu().numpy()], 1).astype(np.int64)
tr_e, te = split_edges(edges)
dg = DeviceGraph.from_coo(tr_e[:, 0], tr_e[:, 1], None, n, undirected=True)
t = dg.build_alias_tables(0.25, 4.0)
starts = torch.arange(n, dtype=torch.int32, device="cuda").repeat(R)
neg = build_neg_samples(n, edges, len(te), seed=1)
seeds = list(range(1, 1 + int(os.environ.get("SEEDS", "3"))))
grid = os.environ.get("GRID", "1,1,4096;2,1,4096;4,1,4096;8,1,4096;8,1,16384;8,1,1024;2,2,4096;4,4,4096;8,8,4096;8,4,4096;8,2,4096")
grid = [tuple(int(x) for x in g.split(",")) for g in grid.split(";")]      # optional 4th field: Hogwild width (warps)


def auc_of(s0, order):
    emb = np.zeros((n, 128), np.float32); emb[order.cpu().numpy()] = s0.cpu().numpy()
    return roc_auc_cosine(emb, te, neg)


out = {}
corpora = {}
for seed in seeds:
    walks, lens = dg.walk_alias(t, starts, L, seed=seed)
    corpora[seed] = walks
    counts = torch.bincount(walks[walks >= 0].to(torch.int64), minlength=n)
    ref = SgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=seed)
    ref.train(walks, None, walks.shape[0], L, total_examples=walks.shape[0], sent_per_job=250, negative_sharing=1)
    out.setdefault("sentence_major_shared", []).append(auc_of(ref.syn0, ref.order))
print("sentence_major_shared", np.mean(out["sentence_major_shared"]), flush=True)
for cfg in grid:
    parts, run, pool = cfg[:3]
    gw = cfg[3] if len(cfg) > 3 else None
    key = "parts=%d neg_group=%d pool=%d" % (parts, run, pool) + (" warps=%d" % gw if gw else "")
    for seed in seeds:
        walks = corpora[seed]
        counts = torch.bincount(walks[walks >= 0].to(torch.int64), minlength=n)
        trn = BlockSgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=seed, local_parts=parts, neg_group=run)
        total = walks.shape[0]
        for a in range(0, total, pool):
            b = min(total, a + pool)
            trn.train(walks[a:b], None, b - a, L, total_examples=total, example_base=a, sent_id_base=a, sent_per_job=10000 // L, grid_warps=gw)
        trn.check_overflow()
        s0, _ = trn.gather()
        out.setdefault(key, []).append(auc_of(s0, trn.order))
    print(key, round(float(np.mean(out[key])), 4), [round(x, 4) for x in out[key]], flush=True)
print(json.dumps({k: {"mean": float(np.mean(v)), "std": float(np.std(v)), "runs": v} for k, v in out.items()}))

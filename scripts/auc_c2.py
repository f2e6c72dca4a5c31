"""Link-prediction AUC parity on C2: N=10k, M=333k planted heavy-tailed graph,
main_link.main protocol (main_link.py:519-565): hold out 50 % of the edges (seed 123), walks
R=5 L=40 p=0.25 q=4 on the rest, SGNS d=128 window 10, cosine scores, ROC-AUC; 5 seeds each:
device (shared negatives, per-pair negatives) vs the CPU oracle (gensim restatement, all cores)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from helpers import build_neg_samples, roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import DeviceGraph, WalkCorpus, Word2Vec, synth

n = 10000
lo, hi = synth.planted_edges(n, 333000, seed=42, device="cuda")
edges = np.stack([lo.cpu().numpy(), hi.cpu().numpy()], 1).astype(np.int64)
tr, te = split_edges(edges)
dg = DeviceGraph.from_coo(tr[:, 0], tr[:, 1], None, n, undirected=True)
t = dg.build_alias_tables(0.25, 4.0)
starts = torch.arange(n, dtype=torch.int32).repeat(5)
neg = build_neg_samples(n, edges, len(te), seed=1)
res = {"device_shared": [], "device_per_pair": [], "device_reject_walks_shared": [], "oracle": []}
for seed in (1, 2, 3, 4, 5):
    walks, lens = dg.walk_alias(t, starts, 40, seed=seed)
    corpus = WalkCorpus(walks, lens, None)
    walks_np = walks.cpu().numpy()
    for key, shared in (("device_shared", 1), ("device_per_pair", 0)):
        m = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, seed=seed, shared_negatives=shared)
        emb = np.zeros((n, 128), np.float32); emb[np.asarray([int(w) for w in m.wv.index2word])] = m.wv.syn0
        res[key].append(roc_auc_cosine(emb, te, neg))
    wr, lr = dg.walk_reject(0.25, 4.0, starts, 40, seed=seed)
    m = Word2Vec(WalkCorpus(wr, lr, None), size=128, window=10, min_count=0, sg=1, iter=1, seed=seed)
    emb = np.zeros((n, 128), np.float32); emb[np.asarray([int(w) for w in m.wv.index2word])] = m.wv.syn0
    res["device_reject_walks_shared"].append(roc_auc_cosine(emb, te, neg))
    voc = oracle.sgns_vocab(walks_np, n)
    tok = voc.id2index[np.maximum(walks_np, 0)].astype(np.int32); tok[walks_np < 0] = -1
    off = np.arange(walks_np.shape[0] + 1, dtype=np.int64) * 40
    s0, _, pairs = oracle.sgns_train(tok, off, voc, dim=128, window=10, negative=5, workers=os.cpu_count(), rng_mode=0, seed=seed)
    emb = np.zeros((n, 128), np.float32); emb[voc.index2id] = s0
    res["oracle"].append(roc_auc_cosine(emb, te, neg))
    print("seed", seed, {k: round(v[-1], 4) for k, v in res.items()}, flush=True)
out = {k: {"mean": float(np.mean(v)), "std": float(np.std(v)), "runs": v} for k, v in res.items()}
for k in out:
    out[k]["delta_vs_oracle"] = out[k]["mean"] - out["oracle"]["mean"]
print(json.dumps(out))

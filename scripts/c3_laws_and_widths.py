"""C3-shaped user-item graph (tests/test_gpu_auc.py): link-prediction AUC of both device laws at full and narrow Hogwild\nwidth against the CPU oracle under both laws (16 workers and 1 worker) on the same walks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from helpers import roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import DeviceGraph, WalkCorpus, Word2Vec, synth
R, L = 5, 40
nu, ni, m = 10_000, 40_000, 1_000_000
u, it, w, n = synth.bipartite_edges(nu, ni, m, seed=7, device="cuda")
edges = np.stack([u.cpu().numpy(), it.cpu().numpy()], 1).astype(np.int64)
wts = w.cpu().numpy()
tr_i, te_i = split_edges(np.arange(len(edges)))
tr, te = edges[tr_i], edges[te_i[:100_000]]
dg = DeviceGraph.from_coo(tr[:, 0], tr[:, 1], wts[tr_i], n, undirected=True)
rng = np.random.RandomState(5)
true = set(map(tuple, edges.tolist()))
neg = []
while len(neg) < len(te):
    a, b = int(rng.randint(0, nu)), int(nu + rng.randint(0, ni))
    if (a, b) not in true:
        neg.append((a, b))
neg = np.asarray(neg, dtype=np.int64)
starts = torch.arange(n, dtype=torch.int32).repeat(R)
for seed in (1, 2):
    wk, ln = dg.walk_reject(0.25, 4.0, starts, L, seed=seed)
    res = {}
    for name, kw in (("dev shared", dict(shared_negatives=1)), ("dev per-pair", dict(shared_negatives=0)),
                     ("dev shared narrow", dict(shared_negatives=1, hogwild_warps=64)), ("dev per-pair narrow", dict(shared_negatives=0, hogwild_warps=64))):
        mdl = Word2Vec(WalkCorpus(wk, ln, None), size=128, window=10, min_count=0, sg=1, iter=1, seed=seed, **kw)
        emb = np.zeros((n, 128), np.float32); emb[np.asarray([int(x) for x in mdl.wv.index2word])] = mdl.wv.syn0
        res[name] = round(roc_auc_cosine(emb, te, neg), 4)
    wn = wk.cpu().numpy()
    voc = oracle.sgns_vocab(wn, n)
    tok = voc.id2index[np.maximum(wn, 0)].astype(np.int32); tok[wn < 0] = -1
    off = np.arange(wn.shape[0] + 1, dtype=np.int64) * L
    for name, mode, wk_ in (("oracle per-pair 16w", 0, os.cpu_count()), ("oracle shared 16w", 2, os.cpu_count()), ("oracle per-pair 1w", 0, 1)):
        s0, _, _ = oracle.sgns_train(tok, off, voc, dim=128, window=10, negative=5, workers=wk_, rng_mode=mode, seed=seed)
        emb = np.zeros((n, 128), np.float32); emb[voc.index2id] = s0
        res[name] = round(roc_auc_cosine(emb, te, neg), 4)
    print(seed, res, flush=True)

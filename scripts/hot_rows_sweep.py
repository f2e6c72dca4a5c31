"""Link-prediction AUC of the shared-negative kernel on the C3-shaped user-item graph of tests/test_gpu_auc.py against\nthe number of hot (uncarried) negative rows and the Hogwild width; the CPU oracle under the same law reads 0.2342."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
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
wk, ln = dg.walk_reject(0.25, 4.0, starts, L, seed=1)
import time
for hot in ("default", "0", "100", "1000", "5000", "60000"):
    for width in (None, 740, 64):
        if hot == "default":
            os.environ.pop("N2V_SGNS_HOT_ROWS", None)
        else:
            os.environ["N2V_SGNS_HOT_ROWS"] = hot
        torch.cuda.synchronize(); t0 = time.time()
        mdl = Word2Vec(WalkCorpus(wk, ln, None), size=128, window=10, min_count=0, sg=1, iter=1, seed=1, shared_negatives=1, hogwild_warps=width)
        torch.cuda.synchronize(); dt = time.time() - t0
        emb = np.zeros((n, 128), np.float32); emb[np.asarray([int(x) for x in mdl.wv.index2word])] = mdl.wv.syn0
        T = mdl.trainer
        print("hot", hot, "->", T.hot_rows(width or T.default_hogwild_warps(True)), "width", width or T.default_hogwild_warps(True),
              "auc", round(roc_auc_cosine(emb, te, neg), 4), "s", round(dt, 2), flush=True)

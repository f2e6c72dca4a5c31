"""Hogwild-width / update-mode sweep of link-prediction AUC (main_link.py:519-565 protocol) on a
planted-partition graph; device trainer vs the CPU oracle on the same walk corpus."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from helpers import build_neg_samples, chung_lu_graph, roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import DeviceGraph, WalkCorpus, Word2Vec

n, m, comm, maxdeg = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
widths = [int(x) for x in sys.argv[5].split(",")]
edges = chung_lu_graph(n, m, seed=42, max_deg=maxdeg, communities=comm)
tr, te = split_edges(edges)
dg = DeviceGraph.from_coo(tr[:, 0], tr[:, 1], None, n, undirected=True)
t = dg.build_alias_tables(0.25, 4.0)
starts = torch.arange(n, dtype=torch.int32).repeat(5)
walks, lens = dg.walk_alias(t, starts, 40, seed=9)
neg = build_neg_samples(n, edges, len(te), seed=1)
corpus = WalkCorpus(walks, lens, None)
walks_np = walks.cpu().numpy()

def auc_of(m_):
    emb = np.zeros((n, 128), dtype=np.float32)
    emb[np.asarray([int(w) for w in m_.wv.index2word])] = m_.wv.syn0
    return roc_auc_cosine(emb, te, neg)

shared = int(os.environ.get("SWEEP_SHARED", "0"))
for atomic in (0, 1):
    for wdt in widths:
        aucs = []
        for seed in (1, 2, 3):
            torch.cuda.synchronize(); t0 = time.time()
            mm = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, seed=seed,
                          hogwild_warps=wdt, atomic_updates=atomic, shared_negatives=shared)
            torch.cuda.synchronize(); dt = time.time() - t0
            aucs.append(auc_of(mm))
        print(f"shared={shared} atomic={atomic} width={wdt} auc={np.round(aucs, 4)} mean={np.mean(aucs):.4f} pairs={mm.pairs_trained} t={dt:.2f}s", flush=True)

voc = oracle.sgns_vocab(walks_np, n)
tok = voc.id2index[np.maximum(walks_np, 0)].astype(np.int32); tok[walks_np < 0] = -1
off = np.arange(walks_np.shape[0] + 1, dtype=np.int64) * 40
aucs = []
for seed in (1, 2, 3):
    s0, _, pairs = oracle.sgns_train(tok, off, voc, dim=128, window=10, negative=5, workers=os.cpu_count(), rng_mode=0, seed=seed)
    emb = np.zeros((n, 128), np.float32); emb[voc.index2id] = s0
    aucs.append(roc_auc_cosine(emb, te, neg))
print(f"oracle workers={os.cpu_count()} auc={np.round(aucs, 4)} mean={np.mean(aucs):.4f} pairs={pairs}")

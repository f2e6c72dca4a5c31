"""Every kernel of libn2v_b200.so once on karate-sized inputs, all modes and odd sizes (written as the
driver for `compute-sanitizer --tool memcheck`; the sanitizer is closed on this GPU pool, so the
out-of-bounds-write checks live in tests/test_gpu_guards.py instead)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from node2vec_by_ecc_b200 import DeviceGraph, WalkCorpus, Word2Vec, KeyedVectors
from node2vec_by_ecc_b200.data import KARATE_EDGES

e = np.asarray(KARATE_EDGES, dtype=np.int64) - 1
rng = np.random.RandomState(0)
for weighted in (False, True):
    w = rng.rand(len(e)) + 0.1 if weighted else None
    dg = DeviceGraph.from_coo(e[:, 0], e[:, 1], w, 34, undirected=True)
    t = dg.build_alias_tables(0.25, 4.0, keep_raw=True)
    tp = dg.build_alias_tables(0.25, 4.0, popwalk=True, pop_edges=True)
    starts = torch.arange(34, dtype=torch.int32).repeat(3)
    for L in (1, 2, 7, 8, 9, 33):
        for packed in (True, False):
            dg.walk_alias(t, starts, L, 1, 5, packed=packed)
        cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
        for indexed in (True, False):
            dg.walk_reject(0.25, 4.0, starts, L, 2, 7, counters=cnt, indexed=indexed)
    walks, lens = dg.walk_alias(t, starts.repeat(4), 40, 3)
    c = WalkCorpus(walks, lens, np.arange(34))
    c.format_walks()
    for shared, neg, dim in ((1, 5, 128), (1, 5, 64), (0, 5, 128), (0, 3, 100), (0, 5, 256)):
        for atomic in (1, 0):
            for width in (1, 8):
                m = Word2Vec(c, size=dim, window=10, min_count=0, sg=1, iter=1, negative=neg, hogwild_warps=width,
                             atomic_updates=atomic, shared_negatives=shared)
    m.wv.similarity_pairs([("1", "2"), ("3", "nope")])
# directed graph with sinks
a, b = rng.randint(0, 40, 150), rng.randint(0, 50, 150)
dg = DeviceGraph.from_coo(a, b, rng.rand(150), 50, undirected=False)
t = dg.build_alias_tables(0.5, 2.0)
st = torch.arange(50, dtype=torch.int32)
dg.walk_alias(t, st, 20, 1); dg.walk_alias(t, st, 20, 1, packed=False)
dg.walk_reject(0.5, 2.0, st, 20, 1); dg.walk_reject(0.5, 2.0, st, 20, 1, indexed=False)
torch.cuda.synchronize()
print("exercise_all_kernels: done")

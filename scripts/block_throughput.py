"""Per-GPU throughput of the block-partitioned SGNS on the bench workload (C4), one device: a pool
of POOL walks is expanded and trained with all `parts` on this device. With parts = n this device
does the pair expansion of n GPUs' centre parts one after another and trains every bucket, i.e.
n x the per-GPU work of an n-GPU step over the same pool: per-GPU step time ~ total / n.
   PARTS=1,8 NEG_GROUPS=1 POOL=524288 python scripts/block_throughput.py"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
from node2vec_by_ecc_b200 import BlockSgnsTrainer, DeviceGraph, synth
from node2vec_by_ecc_b200._lib import check, lib, ptr, stream

scale, edges = int(os.environ.get("SCALE", "22")), float(os.environ.get("EDGES", "100e6"))
POOL, L = int(os.environ.get("POOL", str(1 << 19))), 80
dev = torch.device("cuda", 0)
lo, hi, n = synth.rmat_edges(scale, int(edges), seed=1, device=dev)
dg = DeviceGraph.from_coo(lo, hi, None, n, undirected=True)
del lo, hi
walks = torch.empty((POOL, L), dtype=torch.int32, device=dev)
lens = torch.empty(POOL, dtype=torch.int32, device=dev)
counts = torch.zeros(n, dtype=torch.int64, device=dev)
for s in range(0, n, POOL):                       # vocabulary: one walk per node
    e = min(n, s + POOL)
    st = torch.arange(s, s + POOL, dtype=torch.int32, device=dev) % n
    dg.walk_reject(0.25, 4.0, st, L, 1, (1 << 40) + s, out=(walks, lens))
    check(lib().n2v_vocab_count(ptr(walks), C.c_int64((e - s) * L), C.c_int32(n), ptr(counts), stream()))
ev = lambda: torch.cuda.Event(enable_timing=True)
res = []
for parts in [int(x) for x in os.environ.get("PARTS", "1,8").split(",")]:
    for run in [int(x) for x in os.environ.get("NEG_GROUPS", "1").split(",")]:
        trn = BlockSgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=1, local_parts=parts, neg_group=run)
        P = trn._params(0, None, total_examples=10 * n, sent_per_job=125)
        for it in range(3):                       # 2 warm-up pools, 1 timed
            st = (torch.arange(POOL, dtype=torch.int64, device=dev) + it * POOL) % n
            dg.walk_reject(0.25, 4.0, st.to(torch.int32), L, 1, it * POOL, out=(walks, lens))
            p0 = int(trn.pairs[0]); c0 = int(trn.pairs[1])
            t_pairs = t_train = 0.0
            for k in range(parts):                # one centre part after the other, as GPU k would
                a0, a1, a2 = ev(), ev(), ev()
                a0.record()
                _, bounds = trn.make_groups(walks, None, POOL, L, it * POOL, P, k)
                a1.record()
                bev = []
                for b in range(parts):
                    x0, x1 = ev(), ev()
                    x0.record()
                    trn.train_bucket(k, b, trn.parts0[b], P, POOL, it * POOL, bounds)
                    x1.record()
                    bev.append((x0, x1, bounds[b + 1] - bounds[b]))
                a2.record(); torch.cuda.synchronize()
                t_pairs += a0.elapsed_time(a1); t_train += a1.elapsed_time(a2)
                if it == 2 and os.environ.get("BUCKET_TIMES"):     # ms and pairs of every bucket (k, b) of the timed pool
                    print(json.dumps({"centre_part": k, "bucket_ms": [round(x.elapsed_time(y), 3) for x, y, _ in bev],
                                      "bucket_words": [int(n_) for _, _, n_ in bev]}), flush=True)
            npairs = int(trn.pairs[0]) - p0
        trn.check_overflow()
        r = {"parts": parts, "neg_group": run, "pool_walks": POOL, "pairs": npairs, "make_groups_ms": t_pairs, "rows_per_pair": 1.0 + (int(trn.pairs[1]) - c0) / max(npairs, 1), "train_ms": t_train,
             "train_pairs_per_s": npairs / (t_train / 1e3), "pairs_per_s_incl_expansion": npairs / ((t_pairs + t_train) / 1e3)}
        print(json.dumps(r), flush=True)
        res.append(r)
        del trn
        torch.cuda.empty_cache()

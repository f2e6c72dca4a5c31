"""Where the expansion of a pool goes: count kernel + scan vs fill kernel, per centre part, on the bench
workload (C4). PARTS=2,8 POOL=524288 python scripts/expansion_split.py"""
import os, sys, json
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
for s in range(0, n, POOL):
    e = min(n, s + POOL)
    st = torch.arange(s, s + POOL, dtype=torch.int32, device=dev) % n
    dg.walk_reject(0.25, 4.0, st, L, 1, (1 << 40) + s, out=(walks, lens))
    check(lib().n2v_vocab_count(ptr(walks), C.c_int64((e - s) * L), C.c_int32(n), ptr(counts), stream()))
ev = lambda: torch.cuda.Event(enable_timing=True)
for parts in [int(x) for x in os.environ.get("PARTS", "2,8").split(",")]:
    trn = BlockSgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=1, local_parts=parts)
    P = trn._params(0, None, total_examples=10 * n, sent_per_job=125)
    for k in range(min(parts, 2)):
        words, bounds = trn.make_groups(walks, None, POOL, L, 0, P, k)          # sizes the buffers
        b = trn._buf[k]
        keep = ptr(trn.keep_thr)
        head = (ptr(walks), None, C.c_int64(POOL), C.c_int32(L), C.c_int64(0), ptr(trn.vocab_of_id), keep, C.byref(P),
                C.c_int32(k), C.c_int32(parts))
        t_c = t_f = 0.0
        for it in range(5):
            a0, a1, a2 = ev(), ev(), ev()
            a0.record()
            check(lib().n2v_sgns_groups_count(*head, ptr(b["offsets"]), ptr(b["ws"]), C.c_size_t(b["ws"].numel()), stream()))
            a1.record()
            check(lib().n2v_sgns_groups_fill(*head, ptr(trn.cum_table), ptr(trn.bucket_lo), C.c_int32(1), ptr(b["offsets"]),
                                             ptr(b["words"]), C.c_int64(b["cap"]), ptr(b["overflow"]), stream()))
            a2.record(); torch.cuda.synchronize()
            if it >= 2:
                t_c += a0.elapsed_time(a1) / 3; t_f += a1.elapsed_time(a2) / 3
        print(json.dumps({"parts": parts, "centre_part": k, "pool_walks": POOL, "stream_words": bounds[-1],
                          "count_scan_ms": round(t_c, 3), "fill_ms": round(t_f, 3)}), flush=True)
    del trn
    torch.cuda.empty_cache()

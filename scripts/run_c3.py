"""C3 (BASELINE.json configs[2]): synthetic bipartite user-item graph, 1 M nodes / 20 M weighted
edges, walked (R=10, L=80) and embedded (d=128, window 10) end to end on one GPU, p=q=1 and
p=0.25,q=4. Prints one JSON line per (p, q)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from node2vec_by_ecc_b200 import DeviceGraph, SgnsTrainer, synth
from node2vec_by_ecc_b200._lib import check, lib, ptr, stream
import ctypes as C

dev = torch.device("cuda", 0)
scale = float(os.environ.get("C3_SCALE", "1.0"))
t0 = time.time()
u, it, w, n = synth.bipartite_edges(int(200_000 * scale), int(800_000 * scale), int(20_000_000 * scale), device=dev)
torch.cuda.synchronize(); t_gen = time.time() - t0
t0 = time.time()
dg = DeviceGraph.from_coo(u, it, w, n, undirected=True)
torch.cuda.synchronize(); t_csr = time.time() - t0
deg = (dg.row_ptr[1:] - dg.row_ptr[:-1])
sds = dg.sum_deg_sq()
R, L, B = 10, 80, 1 << 20
for p, q in ((1.0, 1.0), (0.25, 4.0)):
    rec = {"config": "C3 bipartite", "n": n, "edges": int(u.numel()), "nnz": dg.nnz, "max_deg": int(deg.max()),
           "sum_deg_sq": sds, "edge_table_GB": 8 * sds / 1e9, "p": p, "q": q, "gen_s": t_gen, "csr_s": t_csr}
    torch.cuda.synchronize(); t0 = time.time()
    free = torch.cuda.mem_get_info()[0]
    use_alias = 8 * sds * 2.5 < 0.6 * free
    tables = dg.build_alias_tables(p, q) if use_alias else dg.build_node_tables()
    if not use_alias:
        dg.reject_index()
    torch.cuda.synchronize(); rec["mode"] = "alias" if use_alias else "reject(indexed, weighted)"
    rec["preprocess_s"] = time.time() - t0
    walks = torch.empty((B, L), dtype=torch.int32, device=dev); lens = torch.empty(B, dtype=torch.int32, device=dev)
    counts = torch.zeros(n, dtype=torch.int64, device=dev)
    total = R * n
    def walk(g0):
        st = ((g0 + torch.arange(B, device=dev)) % n).to(torch.int32)
        if use_alias:
            dg.walk_alias(tables, st, L, 1, g0, out=(walks, lens))
        else:
            dg.walk_reject(p, q, st, L, 1, g0, node_tables=tables, out=(walks, lens))
    # pass 1: vocabulary counts over the whole corpus (scan_vocab)
    torch.cuda.synchronize(); t0 = time.time()
    for g0 in range(0, total, B):
        walk(g0)
        nb = min(B, total - g0)
        check(lib().n2v_vocab_count(ptr(walks), C.c_int64(nb * L), C.c_int32(n), ptr(counts), stream()))
    torch.cuda.synchronize(); rec["count_pass_s"] = time.time() - t0
    tr = SgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=1)
    # pass 2: walk + train
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    w_ms = s_ms = 0.0
    steps = 0
    torch.cuda.synchronize(); t0 = time.time()
    for g0 in range(0, total, B):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        nb = min(B, total - g0)
        a.record(); walk(g0); b.record()
        tr.train(walks, None, nb, L, total_examples=total, example_base=g0, sent_id_base=g0, sent_per_job=125,
                 negative_sharing=1)
        c.record(); torch.cuda.synchronize()
        w_ms += a.elapsed_time(b); s_ms += b.elapsed_time(c)
        steps += int((lens[:nb].to(torch.int64) - 1).clamp_(min=0).sum().item())
    rec["train_pass_s"] = time.time() - t0
    pairs = int(tr.pairs[0].item())
    rec.update(walk_steps=steps, walk_steps_per_s=steps / (w_ms / 1e3), pairs=pairs, sgns_pairs_per_s=pairs / (s_ms / 1e3),
               finite=bool(torch.isfinite(tr.syn0).all().item()))
    print(json.dumps(rec), flush=True)
    del tables, tr
    torch.cuda.empty_cache()

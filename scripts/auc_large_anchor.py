"""At-size parity anchor: the 1 M-node / 20 M-edge planted graph of
scripts/auc_block_large.py under the main_link protocol (50 % of the edges held out, walks R=5 L=40 p=0.25
q=4 on the rest, d=128, window 10, one epoch), ONE corpus (5 M rejection walks, seed 1), link-prediction AUC
on 1 M held-out edges vs 1 M non-edges for
  * the CPU oracle, gensim's per-pair law, all host cores (oracle/sgns_oracle.c -- parity unpinned),
  * the sentence-major kernels (shared-negative v3, per-pair v2),
  * the block-partitioned trainer at PARTS parts (one device == N ranks), pools of 2^19 walks.
One JSON line per run.   PARTS=1,2,4,8 python scripts/auc_large_anchor.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch
from sklearn.metrics import roc_auc_score
import oracle
from node2vec_by_ecc_b200 import BlockSgnsTrainer, DeviceGraph, SgnsTrainer, synth
from node2vec_by_ecc_b200._lib import check, lib, ptr, stream

dev = torch.device("cuda", 0)
n, m = int(float(os.environ.get("N", 1e6))), int(float(os.environ.get("M", 2e7)))
lo, hi = synth.planted_edges(n, m, seed=42, max_deg=4000, communities=max(39, n // 1000), device=dev)
g = torch.Generator(device=dev); g.manual_seed(123)
perm = torch.randperm(lo.numel(), device=dev, generator=g)
half = lo.numel() // 2
tr_i, te_i = perm[:half], perm[half:half + 1_000_000]
dg = DeviceGraph.from_coo(lo[tr_i], hi[tr_i], None, n, undirected=True)
pos_a, pos_b = lo[te_i].clone().long(), hi[te_i].clone().long()
keys = torch.sort(lo.to(torch.int64) * n + hi.to(torch.int64)).values
ra = torch.randint(0, n, (1_300_000,), device=dev, generator=g); rb = torch.randint(0, n, (1_300_000,), device=dev, generator=g)
a_, b_ = torch.minimum(ra, rb), torch.maximum(ra, rb)
k_ = a_ * n + b_
ok = (a_ != b_) & (keys[torch.searchsorted(keys, k_).clamp_(max=keys.numel() - 1)] != k_)
neg_a, neg_b = a_[ok][:1_000_000], b_[ok][:1_000_000]
R, L, B = 5, 40, 1 << 19
total = R * n
walks_all = torch.empty((total, L), dtype=torch.int32, device=dev)
lens = torch.empty(B, dtype=torch.int32, device=dev)
for g0 in range(0, total, B):
    nb = min(B, total - g0)
    st = ((g0 + torch.arange(nb, device=dev)) % n).to(torch.int32)
    dg.walk_reject(0.25, 4.0, st, L, 1, g0, out=(walks_all[g0:g0 + nb], lens[:nb]))
counts = torch.bincount(walks_all[walks_all >= 0].to(torch.int64), minlength=n)
y = np.concatenate([np.ones(pos_a.numel()), np.zeros(neg_a.numel())])


def auc_rows(syn0_dev, vocab_of_id):
    out = torch.empty(2_000_000, dtype=torch.float32, device=dev)
    ia = torch.cat([vocab_of_id[pos_a], vocab_of_id[neg_a]]).to(torch.int32).contiguous()
    ib = torch.cat([vocab_of_id[pos_b], vocab_of_id[neg_b]]).to(torch.int32).contiguous()
    check(lib().n2v_cosine_pairs(ptr(syn0_dev), C.c_int32(128), ptr(ia), ptr(ib), C.c_int64(ia.numel()), ptr(out), stream()))
    return float(roc_auc_score(y, out.cpu().numpy()))


def emit(**kw):
    print(json.dumps(kw), flush=True)


spj = 10000 // L
for shared in (1, 0):
    t0 = time.time()
    tr = SgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=1)
    for g0 in range(0, total, B):
        nb = min(B, total - g0)
        tr.train(walks_all[g0:g0 + nb], None, nb, L, total_examples=total, example_base=g0, sent_id_base=g0, sent_per_job=spj,
                 negative_sharing=shared)
    emit(run="sentence-major kernel", negatives="one set per centre occurrence" if shared else "fresh set per pair (gensim's law)",
         auc=auc_rows(tr.syn0, tr.vocab_of_id), pairs=int(tr.pairs[0]), seconds=round(time.time() - t0, 1))
    del tr
for parts in [int(x) for x in os.environ.get("PARTS", "1,2,4,8").split(",") if x]:
    for G in [int(x) for x in os.environ.get("NEG_GROUPS", "1").split(",")]:
        t0 = time.time()
        tr = BlockSgnsTrainer(counts, dim=128, window=10, negative=5, sample=1e-3, seed=1, local_parts=parts, neg_group=G)
        for g0 in range(0, total, B):
            nb = min(B, total - g0)
            tr.train(walks_all[g0:g0 + nb], None, nb, L, total_examples=total, example_base=g0, sent_id_base=g0, sent_per_job=spj)
        tr.check_overflow()
        s0, _ = tr.gather()
        emit(run="block-partitioned trainer", parts=parts, neg_group=G, pool_walks=B, auc=auc_rows(s0, tr.vocab_of_id),
             pairs=int(tr.pairs[0]), seconds=round(time.time() - t0, 1))
        del tr, s0
if os.environ.get("ORACLE", "1") == "1":
    t0 = time.time()
    wn = walks_all.cpu().numpy()
    voc = oracle.vocab_from_counts(counts.cpu().numpy())
    tok = np.where(wn >= 0, voc.id2index[np.maximum(wn, 0)], -1).astype(np.int32)
    off = np.arange(wn.shape[0] + 1, dtype=np.int64) * L
    s0, _, pairs = oracle.sgns_train(tok, off, voc, dim=128, window=10, negative=5, workers=os.cpu_count(), rng_mode=0, seed=1)
    id2 = torch.as_tensor(voc.id2index.astype(np.int64), device=dev)
    emit(run="CPU oracle (gensim per-pair law, parity unpinned)", workers=os.cpu_count(), auc=auc_rows(torch.as_tensor(s0).to(dev), id2),
         pairs=int(pairs), seconds=round(time.time() - t0, 1))

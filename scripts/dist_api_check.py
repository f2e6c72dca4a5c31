"""The reference-facing calls on N GPUs (run under torchrun, one rank per GPU): every rank builds the
graph, node2vec.Graph(..., distributed=True).simulate_walks returns the rank's share of the walks
(global walk ids), learn_embeddings' Word2Vec([map(str, walk) for walk in walks], ...) recognises the
device corpus and trains block-partitioned over the NCCL ring; every rank ends with the full table.
C2 protocol (main_link.py:519-565), link-prediction AUC on rank 0.
   torchrun --nproc-per-node N scripts/dist_api_check.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import build_neg_samples, roc_auc_cosine, split_edges
from node2vec_by_ecc_b200 import DeviceGraph, Graph, Word2Vec, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
os.environ.setdefault("N2V_POOL_WALKS", "4096")
n, R, L = 10000, 5, 40
lo, hi = synth.planted_edges(n, 333000, seed=42, device=dev)
edges = np.stack([lo.cpu().numpy(), hi.cpu().numpy()], 1).astype(np.int64)
tr_e, te = split_edges(edges)
neg = build_neg_samples(n, edges, len(te), seed=1)
dg = DeviceGraph.from_coo(tr_e[:, 0], tr_e[:, 1], None, n, undirected=True)
aucs, shares = [], []
for seed in (1, 2, 3):
    G = Graph(dg, False, 0.25, 4.0, seed=seed, distributed=True)
    G.preprocess_transition_probs()
    walks = G.simulate_walks(R, L)
    shares.append(len(walks))
    model = Word2Vec([map(str, walk) for walk in walks], size=128, window=10, min_count=0, sg=1, workers=8, iter=1, seed=seed)
    emb = np.zeros((n, 128), np.float32)
    emb[np.asarray([int(w) for w in model.wv.index2word])] = model.wv.syn0
    aucs.append(roc_auc_cosine(emb, te, neg))
    pairs = model.pairs_trained
if rank == 0:
    print(json.dumps({"world": world, "walks_per_rank": shares, "corpus_count": model.corpus_count, "pairs": pairs,
                      "trainer": type(model.trainer).__name__, "auc_mean": float(np.mean(aucs)), "auc_runs": aucs}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

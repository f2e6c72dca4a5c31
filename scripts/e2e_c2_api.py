"""C2 through the reference-facing Python API, exactly as src/main.py:92-101 drives it: networkx
graph -> node2vec.Graph -> preprocess_transition_probs -> simulate_walks(10, 80) ->
learn_embeddings' Word2Vec(...). Wall-clock per phase (includes the Python glue the API implies)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "node2vec_by_ecc_b200", "dropin")); sys.path.insert(0, ROOT)
import networkx as nx, numpy as np, torch
import node2vec
from gensim.models import Word2Vec
from node2vec_by_ecc_b200 import synth

lo, hi = synth.planted_edges(10000, 333000, seed=42, device="cuda")
edges = np.stack([lo.cpu().numpy(), hi.cpu().numpy()], 1)
t = {}
t0 = time.time()
nx_G = nx.Graph()
nx_G.add_nodes_from(range(10000))
nx_G.add_edges_from(((int(a), int(b), {"weight": 1}) for a, b in edges))
t["networkx_graph_s"] = time.time() - t0
torch.cuda.synchronize(); t0 = time.time()
G = node2vec.Graph(nx_G, False, 0.25, 4.0)
G.preprocess_transition_probs()
torch.cuda.synchronize(); t["ingest_csr_and_alias_tables_s"] = time.time() - t0
t0 = time.time()
walks = G.simulate_walks(10, 80)
torch.cuda.synchronize(); t["simulate_walks_s"] = time.time() - t0
t0 = time.time()
model_dev = Word2Vec(walks, size=128, window=10, min_count=0, sg=1, workers=8, iter=1)      # corpus stays on the device
torch.cuda.synchronize(); t["word2vec_device_corpus_s"] = time.time() - t0
t0 = time.time()
sents = [map(str, walk) for walk in walks]                                                   # main.py:86 verbatim
t["walks_to_python_lists_s"] = time.time() - t0
t0 = time.time()
model = Word2Vec(sents, size=128, window=10, min_count=0, sg=1, workers=8, iter=1)
torch.cuda.synchronize(); t["word2vec_string_corpus_s"] = time.time() - t0
t.update(walks=len(walks), steps=walks.num_steps(), pairs=model.pairs_trained, vocab=len(model.wv.vocab),
         edge_table_entries=G._dg.sum_deg_sq())
print(json.dumps(t))

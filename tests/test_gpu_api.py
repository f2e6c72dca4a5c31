"""GPU: the reference-facing Python API end to end -- what src/main.py:92-101 does, with the
drop-in modules first on sys.path."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def karate_nx():
    import networkx as nx
    from node2vec_by_ecc_b200.data import KARATE_EDGES
    G = nx.DiGraph()
    for a, b in KARATE_EDGES:
        G.add_edge(a, b)
        G[a][b]["weight"] = 1
    return G.to_undirected()


def test_main_py_pipeline_with_dropins():
    sys.path.insert(0, os.path.join(ROOT, "node2vec_by_ecc_b200", "dropin"))
    try:
        import node2vec
        from gensim.models import Word2Vec
        nx_G = karate_nx()
        G = node2vec.Graph(nx_G, False, 1, 1)
        G.preprocess_transition_probs()
        walks = G.simulate_walks(10, 80)
        assert len(walks) == 340 and all(len(w) == 80 for w in walks)
        first = list(nx_G.nodes())
        assert [w[0] for w in walks[:34]] == first and [w[0] for w in walks[34:68]] == first
        for w in walks[:50]:
            assert all(nx_G.has_edge(a, b) for a, b in zip(w[:-1], w[1:]))
        # learn_embeddings (main.py:82-90) verbatim
        sents = [map(str, walk) for walk in walks]
        model = Word2Vec(sents, size=128, window=10, min_count=0, sg=1, workers=8, iter=1)
        assert model.wv.syn0.shape == (34, 128)
        assert set(model.wv.vocab.keys()) == {str(v) for v in nx_G.nodes()}
        # reference dict views of the tables
        J, q = G.alias_nodes[1]
        assert len(J) == nx_G.degree(1) and np.allclose(q, 1.0)
        J, q = G.alias_edges[(1, 2)]
        assert len(J) == nx_G.degree(2)
        J2, q2 = G.get_alias_edge(1, 2)
        assert (J == J2).all() and (q == q2).all()
    finally:
        sys.path.pop(0)


def test_graph_matches_golden_through_public_api():
    """Graph(...).simulate_walks == the walks the reference produced (tests/golden) when seeded alike"""
    from helpers import load_case
    from node2vec_by_ecc_b200 import Graph
    z, _ = load_case("karate_p025_q4")
    G = Graph(karate_nx(), False, 0.25, 4.0, seed=int(z["seed"]), mode="alias")
    G.preprocess_transition_probs()
    walks = G.simulate_walks(int(z["R"]), int(z["L"]))
    labels = z["labels"]
    want = [[int(labels[t]) for t in row[:l]] for row, l in zip(z["walks"], z["lens"])]
    assert [list(w) for w in walks] == want
    # on-the-fly entry point: same tokens (node2vec.py:97-111), fresh walk ids on a fresh Graph
    G2 = Graph(karate_nx(), False, 0.25, 4.0, seed=int(z["seed"]), mode="alias")
    walks2 = G2.simulate_walks_on_the_fly(int(z["R"]), int(z["L"]))
    assert [list(w) for w in walks2] == want
    # rejection mode through the same API: valid walks of the right shape
    G3 = Graph(karate_nx(), False, 0.25, 4.0, mode="reject")
    G3.preprocess_transition_probs()
    w3 = G3.simulate_walks(2, 30)
    assert len(w3) == 68 and all(len(w) == 30 for w in w3)


def test_graph_survives_pickle():
    import pickle
    from node2vec_by_ecc_b200 import Graph
    G = Graph(karate_nx(), False, 1, 1)
    G.preprocess_transition_probs()
    G2 = pickle.loads(pickle.dumps(G))
    assert len(G2.simulate_walks_on_the_fly(1, 10)) == 34

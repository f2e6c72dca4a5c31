"""GPU: the reference-facing Python API end to end -- what src/main.py:92-101 does, with the
drop-in modules first on sys.path."""
import os
import sys

import numpy as np
import torch
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
        # the one-shot map objects were recognised as rows of the device corpus (nothing was stringified):
        # same tables as training on the corpus object, and as the generic string path
        assert model._corpus[0] is walks.walks
        direct = Word2Vec(walks, size=128, window=10, min_count=0, sg=1, workers=8, iter=1, hogwild_warps=1)
        seq = Word2Vec([map(str, walk) for walk in walks], size=128, window=10, min_count=0, sg=1, workers=8, iter=1,
                       hogwild_warps=1)
        generic = Word2Vec([[str(t) for t in walk] for walk in walks], size=128, window=10, min_count=0, sg=1, workers=8,
                           iter=1, hogwild_warps=1)
        assert np.array_equal(seq.wv.syn0, direct.wv.syn0) and seq.wv.index2word == direct.wv.index2word
        assert seq.pairs_trained == generic.pairs_trained
        for w in ("1", "34", "17"):
            assert np.allclose(seq.wv[w], generic.wv[w], atol=1e-6)
        # a shuffled / partial list of rows still trains on the device (row gather), a consumed row does not
        import random
        part = [map(str, walk) for walk in walks][::2]
        random.Random(3).shuffle(part)
        m2 = Word2Vec(part, size=64, window=5, min_count=0, sg=1, iter=1)
        assert m2.corpus_count == 170 and m2._corpus[0].shape == (170, 80)
        used = [map(str, walk) for walk in walks]
        next(used[5])
        m3 = Word2Vec(used, size=64, window=5, min_count=0, sg=1, iter=1)
        assert m3.corpus_count == 340 and m3._corpus[2] == 0          # generic path (ragged token stream)
        # reference dict views of the tables
        J, q = G.alias_nodes[1]
        assert len(J) == nx_G.degree(1) and np.allclose(q, 1.0)
        J, q = G.alias_edges[(1, 2)]
        assert len(J) == nx_G.degree(2)
        J2, q2 = G.get_alias_edge(1, 2)
        assert (J == J2).all() and (q == q2).all()
    finally:
        sys.path.pop(0)


def test_incremental_training_and_device_graph_constructor():
    """gensim's build_vocab + train(new_sentences) on the existing vocabulary, and node2vec.Graph over a
    DeviceGraph (graphs that never exist as networkx objects) with bulk start nodes: the calls bench.py's
    e2e leg makes"""
    from node2vec_by_ecc_b200 import DeviceGraph, Graph, Word2Vec
    e = np.asarray(list(karate_nx().edges()), dtype=np.int64) - 1
    dg = DeviceGraph.from_coo(e[:, 0], e[:, 1], None, 34, undirected=True)
    G = Graph(dg, False, 0.5, 2.0, seed=3, mode="reject")
    G.preprocess_transition_probs()
    every = G.simulate_walks(1, 20)
    assert len(every) == 34 and [w[0] for w in every] == list(range(34))
    model = Word2Vec(size=64, window=5, min_count=0, sg=1, workers=4, iter=1)
    model.build_vocab(every)
    first = model.wv.syn0.copy()
    for step in range(3):
        nodes = np.arange(34, dtype=np.int32)[::-1].copy()
        walks = G.simulate_walks(4, 20, nodes=torch.as_tensor(nodes))
        assert len(walks) == 136 and walks[0][0] == 33
        model.train([map(str, w) for w in walks], total_examples=len(walks), epochs=1, start_alpha=0.025 - 0.005 * step,
                    end_alpha=0.02 - 0.005 * step)
    assert model.pairs_trained > 3000 and model.train_count == 3        # (sample=1e-3 thins a 34-word corpus hard)
    assert np.abs(model.wv.syn0 - first).max() > 1e-3 and np.isfinite(model.wv.syn0).all()
    # sentences in another id space (strings): mapped through the vocabulary, unknown words ignored
    before = model.pairs_trained
    model.train([["1", "2", "zzz", "3", "2", "1"]] * 10, total_examples=10, epochs=1)
    assert 0 < model.pairs_trained - before <= 10 * 20


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


def test_reference_main_py_runs_unmodified_with_the_dropins():
    """src/main.py:92-103 executed as a script (runpy), byte for byte, with node2vec_by_ecc_b200/dropin
    first on sys.path -- where the reference tree is mounted (the build container; not the GPU box)"""
    import runpy
    import pytest
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "main.py")):
        pytest.skip("reference tree not mounted here")
    argv, cwd = sys.argv, os.getcwd()
    sys.path.insert(0, os.path.join(ROOT, "node2vec_by_ecc_b200", "dropin"))
    for m in ("node2vec", "gensim", "gensim.models", "gensim.models.word2vec"):
        sys.modules.pop(m, None)
    try:
        os.chdir(ref)
        sys.argv = ["main.py", "--input", "graph/karate.edgelist", "--num-walks", "4", "--walk-length", "30"]
        ns = runpy.run_path(os.path.join(ref, "src", "main.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = argv
        sys.path.pop(0)
    emb = ns["emb"]
    assert emb.wv.syn0.shape == (34, 128) and np.isfinite(emb.wv.syn0).all()
    assert set(emb.wv.vocab.keys()) == {str(v) for v in range(1, 35)}


def test_graph_survives_pickle():
    import pickle
    from node2vec_by_ecc_b200 import Graph
    G = Graph(karate_nx(), False, 1, 1)
    G.preprocess_transition_probs()
    G2 = pickle.loads(pickle.dumps(G))
    assert len(G2.simulate_walks_on_the_fly(1, 10)) == 34
    # main_link.py:216-226,:277-296: preprocess in the parent, node2vec_walk / simulate_walks in the
    # pool workers -- the unpickled copy rebuilds the tables of the preprocess call on record
    G3 = pickle.loads(pickle.dumps(G))
    w = G3.node2vec_walk(walk_length=12, start_node=1)
    assert len(w) == 12 and w[0] == 1
    assert len(G3.simulate_walks(2, 10, nodes=[1, 2, 3])) == 6
    J, q = G3.alias_edges[(1, 2)]
    J0, q0 = G.alias_edges[(1, 2)]
    assert (J == J0).all() and (q == q0).all()
    assert G3._walk_id_base != G._walk_id_base            # workers do not reuse the parent's Philox streams
    # the popularity variant survives too, and a Graph that was never preprocessed still refuses
    Gp = Graph(karate_nx(), False, 1, 1, popwalk="pop")
    Gp.preprocess_transition_probs_popularity()
    Gp2 = pickle.loads(pickle.dumps(Gp))
    assert Gp2._prep == "pop" and len(Gp2.simulate_walks(1, 10)) == 34
    import pytest
    with pytest.raises(AttributeError):
        pickle.loads(pickle.dumps(Graph(karate_nx(), False, 1, 1))).simulate_walks(1, 5)


def test_popularity_walks_in_rejection_mode_follow_the_reference_laws():
    """popwalk="pop" on a graph whose edge tables "do not fit" (mode="reject"): both entry points of
    main_link.py:206-219,:274-281 run, produce only arcs, and follow the popularity node law on the
    first step (chi-square; the edge laws are tested through DeviceGraph in test_gpu_walks.py)."""
    import numpy as np
    from helpers import chi_square_p
    from node2vec_by_ecc_b200 import Graph
    import networkx as nx
    rng = np.random.RandomState(3)
    G = nx.Graph()
    users, items = list(range(1, 41)), [int("9999999" + str(i)) for i in range(30)]
    for u in users:
        for it in rng.choice(items, size=rng.randint(2, 12), replace=False):
            G.add_edge(u, int(it), weight=int(rng.randint(1, 4)))
    g = Graph(G, False, 0.5, 2.0, popwalk="pop", mode="reject", seed=5)
    g.preprocess_transition_probs_popularity()
    assert g._tables.edge_slots is None
    for walks in (g.simulate_walks(200, 6), g.simulate_walks_on_the_fly(200, 6)):
        assert len(walks) == 200 * G.number_of_nodes()
        for w in walks[:300]:
            assert all(G.has_edge(a, b) for a, b in zip(w[:-1], w[1:]))
    # first step from user 1: weight / len(G[nbr]) (node2vec.py:213-218)
    nbrs = sorted(G.neighbors(1))
    pr = np.asarray([G[1][x]["weight"] / len(G[x]) for x in nbrs], dtype=np.float64)
    walks = g.simulate_walks(20000, 2, nodes=[1])
    obs = np.bincount([nbrs.index(w[1]) for w in walks], minlength=len(nbrs))
    assert chi_square_p(obs, pr / pr.sum()) > 1e-4


def golden_nx(name):
    """networkx graph of a golden case, nodes inserted in the recorded list(G.nodes()) order"""
    import networkx as nx
    from helpers import load_case
    z, g = load_case(name)
    labels = z["labels"]
    G = nx.DiGraph() if int(z["directed"]) else nx.Graph()
    G.add_nodes_from(int(labels[i]) for i in z["order"])
    w = z["w"]
    for u in range(g.n):
        for e in range(g.row_ptr[u], g.row_ptr[u + 1]):
            wt = float(w[e])
            G.add_edge(int(labels[u]), int(labels[g.col[e]]), weight=int(wt) if wt.is_integer() else wt)
    return z, G


def as_lists(z, key_w, key_l):
    labels = z["labels"]
    return [[int(labels[t]) for t in row[:l]] for row, l in zip(z[key_w], z[key_l])]


def test_popularity_walks_through_public_api():
    """main_link.simulate_walk_popularity (main_link.py:206-219) flows: popwalk 'pop' and 'both'"""
    from node2vec_by_ecc_b200 import Graph
    z, nxG = golden_nx("bip_pop_p1_q1")
    G = Graph(nxG, False, float(z["p"]), float(z["q"]), "pop", seed=int(z["seed"]), mode="alias")
    G.preprocess_transition_probs_popularity()
    assert [list(w) for w in G.simulate_walks(int(z["R"]), int(z["L"]))] == as_lists(z, "walks", "lens")
    G2 = Graph(nxG, False, float(z["p"]), float(z["q"]), "pop", seed=int(z["seed"]), mode="alias")
    assert [list(w) for w in G2.simulate_walks_on_the_fly(int(z["R"]), int(z["L"]))] == as_lists(z, "walks_otf", "lens_otf")
    # popwalk == "both": plain walks extended by popularity walks
    G3 = Graph(nxG, False, 1.0, 1.0, "both")
    G3.preprocess_transition_probs()
    walks = G3.simulate_walks(1, 10)
    G3.preprocess_transition_probs_popularity()
    walks.extend(G3.simulate_walks(1, 10))
    assert len(walks) == 2 * nxG.number_of_nodes()


def test_directed_weighted_graph_through_public_api():
    from node2vec_by_ecc_b200 import Graph
    z, nxG = golden_nx("dir_p025_q4")
    G = Graph(nxG, True, float(z["p"]), float(z["q"]), seed=int(z["seed"]), mode="alias")
    G.preprocess_transition_probs()
    assert [list(w) for w in G.simulate_walks(int(z["R"]), int(z["L"]))] == as_lists(z, "walks", "lens")
    sub = [int(z["labels"][i]) for i in z["sub_starts"]]
    G._walk_id_base = 1000
    assert [list(w) for w in G.simulate_walks(2, int(z["L"]), nodes=sub)] == as_lists(z, "walks_sub", "lens_sub")


def test_walk_file_round_trip_and_batched_link_scores(tmp_path):
    """walk file (main_link.py:544-546) -> LineSentence (:340) -> same model; get_roc_score's cosine
    loop (:173-189) as one n2v_cosine_pairs launch"""
    from node2vec_by_ecc_b200 import Graph, LineSentence, Word2Vec
    G = Graph(karate_nx(), False, 0.25, 4.0, seed=5)
    G.preprocess_transition_probs()
    walks = G.simulate_walks(4, 30)
    path = tmp_path / "walks.txt"
    walks.save_walks(str(path))
    lines = path.read_text().splitlines()
    assert len(lines) == len(walks) and lines[0] == " ".join(map(str, walks[0]))
    m1 = Word2Vec([list(map(str, w)) for w in walks], size=32, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1)
    m2 = Word2Vec(LineSentence(str(path)), size=32, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1)
    assert m1.wv.index2word == m2.wv.index2word and np.array_equal(m1.wv.syn0, m2.wv.syn0)
    edges = [(str(a), str(b)) for a, b in list(karate_nx().edges())[:40]] + [("1", "nope"), ("nope", "2")]
    got = m1.wv.similarity_pairs(edges)
    want = [m1.wv.similarity(a, b) if (a in m1.wv and b in m1.wv) else 0.0 for a, b in edges]
    assert np.abs(got - np.asarray(want, dtype=np.float32)).max() < 1e-5 and got[-1] == 0 and got[-2] == 0


def test_all_pairs_top_k_links_matches_brute_force():
    """link_prediction's all-pairs scoring (main_link.py:69-171) on the device vs numpy"""
    from node2vec_by_ecc_b200 import KeyedVectors, Vocab
    rng = np.random.RandomState(0)
    V, d = 300, 128
    kv = KeyedVectors(d)
    kv.index2word = [str(i) for i in range(V)]
    kv.vocab = {w: Vocab(i, 1) for i, w in enumerate(kv.index2word)}
    kv.syn0 = rng.randn(V, d).astype(np.float32)
    e = kv.syn0 / np.linalg.norm(kv.syn0, axis=1, keepdims=True)
    users = [str(i) for i in range(0, 120)]
    items = [str(i) for i in range(120, 300)]
    train = [(str(rng.randint(0, 120)), str(rng.randint(120, 300))) for _ in range(500)]
    train += [(b, a) for a, b in train[:50]]                   # either orientation is excluded
    got = kv.top_k_links(users, items, k=50, exclude=train, block_rows=37)
    S = e[:120] @ e[120:].T
    for a, b in train:
        if int(a) >= 120:
            a, b = b, a
        S[int(a), int(b) - 120] = -np.inf
    flat = np.argsort(-S.ravel(), kind="stable")[:50]
    want = [(str(i // 180), str(120 + i % 180)) for i in flat]
    assert [p for p, _ in got] == want
    assert np.allclose([s for _, s in got], S.ravel()[flat], atol=1e-5)
    # unseparated mode: unordered pairs i < j of one node list
    nodes = [str(i) for i in range(200)]
    got = kv.top_k_links(nodes, None, k=30, exclude=[("3", "7")], block_rows=64)
    S = e[:200] @ e[:200].T
    S[np.tril_indices(200)] = -np.inf
    S[3, 7] = -np.inf
    flat = np.argsort(-S.ravel(), kind="stable")[:30]
    assert [p for p, _ in got] == [(str(i // 200), str(i % 200)) for i in flat]


def test_device_walk_formatter_matches_python_join(tmp_path):
    """n2v_format_walks_* == " ".join(map(str, walk)) per line (main_link.py:544-546): ragged walks,
    negative and 19-digit labels, the '9999999' item prefix, several chunks"""
    import torch
    from node2vec_by_ecc_b200 import WalkCorpus
    rng = np.random.RandomState(1)
    labels = np.concatenate([[0, -7, 9, 10, 99, 100, 2 ** 62, -2 ** 62, 99999991, 999999912345],
                             rng.randint(0, 10 ** 9, size=90)]).astype(np.int64)
    n, L = 1000, 17
    walks = rng.randint(0, len(labels), size=(n, L)).astype(np.int32)
    lens = rng.randint(1, L + 1, size=n).astype(np.int32)
    lens[:5] = [L, 1, 2, L, 1]
    for i in range(n):
        walks[i, lens[i]:] = -1
    c = WalkCorpus(torch.as_tensor(walks).cuda(), torch.as_tensor(lens).cuda(), labels)
    want = "".join(" ".join(str(int(labels[t])) for t in walks[i, :lens[i]]) + "\n" for i in range(n))
    assert c.format_walks().cpu().numpy().tobytes().decode() == want
    p = tmp_path / "w.txt"
    c.save_walks(str(p), chunk_walks=300)
    assert p.read_text() == want
    c2 = WalkCorpus(c.walks, c.lens, None)                # no labels: compact ids
    assert c2.format_walks(0, 3).cpu().numpy().tobytes().decode() == "".join(
        " ".join(str(int(t)) for t in walks[i, :lens[i]]) + "\n" for i in range(3))


def test_walk_file_device_parser_equals_generic_path(tmp_path):
    """LineSentence over an integer walk file is tokenised on the device (n2v_parse_walks_*): same
    tokens, sentences, vocabulary order and trained rows as the generic Python path; tabs, blank
    lines, CRLF, a missing final newline and negative labels included"""
    from node2vec_by_ecc_b200 import LineSentence, Word2Vec
    rng = np.random.RandomState(3)
    lines = [" ".join(str(int(v)) for v in rng.randint(0, 50, size=rng.randint(1, 40))) for _ in range(200)]
    lines[3] = ""                                   # blank line: no sentence
    lines[5] = "7\t8   9 \r"                        # tabs, runs of spaces, CR
    lines[9] = "-3 4 99999991234 -3"
    text = "\n".join(lines)                         # no trailing newline
    p = tmp_path / "walks.txt"
    p.write_text(text)
    m_fast = Word2Vec(LineSentence(str(p)), size=32, window=5, min_count=0, sg=1, iter=1, hogwild_warps=1)
    sents = [l.split() for l in text.split("\n") if l.split()]
    m_ref = Word2Vec(sents, size=32, window=5, min_count=0, sg=1, iter=1, hogwild_warps=1)
    assert m_fast.corpus_count == m_ref.corpus_count == len(sents)
    assert m_fast.wv.index2word == m_ref.wv.index2word
    assert m_fast.pairs_trained == m_ref.pairs_trained and np.array_equal(m_fast.wv.syn0, m_ref.wv.syn0)
    # a non-integer token falls back to the generic path
    p2 = tmp_path / "words.txt"
    p2.write_text("a b c\nb c d\n")
    m2 = Word2Vec(LineSentence(str(p2)), size=8, window=2, min_count=0, sg=1, iter=1, hogwild_warps=1)
    assert sorted(m2.wv.index2word) == ["a", "b", "c", "d"]


def test_fused_similarity_selections_at_sampled_threshold_sizes():
    """sizes at which the thresholds come from a sample (scoring.py): global top-k over 4 M candidate
    pairs and per-user top share over 6,000 users, both exact against numpy"""
    from node2vec_by_ecc_b200 import KeyedVectors, Vocab
    from node2vec_by_ecc_b200.augment import user_edges
    rng = np.random.RandomState(7)
    V, d = 6000, 64
    kv = KeyedVectors(d)
    kv.index2word = [str(i) for i in range(V)]
    kv.vocab = {w: Vocab(i, 1) for i, w in enumerate(kv.index2word)}
    base = rng.randn(40, d).astype(np.float32)                 # clustered rows: a heavy upper tail of scores
    kv.syn0 = (base[rng.randint(0, 40, V)] + 0.7 * rng.randn(V, d)).astype(np.float32)
    e = kv.syn0 / np.linalg.norm(kv.syn0, axis=1, keepdims=True)
    users, items = [str(i) for i in range(2000)], [str(i) for i in range(2000, 4100)]
    train = [(str(rng.randint(0, 2000)), str(rng.randint(2000, 4100))) for _ in range(3000)]
    got = kv.top_k_links(users, items, k=1000, exclude=train)
    S = e[:2000] @ e[2000:4100].T
    for a, b in train:
        S[int(a), int(b) - 2000] = -np.inf
    flat = np.argsort(-S.ravel(), kind="stable")[:1000]
    gs = np.asarray([s for _, s in got])
    assert np.allclose(gs, S.ravel()[flat], atol=2e-6)
    want = {(str(i // 2100), str(2000 + i % 2100)) for i in flat[:990]}     # the last few may tie within float32 round-off
    assert len(want - {p for p, _ in got}) <= 2
    k = int(V * 0.01)
    src, dst, w = user_edges(kv, list(range(V)), "relu-ratio", 0.01, "cos")
    assert src.numel() == V * k and bool((src.view(V, k) == torch.arange(V, device=src.device)[:, None]).all())
    M = e @ e.T
    np.fill_diagonal(M, 0.0)
    kth = -np.sort(-M, axis=1)[:, :k]
    assert np.allclose(w.view(V, k).cpu().numpy(), kth, atol=2e-6)


def test_user_edge_augmentation_matches_brute_force():
    """add_user_edge's selections (main_link.py:358-453) on the device vs the reference's loops
    restated with numpy/scipy"""
    from scipy.stats import pearsonr
    from node2vec_by_ecc_b200 import KeyedVectors, Vocab
    from node2vec_by_ecc_b200.augment import as_tuples, user_edges
    rng = np.random.RandomState(4)
    U, d = 150, 32
    kv = KeyedVectors(d)
    kv.index2word = [str(i) for i in range(U + 20)]
    kv.vocab = {w: Vocab(i, 1) for i, w in enumerate(kv.index2word)}
    kv.syn0 = rng.randn(U + 20, d).astype(np.float32)
    users = list(range(5, 5 + U))

    def sim(a, b, method):
        x, y = kv[str(a)].astype(np.float64), kv[str(b)].astype(np.float64)
        return pearsonr(x, y)[0] if method == "pearson" else float(x @ y / np.linalg.norm(x) / np.linalg.norm(y))

    for method in ("cos", "pearson"):
        M = np.array([[0.0 if a == b else sim(a, b, method) for b in users] for a in users])
        # ratio: top int(U * ratio) per user, weight 1
        got = as_tuples(users, *user_edges(kv, users, "ratio", 0.04, method, block_rows=64))
        k = int(U * 0.04)
        assert len(got) == U * k
        for i, u in enumerate(users):
            mine = {b for a, b, _ in got[i * k:(i + 1) * k]}
            want = {users[j] for j in np.argsort(-M[i], kind="stable")[:k]}
            assert mine == want and all(a == u and w == 1.0 for a, _, w in got[i * k:(i + 1) * k])
        # step / relu: similarity above a threshold
        thre = 0.3
        for mode in ("step", "relu"):
            got = as_tuples(users, *user_edges(kv, users, mode, thre, method, block_rows=64))
            want = {(users[i], users[j]) for i in range(U) for j in range(U) if M[i, j] > thre + 1e-6}
            loose = {(users[i], users[j]) for i in range(U) for j in range(U) if M[i, j] > thre - 1e-6}
            pairs = {(a, b) for a, b, _ in got}
            assert want <= pairs <= loose
            for a, b, w in got[:200]:
                assert abs(w - (1.0 if mode == "step" else M[users.index(a), users.index(b)])) < 1e-5
    got = user_edges(kv, users[:20], "linear", None, "cos")
    assert got[0].numel() == 400

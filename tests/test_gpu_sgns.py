"""GPU: skip-gram negative sampling through the C ABI vs the C restatement of gensim 3.2.0
(oracle/sgns_oracle.c -- PARITY UNPINNED: gensim is absent, see the oracle's header).
 * vocabulary / sub-sampling / cum-table / exp-table preparation: equal to the oracle's
 * sequential device run (1 warp) == one-worker oracle run with the same Philox streams,
   to fp32 round-off (tolerance stated below)
 * Hogwild run: link-prediction AUC within tolerance of the oracle's on the same corpus."""
import numpy as np
import pytest
import torch

import oracle
from helpers import (build_neg_samples, chung_lu_graph, load_case, roc_auc_cosine, split_edges)

pytestmark = pytest.mark.gpu


def corpus_from_golden(name="karate_p1_q1"):
    from node2vec_by_ecc_b200 import WalkCorpus
    z, g = load_case(name)
    walks = torch.as_tensor(z["walks"]).cuda()
    lens = torch.as_tensor(z["lens"]).cuda()
    return z, g, WalkCorpus(walks, lens, None)


def oracle_inputs(model, walks_np):
    """tokens as vocabulary indices + offsets, with the DEVICE's vocabulary tables (so that both
    sides draw identical negatives; the tables themselves are compared separately)."""
    i2w = np.asarray([int(w) for w in model.wv.index2word], dtype=np.int64)
    n_ids = int(walks_np.max()) + 1
    id2index = np.full(n_ids, -1, dtype=np.int32)
    id2index[i2w] = np.arange(len(i2w), dtype=np.int32)
    tok = np.where(walks_np >= 0, id2index[np.maximum(walks_np, 0)], -1).astype(np.int32)
    off = np.arange(walks_np.shape[0] + 1, dtype=np.int64) * walks_np.shape[1]
    counts = np.asarray([model.wv.vocab[w].count for w in model.wv.index2word], dtype=np.int64)
    keep = model._keep_thr.cpu().numpy().view(np.uint32).astype(np.uint64)
    keep = np.where(keep == 0xFFFFFFFF, np.uint64(1) << np.uint64(32), keep)
    cum = model._cum_table.cpu().numpy().view(np.uint32)
    voc = oracle.Vocab(counts, i2w.astype(np.int32), id2index, keep, cum.copy())
    return tok, off, voc


def test_vocab_tables_match_oracle():
    from node2vec_by_ecc_b200 import Word2Vec
    z, g, corpus = corpus_from_golden()
    m = Word2Vec(corpus, size=32, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1)
    vo = oracle.sgns_vocab(z["walks"], g.n, sample=1e-3)
    counts = np.asarray([m.wv.vocab[w].count for w in m.wv.index2word])
    assert (counts == vo.counts).all() and (np.diff(counts) <= 0).all()
    assert counts.sum() == int((z["walks"] >= 0).sum())
    # same multiset per count value (tie order is unspecified in gensim under Python 2)
    ids_dev = np.asarray([int(w) for w in m.wv.index2word])
    for c in np.unique(counts):
        assert sorted(ids_dev[counts == c]) == sorted(vo.index2id[vo.counts == c])
    keep = m._keep_thr.cpu().numpy().view(np.uint32).astype(np.uint64)
    want = np.minimum(vo.sample_int, np.uint64(0xFFFFFFFF))
    assert (np.abs(keep.astype(np.int64) - want.astype(np.int64)) <= 1).all()
    cum = m._cum_table.cpu().numpy().view(np.uint32).astype(np.int64)
    assert (np.abs(cum - vo.cum_table.astype(np.int64)) <= 1).all() and cum[-1] == 2 ** 31 - 1
    # bucket index really brackets bisect_left
    bl = m._bucket_lo.cpu().numpy()
    shift = 31 - m._bucket_bits
    for b in range(0, len(bl), max(1, len(bl) // 257)):
        assert bl[b] == np.searchsorted(cum, b << shift, side="left")


def test_syn0_init_matches_oracle():
    from node2vec_by_ecc_b200 import Word2Vec
    _, _, corpus = corpus_from_golden()
    m = Word2Vec.__new__(Word2Vec)
    Word2Vec.__init__(m, None, size=100, sg=1, min_count=0, seed=7)
    m.build_vocab(corpus)
    want = oracle.sgns_init_syn0(len(m.wv.index2word), 100, seed=7)
    assert np.array_equal(m.wv.syn0, want)
    assert float(m.syn1neg_dev.abs().max()) == 0.0


@pytest.mark.parametrize("size,negative,iters,sample", [(128, 5, 1, 1e-3), (64, 5, 2, 1e-3), (256, 5, 1, 0.0),
                                                        (128, 3, 1, 1e-2), (32, 7, 1, 1e-3)])
def test_sequential_device_run_equals_oracle(size, negative, iters, sample):
    """1 warp == gensim with one worker: same sub-sampling, window shrink, negatives (Philox),
    same update order. fp32 dot products are reduced in a different order (4 partial sums per
    lane + xor tree vs left to right), so rows agree to round-off: tolerance 2e-4 absolute on
    values of magnitude ~1e-2..1, typically 1e-6."""
    from node2vec_by_ecc_b200 import Word2Vec
    z, g, corpus = corpus_from_golden("karate_p025_q4")
    m = Word2Vec(corpus, size=size, window=10, min_count=0, sg=1, iter=iters, negative=negative,
                 sample=sample, seed=1, hogwild_warps=1, shared_negatives=0)
    tok, off, voc = oracle_inputs(m, z["walks"])
    s0, s1, pairs = oracle.sgns_train(tok, off, voc, dim=size, window=10, negative=negative, iters=iters,
                                      workers=1, rng_mode=1, seed=1, subsample=sample > 0)
    assert m.pairs_trained == pairs and pairs > 1000
    d0 = np.abs(m.wv.syn0 - s0).max()
    d1 = np.abs(m.syn1neg_dev.cpu().numpy() - s1).max()
    assert d0 < 2e-4 and d1 < 2e-4, (d0, d1)
    assert np.abs(s0).max() > 0.05            # training moved the rows well away from init


@pytest.mark.parametrize("size,iters,sample,case", [(128, 1, 1e-3, "rndw_p05_q2"), (64, 2, 0.0, "karate_p4_q025"),
                                                   (100, 1, 1e-2, "karate_p1_q1")])
def test_shared_negative_mode_sequential_equals_oracle(size, iters, sample, case):
    """shared_negatives=1 (one negative set per centre, output rows carried in registers across its
    context pairs) == the oracle's shared mode (rng_mode 3 = Philox | shared) run by one worker."""
    from node2vec_by_ecc_b200 import Word2Vec
    z, g, corpus = corpus_from_golden(case)
    m = Word2Vec(corpus, size=size, window=10, min_count=0, sg=1, iter=iters, negative=5, sample=sample,
                 seed=3, hogwild_warps=1, shared_negatives=1)
    tok, off, voc = oracle_inputs(m, z["walks"])
    s0, s1, pairs = oracle.sgns_train(tok, off, voc, dim=size, window=10, negative=5, iters=iters,
                                      workers=1, rng_mode=3, seed=3, subsample=sample > 0)
    assert m.pairs_trained == pairs and pairs > 1000
    d0 = np.abs(m.wv.syn0 - s0).max()
    d1 = np.abs(m.syn1neg_dev.cpu().numpy() - s1).max()
    assert d0 < 2e-4 and d1 < 2e-4, (d0, d1)
    # and it is a different stream of negatives than the per-pair mode
    s0p, _, _ = oracle.sgns_train(tok, off, voc, dim=size, window=10, negative=5, iters=iters,
                                  workers=1, rng_mode=1, seed=3, subsample=sample > 0)
    assert np.abs(s0p - s0).max() > 1e-3


@pytest.mark.parametrize("atomic,hot", [(1, 1 << 30), (0, 1 << 30), (1, 5)])
def test_hot_rows_keep_the_sequential_semantics(monkeypatch, atomic, hot):
    """negatives among the `hot_rows` most frequent words are reduced and re-read pair by pair instead
    of carried across the centre's window (bounded staleness under wide Hogwild): with one warp the run
    still equals the oracle's shared-negative law -- every row hot, the top 5 hot, both update modes"""
    from node2vec_by_ecc_b200 import Word2Vec
    monkeypatch.setenv("N2V_SGNS_HOT_ROWS", str(hot))
    z, g, corpus = corpus_from_golden("rndw_p05_q2")
    m = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, negative=5, seed=3, hogwild_warps=1,
                 shared_negatives=1, atomic_updates=atomic)
    tok, off, voc = oracle_inputs(m, z["walks"])
    s0, s1, pairs = oracle.sgns_train(tok, off, voc, dim=128, window=10, negative=5, iters=1, workers=1, rng_mode=3, seed=3)
    assert m.pairs_trained == pairs
    assert np.abs(m.wv.syn0 - s0).max() < 2e-4 and np.abs(m.syn1neg_dev.cpu().numpy() - s1).max() < 2e-4
    # the rule itself: rows whose count^0.75 share puts more than ~8 copies in flight at the given width
    monkeypatch.delenv("N2V_SGNS_HOT_ROWS")
    T = m.trainer
    assert T.hot_rows(1) == 0 and 0 < T.hot_rows(100000) <= T.V
    assert T.hot_rows(100000) >= T.hot_rows(1000)


def test_tensor_core_window_batch_experiment_trains(monkeypatch):
    """csrc/n2v_sgns_mma.cu (N2V_SGNS_TUNING=16): the same pairs and negative sets as the shared-negative
    kernel, window-batch semantics on TF32 tensor cores -- not a parity mode; here only: same pair count,
    finite tables, rows close to the sequential kernel's"""
    from node2vec_by_ecc_b200 import Word2Vec
    _, _, corpus = corpus_from_golden("rndw_p05_q2")
    ref = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, seed=3, hogwild_warps=1, shared_negatives=1)
    monkeypatch.setenv("N2V_SGNS_TUNING", "16")
    m = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, seed=3, hogwild_warps=1, shared_negatives=1)
    assert m.pairs_trained == ref.pairs_trained and np.isfinite(m.wv.syn0).all()
    a, b = m.wv.syn0, ref.wv.syn0
    cos = (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)
    assert np.abs(a - oracle.sgns_init_syn0(a.shape[0], 128, 3)).max() > 0.01 and cos.mean() > 0.9, cos.mean()


def test_atomic_update_mode_sequential_equals_plain():
    from node2vec_by_ecc_b200 import Word2Vec
    _, _, corpus = corpus_from_golden()
    a = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1, atomic_updates=0)
    b = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1, atomic_updates=1)
    c = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1, atomic_updates=0,
                 shared_negatives=0)
    d = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1, atomic_updates=1,
                 shared_negatives=0)
    assert np.abs(c.wv.syn0 - d.wv.syn0).max() < 2e-4
    assert np.abs(a.wv.syn0 - b.wv.syn0).max() < 2e-4


def test_generic_string_corpus_and_keyedvectors_surface(tmp_path):
    """what learn_embeddings does (main.py:86-87): one-shot map objects of str tokens"""
    from node2vec_by_ecc_b200 import LineSentence, Word2Vec
    z, g, corpus = corpus_from_golden()
    walks = [w for w in corpus]
    sents = [map(str, w) for w in walks]
    m = Word2Vec(sents, size=16, window=10, min_count=0, sg=1, workers=8, iter=1, hogwild_warps=1)
    wv = m.wv
    assert len(wv.vocab) == g.n and set(wv.vocab.keys()) == {str(i) for i in range(g.n)}
    assert wv["3"].shape == (16,) and wv.syn0.dtype == np.float32
    assert abs(wv.similarity("0", "0") - 1.0) < 1e-6 and -1 <= wv.similarity("0", "33") <= 1
    assert m["3"] is not None and m.similarity("1", "2") == wv.similarity("1", "2")
    with pytest.raises(KeyError):
        wv["not-a-node"]
    counts = [wv.vocab[w].count for w in wv.index2word]
    assert counts == sorted(counts, reverse=True) and wv.vocab[wv.index2word[0]].index == 0
    # walk file -> LineSentence -> same vocabulary / same training (main_link.py:340-349)
    path = tmp_path / "walks.txt"
    path.write_text("\n".join(" ".join(map(str, w)) for w in walks) + "\n")
    m2 = Word2Vec(LineSentence(str(path)), size=16, window=10, min_count=0, sg=1, iter=1, hogwild_warps=1)
    assert m2.wv.index2word == wv.index2word and np.array_equal(m2.wv.syn0, wv.syn0)
    out = tmp_path / "emb.txt"
    wv.save_word2vec_format(str(out))
    lines = out.read_text().splitlines()
    assert lines[0] == f"{g.n} 16" and len(lines) == g.n + 1 and len(lines[1].split()) == 17


def test_unsupported_modes_raise():
    from node2vec_by_ecc_b200 import Word2Vec
    with pytest.raises(NotImplementedError):
        Word2Vec([["a", "b"]], sg=0)
    with pytest.raises(NotImplementedError):
        Word2Vec([["a", "b"]], sg=1, hs=1, negative=0)


@pytest.mark.parametrize("shared", [0, 1])
def test_hogwild_auc_matches_oracle(shared):
    """main_link.main protocol (main_link.py:519-565) on a 3k-node heavy-tailed graph: hold out
    50 % of the edges (seed 123), walk the rest (R=5, L=40, p=0.25, q=4), train, score held-out
    edges vs sampled non-edges by cosine, ROC-AUC. Device Hogwild vs the oracle with 8 workers on
    the SAME corpus, mean of 3 seeds each; tolerance = the north-star's +-0.005."""
    from node2vec_by_ecc_b200 import DeviceGraph, WalkCorpus, Word2Vec
    n = 3000
    edges = chung_lu_graph(n, 60000, seed=42, max_deg=600, communities=20)
    tr, te = split_edges(edges)
    dg = DeviceGraph.from_coo(tr[:, 0], tr[:, 1], None, n, undirected=True)
    t = dg.build_alias_tables(0.25, 4.0)
    starts = torch.arange(n, dtype=torch.int32).repeat(5)
    walks, lens = dg.walk_alias(t, starts, 40, seed=9)
    neg = build_neg_samples(n, edges, len(te), seed=1)
    corpus = WalkCorpus(walks, lens, None)
    walks_np = walks.cpu().numpy()
    auc_dev, auc_ref = [], []
    for seed in (1, 2, 3):
        m = Word2Vec(corpus, size=128, window=10, min_count=0, sg=1, workers=8, iter=1, seed=seed,
                     shared_negatives=shared)
        emb = np.zeros((n, 128), dtype=np.float32)
        emb[np.asarray([int(w) for w in m.wv.index2word])] = m.wv.syn0
        auc_dev.append(roc_auc_cosine(emb, te, neg))
        tok, off, voc = oracle_inputs(m, walks_np)
        s0, _, _ = oracle.sgns_train(tok, off, voc, dim=128, window=10, negative=5, iters=1, workers=8,
                                     rng_mode=0, seed=seed)
        emb = np.zeros((n, 128), dtype=np.float32)
        emb[voc.index2id] = s0
        auc_ref.append(roc_auc_cosine(emb, te, neg))
    print('AUC device', auc_dev, 'oracle', auc_ref)
    assert np.mean(auc_ref) > 0.6, (auc_dev, auc_ref)   # the protocol is learning something
    assert abs(np.mean(auc_dev) - np.mean(auc_ref)) <= 0.005, (auc_dev, auc_ref)


@pytest.mark.parametrize("shared,negative", [(1, 5), (0, 5), (0, 3)])
def test_long_ragged_sentences_stream_with_exact_windows(shared, negative):
    """Sentences longer than the 256-token per-warp staging buffer are streamed with a carry-over
    of 2*window kept tokens: windows never break at a chunk boundary. Ragged corpus (0..1000
    tokens per sentence) through the generic string path, all three kernels, vs the oracle."""
    from node2vec_by_ecc_b200 import Word2Vec
    rng = np.random.RandomState(5)
    lens = [1000, 224, 225, 257, 3, 1, 0, 500, 300]
    sents = [[str(t) for t in rng.randint(0, 60, size=n)] for n in lens]
    m = Word2Vec(sents, size=32, window=10, min_count=0, sg=1, iter=1, negative=negative, sample=1e-2,
                 seed=2, hogwild_warps=1, shared_negatives=shared)
    tok = np.asarray([m.wv.vocab[w].index for s in sents for w in s], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    counts = np.asarray([m.wv.vocab[w].count for w in m.wv.index2word], dtype=np.int64)
    keep = m._keep_thr.cpu().numpy().view(np.uint32).astype(np.uint64)
    keep = np.where(keep == 0xFFFFFFFF, np.uint64(1) << np.uint64(32), keep)
    voc = oracle.Vocab(counts, np.arange(len(counts), dtype=np.int32), np.arange(len(counts), dtype=np.int32),
                       keep, m._cum_table.cpu().numpy().view(np.uint32).copy())
    s0, s1, pairs = oracle.sgns_train(tok, off, voc, dim=32, window=10, negative=negative, iters=1, workers=1,
                                      rng_mode=3 if shared else 1, seed=2)
    assert m.pairs_trained == pairs and pairs > 20000
    assert np.abs(m.wv.syn0 - s0).max() < 2e-4 and np.abs(m.syn1neg_dev.cpu().numpy() - s1).max() < 2e-4


@pytest.mark.parametrize("n_parts,size", [(2, 128), (4, 64), (8, 128)])
def test_sharded_tables_sequential_equal_oracle(n_parts, size):
    """n2v_sgns_train_sharded: the same tables spread over n_parts allocations (row i in part
    i % n_parts) -- here all on one device; sequential run == the oracle's shared mode."""
    from node2vec_by_ecc_b200 import PeerSgnsTrainer
    z, g, corpus = corpus_from_golden("karate_p025_q4")
    walks = corpus.walks
    counts = torch.bincount(walks[walks >= 0].to(torch.int64), minlength=g.n)
    tr = PeerSgnsTrainer(counts, dim=size, window=10, negative=5, sample=1e-3, seed=4, local_parts=n_parts)
    tr.train(walks, None, walks.shape[0], walks.shape[1], total_examples=walks.shape[0], sent_per_job=125, grid_warps=1)
    s0_dev, s1_dev = tr.gather()
    order = tr.order.cpu().numpy()
    id2index = np.full(g.n, -1, dtype=np.int32); id2index[order] = np.arange(len(order), dtype=np.int32)
    keep = tr.keep_thr.cpu().numpy().view(np.uint32).astype(np.uint64)
    keep = np.where(keep == 0xFFFFFFFF, np.uint64(1) << np.uint64(32), keep)
    voc = oracle.Vocab(tr.counts.cpu().numpy(), order.astype(np.int32), id2index, keep,
                       tr.cum_table.cpu().numpy().view(np.uint32).copy())
    wn = z["walks"]
    tok = np.where(wn >= 0, id2index[np.maximum(wn, 0)], -1).astype(np.int32)
    off = np.arange(wn.shape[0] + 1, dtype=np.int64) * wn.shape[1]
    s0, s1, pairs = oracle.sgns_train(tok, off, voc, dim=size, window=10, negative=5, iters=1, workers=1, rng_mode=3, seed=4)
    assert int(tr.pairs[0]) == pairs
    assert np.abs(s0_dev.cpu().numpy() - s0).max() < 2e-4 and np.abs(s1_dev.cpu().numpy() - s1).max() < 2e-4

"""CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/_build/libn2v_oracle.so`` (``n2v_oracle.c`` + ``sgns_oracle.c``),
the plain-C restatement of the reference's alias-table build, second-order walk
(/root/reference/src/node2vec.py) and of gensim-3.2.0 skip-gram negative sampling
(behind /root/reference/src/main.py:82-90).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package. The product package
(``node2vec_by_ecc_b200``) never does and has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libn2v_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("n2v_oracle.c", "sgns_oracle.c", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.n2v_oracle_sum_deg_sq.restype = C.c_int64
        _lib.sgns_oracle_vocab.restype = C.c_int32
    return _lib


# --------------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().n2v_oracle_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def walk_uniforms(seed: int, walk_id: int, step: int):
    """The two float64 uniforms injected at (walk_id, step)."""
    u = np.zeros(2, dtype=np.float64)
    lib().n2v_oracle_walk_uniforms(C.c_uint64(seed), C.c_uint64(walk_id), C.c_uint32(step),
                                   _p(u, C.c_double))
    return float(u[0]), float(u[1])


def alias_setup(probs):
    """node2vec.py:240-269 -> (J int64[K], q float64[K])"""
    pr = np.ascontiguousarray(probs, dtype=np.float64)
    K = pr.shape[0]
    J = np.zeros(K, dtype=np.int64)
    q = np.zeros(K, dtype=np.float64)
    rc = lib().n2v_oracle_alias_setup(_p(pr, C.c_double), C.c_int64(K), _p(J, C.c_int64),
                                      _p(q, C.c_double))
    assert rc == 0
    return J, q


@dataclass
class CSR:
    """Sorted-row CSR over compact ids; ``w`` None == unweighted (every weight 1)."""
    row_ptr: np.ndarray      # int64[N+1]
    col: np.ndarray          # int32[nnz], ascending per row
    w: np.ndarray | None     # float64[nnz]

    @property
    def n(self):
        return self.row_ptr.shape[0] - 1

    @property
    def nnz(self):
        return int(self.row_ptr[-1])


def csr_from_coo(src, dst, w, n_nodes, undirected: bool) -> CSR:
    """numpy restatement of what networkx's add_edge / to_undirected gives the reference:
    duplicates collapse (the LAST weight wins), undirected graphs hold both orientations,
    rows are sorted ascending (== sorted(G.neighbors(v)))."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    ww = None if w is None else np.asarray(w, dtype=np.float64)
    if undirected:
        # edge i contributes arcs 2i (as given) and 2i+1 (reversed): "last" is by input order
        src, dst = np.stack([src, dst], 1).ravel(), np.stack([dst, src], 1).ravel()
        if ww is not None:
            ww = np.repeat(ww, 2)
    key = src * np.int64(n_nodes) + dst
    order = np.argsort(key, kind="stable")
    key = key[order]
    last = np.ones(key.shape[0], dtype=bool)
    last[:-1] = key[1:] != key[:-1]
    sel = order[last]
    s, d = src[sel], dst[sel]
    row_ptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(row_ptr, s + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return CSR(row_ptr, d.astype(np.int32), None if ww is None else np.ascontiguousarray(ww[sel]))


@dataclass
class Tables:
    nJ: np.ndarray
    nq: np.ndarray
    etab_ptr: np.ndarray
    eJ: np.ndarray | None
    eq: np.ndarray | None


def sum_deg_sq(g: CSR) -> int:
    return int(lib().n2v_oracle_sum_deg_sq(_p(g.row_ptr, C.c_int64), _p(g.col, C.c_int32),
                                           C.c_int32(g.n)))


def preprocess(g: CSR, p: float, q: float, is_item=None, popwalk_nodes: bool = False,
               edges: bool = True) -> Tables:
    """preprocess_transition_probs (node2vec.py:176-204); popwalk_nodes=True ==
    preprocess_transition_probs_popularity (:206-237)."""
    nnz = g.nnz
    nJ = np.zeros(nnz, dtype=np.int64)
    nq = np.zeros(nnz, dtype=np.float64)
    etab = np.zeros(nnz + 1, dtype=np.int64)
    tot = sum_deg_sq(g)
    eJ = np.zeros(tot, dtype=np.int64) if edges else None
    eq = np.zeros(tot, dtype=np.float64) if edges else None
    it = None if is_item is None else np.ascontiguousarray(is_item, dtype=np.uint8)
    rc = lib().n2v_oracle_preprocess(
        _p(g.row_ptr, C.c_int64), _p(g.col, C.c_int32), _p(g.w, C.c_double), C.c_int32(g.n),
        C.c_double(p), C.c_double(q), _p(it, C.c_uint8), C.c_int(int(popwalk_nodes)),
        _p(nJ, C.c_int64), _p(nq, C.c_double), _p(etab, C.c_int64), _p(eJ, C.c_int64),
        _p(eq, C.c_double))
    assert rc == 0
    return Tables(nJ, nq, etab, eJ, eq)


def edge_table(g: CSR, p, q, src, dst, popwalk=False):
    K = int(g.row_ptr[dst + 1] - g.row_ptr[dst])
    J = np.zeros(K, dtype=np.int64)
    qq = np.zeros(K, dtype=np.float64)
    lib().n2v_oracle_edge_table(_p(g.row_ptr, C.c_int64), _p(g.col, C.c_int32), _p(g.w, C.c_double),
                                C.c_double(p), C.c_double(q), C.c_int(int(popwalk)),
                                C.c_int32(src), C.c_int32(dst), _p(J, C.c_int64), _p(qq, C.c_double))
    return J, qq


def transition_row(g: CSR, p, q, src, dst):
    K = int(g.row_ptr[dst + 1] - g.row_ptr[dst])
    pr = np.zeros(K, dtype=np.float64)
    lib().n2v_oracle_transition_row(_p(g.row_ptr, C.c_int64), _p(g.col, C.c_int32),
                                    _p(g.w, C.c_double), C.c_double(p), C.c_double(q),
                                    C.c_int32(src), C.c_int32(dst), _p(pr, C.c_double))
    return pr


def walks_alias(g: CSR, t: Tables, starts, L: int, seed: int, walk_id_base: int = 0):
    """simulate_walks (node2vec.py:81-95) with Philox uniforms; -> (walks[n,L] padded -1, lens)"""
    st = np.ascontiguousarray(starts, dtype=np.int32)
    n = st.shape[0]
    walks = np.empty((n, L), dtype=np.int32)
    lens = np.zeros(n, dtype=np.int32)
    rc = lib().n2v_oracle_walks_alias(
        _p(g.row_ptr, C.c_int64), _p(g.col, C.c_int32), _p(t.nJ, C.c_int64), _p(t.nq, C.c_double),
        _p(t.etab_ptr, C.c_int64), _p(t.eJ, C.c_int64), _p(t.eq, C.c_double),
        _p(st, C.c_int32), C.c_int64(n), C.c_int32(L), C.c_uint64(seed), C.c_uint64(walk_id_base),
        _p(walks, C.c_int32), _p(lens, C.c_int32))
    assert rc == 0
    return walks, lens


def walks_on_the_fly(g: CSR, p, q, starts, L: int, seed: int, walk_id_base: int = 0,
                     is_item=None, popwalk: bool = False):
    """simulate_walks_on_the_fly (node2vec.py:97-111) with Philox uniforms."""
    st = np.ascontiguousarray(starts, dtype=np.int32)
    n = st.shape[0]
    walks = np.empty((n, L), dtype=np.int32)
    lens = np.zeros(n, dtype=np.int32)
    deg = np.diff(g.row_ptr)
    it = None if is_item is None else np.ascontiguousarray(is_item, dtype=np.uint8)
    rc = lib().n2v_oracle_walks_on_the_fly(
        _p(g.row_ptr, C.c_int64), _p(g.col, C.c_int32), _p(g.w, C.c_double), C.c_double(p),
        C.c_double(q), _p(it, C.c_uint8), C.c_int(int(popwalk)),
        C.c_int64(int(deg.max()) if deg.size else 0), _p(st, C.c_int32), C.c_int64(n),
        C.c_int32(L), C.c_uint64(seed), C.c_uint64(walk_id_base), _p(walks, C.c_int32),
        _p(lens, C.c_int32))
    assert rc == 0
    return walks, lens


# --------------------------------------------------------------------------------------
@dataclass
class Vocab:
    counts: np.ndarray       # int64[V], vocabulary order (count descending, first-seen ties)
    index2id: np.ndarray     # int32[V]
    id2index: np.ndarray     # int32[n_ids], -1 when absent
    sample_int: np.ndarray   # uint64[V]
    cum_table: np.ndarray    # uint32[V]

    @property
    def V(self):
        return self.counts.shape[0]


def sgns_vocab(tokens, n_ids: int, sample: float = 1e-3) -> Vocab:
    tok = np.ascontiguousarray(tokens, dtype=np.int32).ravel()
    counts = np.zeros(n_ids, dtype=np.int64)
    i2id = np.zeros(n_ids, dtype=np.int32)
    id2i = np.zeros(n_ids, dtype=np.int32)
    V = lib().sgns_oracle_vocab(_p(tok, C.c_int32), C.c_int64(tok.shape[0]), C.c_int32(n_ids),
                                _p(counts, C.c_int64), _p(i2id, C.c_int32), _p(id2i, C.c_int32))
    assert V >= 0
    counts = counts[:V].copy()
    si = np.zeros(V, dtype=np.uint64)
    cum = np.zeros(V, dtype=np.uint32)
    lib().sgns_oracle_prepare(_p(counts, C.c_int64), C.c_int32(V), C.c_double(sample),
                              _p(si, C.c_uint64), _p(cum, C.c_uint32))
    return Vocab(counts, i2id[:V].copy(), id2i, si, cum)


def sgns_init_syn0(V: int, dim: int, seed: int = 1):
    syn0 = np.zeros((V, dim), dtype=np.float32)
    lib().sgns_oracle_init_syn0(_p(syn0, C.c_float), C.c_int32(V), C.c_int32(dim), C.c_uint64(seed))
    return syn0


def sgns_exp_table():
    t = np.zeros(1000, dtype=np.float32)
    lib().sgns_oracle_exp_table(_p(t, C.c_float))
    return t


def sgns_train(tok_idx, sent_off, vocab: Vocab, dim=128, window=10, negative=5, alpha=0.025,
               min_alpha=1e-4, iters=1, batch_words=10000, workers=1, rng_mode=0, seed=1,
               subsample=True, syn0=None, syn1neg=None):
    """Word2Vec(sg=1, hs=0).train restated; tok_idx = vocabulary indices (-1 = skip).
    Returns (syn0, syn1neg, pairs)."""
    tok = np.ascontiguousarray(tok_idx, dtype=np.int32).ravel()
    off = np.ascontiguousarray(sent_off, dtype=np.int64)
    V = vocab.V
    if syn0 is None:
        syn0 = sgns_init_syn0(V, dim, seed)
    if syn1neg is None:
        syn1neg = np.zeros((V, dim), dtype=np.float32)
    pairs = C.c_int64(0)
    rc = lib().sgns_oracle_train(
        _p(tok, C.c_int32), _p(off, C.c_int64), C.c_int64(off.shape[0] - 1), C.c_int32(V),
        C.c_int32(dim), C.c_int32(window), C.c_int32(negative),
        _p(vocab.sample_int, C.c_uint64) if subsample else None, _p(vocab.cum_table, C.c_uint32),
        C.c_float(alpha), C.c_float(min_alpha), C.c_int32(iters), C.c_int32(batch_words),
        C.c_int32(workers), C.c_int32(rng_mode), C.c_uint64(seed), _p(syn0, C.c_float),
        _p(syn1neg, C.c_float), C.byref(pairs))
    assert rc == 0
    return syn0, syn1neg, int(pairs.value)


def vocab_from_counts(counts_by_id, sample: float = 1e-3) -> Vocab:
    """Vocabulary from given per-id counts (ids with count 0 are absent): count descending,
    ties by id -- used when the corpus itself is only sampled (bench.py)."""
    c = np.asarray(counts_by_id, dtype=np.int64)
    order = np.argsort(-c, kind="stable")
    V = int((c > 0).sum())
    order = order[:V].astype(np.int32)
    counts = np.ascontiguousarray(c[order])
    id2i = np.full(c.shape[0], -1, dtype=np.int32)
    id2i[order] = np.arange(V, dtype=np.int32)
    si = np.zeros(V, dtype=np.uint64)
    cum = np.zeros(V, dtype=np.uint32)
    lib().sgns_oracle_prepare(_p(counts, C.c_int64), C.c_int32(V), C.c_double(sample),
                              _p(si, C.c_uint64), _p(cum, C.c_uint32))
    return Vocab(counts, order, id2i, si, cum)


# ---- block-partitioned SGNS (csrc/n2v_sgns_block.cu) ---------------------------------------------
def sgns_make_groups(tok_idx, sent_off, vocab: Vocab, part: int, n_parts: int, window=10, seed=1, epoch=0,
                     sent_id_base=0, subsample=True, neg_group=1):
    """-> list of n_parts uint32[n_b] word streams: the pairs whose centre is in `part`, stream b =
    context in part b, grouped per centre occurrence: {0x80000000 | centre local row, sentence index,
    position | pairs << 16, the centre's 5 negatives (local rows of `part`)}, then the context local rows."""
    tok = np.ascontiguousarray(tok_idx, dtype=np.int32).ravel()
    off = np.ascontiguousarray(sent_off, dtype=np.int64)
    lg = int(n_parts).bit_length() - 1
    assert 1 << lg == n_parts
    si = _p(vocab.sample_int, C.c_uint64) if subsample else None
    lens = np.zeros(n_parts, dtype=np.int64)
    args = (_p(tok, C.c_int32), _p(off, C.c_int64), C.c_int64(off.shape[0] - 1), C.c_int64(sent_id_base),
            C.c_int32(window), si, C.c_uint64(seed), C.c_uint32(epoch), C.c_int32(part), C.c_int32(lg),
            C.c_int32(vocab.V), _p(vocab.cum_table, C.c_uint32), C.c_int32(neg_group))
    assert lib().sgns_oracle_make_groups(*args, _p(lens, C.c_int64), None, None) == 0
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    words = np.zeros(max(int(offs[-1]), 1), dtype=np.uint32)
    assert lib().sgns_oracle_make_groups(*args, _p(lens, C.c_int64), _p(offs, C.c_int64), _p(words, C.c_uint32)) == 0
    return [words[offs[b]:offs[b + 1]] for b in range(n_parts)]


def sgns_train_groups(words, syn0_part, syn1_part, *, alpha=0.025, min_alpha=1e-4, total_examples=1,
                      example_base=0, sent_per_job=1) -> int:
    """one stream against (syn0 part, syn1neg part `part`), in place; returns the pairs trained"""
    wd = np.ascontiguousarray(words, dtype=np.uint32)
    assert syn0_part.dtype == np.float32 and syn1_part.dtype == np.float32
    assert syn0_part.flags.c_contiguous and syn1_part.flags.c_contiguous
    pairs = C.c_int64(0)
    rc = lib().sgns_oracle_train_groups(
        _p(wd, C.c_uint32), C.c_int64(wd.shape[0]), _p(syn0_part, C.c_float), _p(syn1_part, C.c_float),
        C.c_int32(syn0_part.shape[1]), C.c_float(alpha), C.c_float(min_alpha), C.c_int64(total_examples),
        C.c_int64(example_base), C.c_int64(sent_per_job), C.byref(pairs))
    assert rc == 0, rc
    return int(pairs.value)


def sgns_block_pool(tok_idx, sent_off, vocab: Vocab, parts0, parts1, *, window=10, alpha=0.025, min_alpha=1e-4,
                    total_examples=1, example_base=0, sent_per_job=1, neg_group=1, seed=1, epoch=0, sent_id_base=0,
                    subsample=True):
    """One pool of sentences through the block schedule, in the order n GPUs would run it: sub-step e,
    GPU k trains stream (k, (k + e) % n) against syn1neg part k and syn0 part (k + e) % n.
    parts0 / parts1: lists of float32[rows_k, dim] (updated in place). Returns the pair count."""
    n = len(parts0)
    streams = [sgns_make_groups(tok_idx, sent_off, vocab, k, n, window, seed, epoch, sent_id_base, subsample, neg_group)
               for k in range(n)]
    pairs = 0
    for e in range(n):
        for k in range(n):
            b = (k + e) % n
            if len(streams[k][b]):
                pairs += sgns_train_groups(streams[k][b], parts0[b], parts1[k], alpha=alpha, min_alpha=min_alpha,
                                           total_examples=total_examples, example_base=example_base,
                                           sent_per_job=sent_per_job)
    return pairs

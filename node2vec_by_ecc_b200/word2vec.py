"""Drop-in for the gensim 3.2.0 surface the reference uses (src/main.py:82-90,
src/main_link.py:36-41,:304-349,:43-61,:127-131,:173-189): ``Word2Vec(sentences, size=, window=,
min_count=, sg=1, workers=, iter=)`` -> model with ``.wv`` (KeyedVectors-compatible), and
``LineSentence``. Training runs in libn2v_b200.so (n2v_sgns_train); there is no CPU fallback.

Only what gensim does for ``sg=1, hs=0, negative>0`` is implemented; other modes raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from ._lib import N2VError, SgnsParams, check, lib, ptr, require_cuda, stream
from .walker import WalkCorpus, rows_of


class Vocab:
    """gensim.models.keyedvectors.Vocab: .index, .count, .sample_int"""
    __slots__ = ("index", "count", "sample_int")

    def __init__(self, index, count, sample_int=None):
        self.index, self.count, self.sample_int = index, count, sample_int

    def __repr__(self):
        return f"Vocab(count:{self.count}, index:{self.index})"


class LineSentence:
    """gensim.models.word2vec.LineSentence: one sentence per line, whitespace separated
    (the walk files of main_link.py:237-239,544-546), at most max_sentence_length words each."""

    def __init__(self, source, max_sentence_length=10000, limit=None):
        self.source, self.max_sentence_length, self.limit = source, max_sentence_length, limit

    def __iter__(self):
        close = False
        f = self.source
        if isinstance(f, (str, os.PathLike)):
            f = open(f, "r")
            close = True
        try:
            if hasattr(f, "seek"):
                f.seek(0)
            for n, line in enumerate(f):
                if self.limit is not None and n >= self.limit:
                    break
                if isinstance(line, bytes):
                    line = line.decode("utf8")
                words = line.split()
                for i in range(0, len(words), self.max_sentence_length):
                    yield words[i:i + self.max_sentence_length]
        finally:
            if close:
                f.close()


class KeyedVectors:
    """The part of gensim.models.KeyedVectors the reference touches."""

    def __init__(self, vector_size=0):
        self.vocab = {}
        self.index2word = []
        self.vector_size = vector_size
        self._syn0_dev_ = None
        self._syn0_host = None
        self._gather = None               # multi-GPU: () -> full device table, called when the rows are first read

    @property
    def _syn0_dev(self):
        if self._syn0_dev_ is None and self._gather is not None:
            self._syn0_dev_ = self._gather()
            self._gather = None
        return self._syn0_dev_

    @_syn0_dev.setter
    def _syn0_dev(self, t):
        self._syn0_dev_, self._gather = t, None

    @property
    def syn0(self):
        """float32[V, d] numpy, copied from the device lazily"""
        if self._syn0_host is None and self._syn0_dev is not None:
            self._syn0_host = self._syn0_dev.cpu().numpy()
        return self._syn0_host

    @syn0.setter
    def syn0(self, v):
        self._syn0_host = np.asarray(v, dtype=np.float32)
        self._syn0_dev = None

    vectors = syn0

    def __contains__(self, word):
        return word in self.vocab

    def word_vec(self, word):
        if word not in self.vocab:
            raise KeyError("word '%s' not in vocabulary" % word)
        return self.syn0[self.vocab[word].index]

    def __getitem__(self, words):
        if isinstance(words, (str, bytes)):
            return self.word_vec(words)
        return np.vstack([self.word_vec(w) for w in words])

    def similarity(self, w1, w2):
        """cosine of the two rows (link_score 'cos', main_link.py:43-49)"""
        a, b = self.word_vec(w1), self.word_vec(w2)
        return float(np.dot(a / np.linalg.norm(a), b / np.linalg.norm(b)))

    def similarity_pairs(self, pairs):
        """Batched link_score 'cos' (main_link.py:43-49) for an iterable of (a, b) words on the
        device (n2v_cosine_pairs): float32[n]; a pair with a word missing from the vocabulary
        scores 0, like the reference's except branch. What get_roc_score (:173-189) loops over."""
        dev = require_cuda()
        pairs = list(pairs)
        v = self.vocab
        ia = np.fromiter((v[a].index if a in v else -1 for a, _ in pairs), dtype=np.int32, count=len(pairs))
        ib = np.fromiter((v[b].index if b in v else -1 for _, b in pairs), dtype=np.int32, count=len(pairs))
        emb = self._syn0_dev if self._syn0_dev is not None else torch.as_tensor(self._syn0_host).to(dev)
        emb = emb.contiguous()
        da, db = torch.as_tensor(ia).to(dev), torch.as_tensor(ib).to(dev)
        out = torch.empty(len(pairs), dtype=torch.float32, device=dev)
        check(lib().n2v_cosine_pairs(ptr(emb), C.c_int32(emb.shape[1]), ptr(da), ptr(db), C.c_int64(len(pairs)),
                                     ptr(out), stream()))
        return out.cpu().numpy()

    def top_k_links(self, words_a, words_b=None, k=10, exclude=(), block_rows=None):
        """All-pairs link scoring of link_prediction / make_links_and_score / links_score
        (main_link.py:69-171): cosine of every candidate pair -- words_a x words_b ("separated"
        user x item mode) or, with words_b None, every unordered pair i < j of words_a -- minus the
        `exclude` pairs (train edges, either orientation), global top-k by score.
        -> list of ((a, b), score), best first. Tiles of the score matrix are computed and selected
        in one kernel (n2v_sim_threshold); the matrix itself is never written (scoring.top_k_links)."""
        from .scoring import top_k_links
        return top_k_links(self, words_a, words_b, k, exclude)

    def most_similar(self, positive, topn=10):
        if isinstance(positive, (str, bytes)):
            positive = [positive]
        m = self.syn0 / np.linalg.norm(self.syn0, axis=1, keepdims=True)
        v = np.mean([m[self.vocab[w].index] for w in positive], axis=0)
        d = m @ (v / np.linalg.norm(v))
        skip = {self.vocab[w].index for w in positive}
        out = [(self.index2word[i], float(d[i])) for i in np.argsort(-d) if i not in skip]
        return out[:topn]

    def save_word2vec_format(self, fname, fvocab=None, binary=False, total_vec=None):
        """text format read back by utils.emb_file_to_user_dict (utils.py:417-426):
        'V d' then 'word v1 ... vd', vocabulary (count-descending) order."""
        if binary:
            raise NotImplementedError("binary word2vec format is not needed by the reference")
        syn0 = self.syn0
        with open(fname, "w") as f:
            f.write("%d %d\n" % (len(self.index2word), self.vector_size))
            for i, word in enumerate(self.index2word):
                f.write("%s %s\n" % (word, " ".join("%f" % v for v in syn0[i])))


class SgnsTrainer:
    """Device state of one SGNS job -- vocabulary tables (scale_vocab / make_cum_table), syn0 and
    syn1neg -- and the launcher of n2v_sgns_train. Word2Vec below is the gensim-shaped front;
    bench.py drives this class directly on device-resident walk buffers."""

    def __init__(self, counts_by_id: torch.Tensor, dim=128, window=10, negative=5, sample=1e-3, seed=1,
                 alpha=0.025, min_alpha=1e-4, min_count=0, batch_words=10000, first_seen=None):
        dev = require_cuda()
        L = lib()
        counts = counts_by_id.to(device=dev, dtype=torch.int64).contiguous()
        n_ids = int(counts.shape[0])
        self.dim, self.window, self.negative, self.sample = int(dim), int(window), int(negative), float(sample)
        self.seed, self.alpha, self.min_alpha, self.batch_words = int(seed), float(alpha), float(min_alpha), int(batch_words)
        keep = counts >= max(int(min_count), 1)
        # count descending, ties by first appearance in the corpus (`first_seen`: position of the id's first
        # token; ids of generic corpora already are in first-seen order) -- gensim's own tie order is dict
        # order, unspecified in Python 2; first-seen makes every ingest path of one corpus agree
        kept = torch.where(keep, counts, torch.zeros_like(counts))
        if first_seen is None:
            order = torch.sort(kept, descending=True, stable=True).indices
        else:
            by_first = torch.argsort(first_seen.to(dev), stable=True)
            order = by_first[torch.sort(kept[by_first], descending=True, stable=True).indices]
        V = int(keep.sum().item())
        if V == 0:
            raise RuntimeError("you must first build vocabulary before training the model")
        self.V = V
        self.order = order[:V].contiguous()                       # vocabulary index -> token id
        self.counts = counts[self.order].contiguous()
        self.raw_words = int(counts.sum().item())
        self.vocab_of_id = torch.full((n_ids,), -1, dtype=torch.int32, device=dev)
        self.vocab_of_id[self.order] = torch.arange(V, dtype=torch.int32, device=dev)
        self.bucket_bits = int(min(20, max(4, int(np.ceil(np.log2(V))) + 2)))
        self.keep_thr = torch.empty(V, dtype=torch.int32, device=dev)
        self.cum_table = torch.empty(V, dtype=torch.int32, device=dev)
        self.bucket_lo = torch.empty((1 << self.bucket_bits) + 1, dtype=torch.int32, device=dev)
        ws_bytes = int(L.n2v_sgns_prepare_workspace_bytes(C.c_int32(V)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(L.n2v_sgns_prepare(ptr(self.counts), C.c_int32(V), C.c_double(self.sample), ptr(self.keep_thr),
                                 ptr(self.cum_table), ptr(self.bucket_lo), C.c_int32(self.bucket_bits),
                                 ptr(ws), C.c_size_t(ws_bytes), stream()))
        self.pairs = torch.zeros(2, dtype=torch.int64, device=dev)   # [pairs, carried centres]
        self._neg_prob = None
        self.reset_weights()

    def reset_weights(self):
        dev = self.counts.device
        self.syn0 = torch.empty((self.V, self.dim), dtype=torch.float32, device=dev)
        self.syn1neg = torch.empty((self.V, self.dim), dtype=torch.float32, device=dev)
        check(lib().n2v_sgns_init(ptr(self.syn0), ptr(self.syn1neg), C.c_int32(self.V), C.c_int32(self.dim),
                                  C.c_uint64(self.seed), stream()))

    def hot_rows(self, grid_warps: int) -> int:
        """Vocabulary rows (a prefix: the table is count-sorted) whose negatives the shared-negative kernels
        must not carry in registers. A negative row is held by about c = grid_warps * 5 * p warps at once
        (p = its share of the count^0.75 mass), each copy stale by what the others add meanwhile; rows with
        c above N2V_SGNS_HOT_COPIES (default 1) are re-read and reduced pair by pair instead. Measured on
        a Zipf(1.0) user-item graph (tests/test_gpu_auc.py, scripts/hot_rows_sweep.py): carrying everything
        at full width moved the link-prediction AUC by +0.019 against the same law run narrow or on the
        CPU; +0.009 with only the hubs (c > 8) uncarried, +0.0006 with c > 0.3. N2V_SGNS_HOT_ROWS overrides
        the count directly."""
        env = os.environ.get("N2V_SGNS_HOT_ROWS")
        if env is not None:
            return int(env)
        if self._neg_prob is None:
            cum = self.cum_table.to(torch.int64)
            self._neg_prob = (torch.diff(cum, prepend=cum.new_zeros(1)).to(torch.float64) / float(cum[-1].item())).cpu()
        tau = float(os.environ.get("N2V_SGNS_HOT_COPIES", "1")) / (5.0 * max(int(grid_warps), 1))
        return int((self._neg_prob > tau).sum().item())

    def default_hogwild_warps(self, shared: bool = False) -> int:
        """Concurrent sentences. gensim runs `workers` (8-12) sentences at a time against the shared
        tables; the GPU runs thousands. Staleness scales with width/V, so the width is capped at
        V/4 for small vocabularies (scripts/auc_sweep.py: |dAUC| <= 0.003 up to V/2 with atomic
        updates) and at the machine width (24 resident warps per SM) otherwise."""
        sms = int(lib().n2v_sm_count())
        return int(max(4, min(sms * (20 if shared else 24), self.V // 4)))

    def train(self, tokens, sent_off, n_sent, stride, *, total_examples, example_base=0, sent_id_base=0,
              epoch=0, sent_per_job=125, grid_warps=None, atomic_updates=1, alpha=None, min_alpha=None,
              tuning=None, negative_sharing=0, vocab_of_id=None, hot_rows=None):
        """One pass over n_sent sentences (asynchronous on the current stream); self.pairs
        (device int64) accumulates the (centre, context) pairs trained. vocab_of_id: token id ->
        vocabulary row (-1 = not in the vocabulary) when the tokens are not in the id space the
        trainer was built over."""
        P = SgnsParams()
        P.V, P.dim, P.window, P.negative = self.V, self.dim, self.window, self.negative
        P.bucket_bits, P.max_sentence_len = self.bucket_bits, 10000
        P.alpha0 = self.alpha if alpha is None else float(alpha)
        P.min_alpha = self.min_alpha if min_alpha is None else float(min_alpha)
        P.total_examples, P.example_base = int(total_examples), int(example_base)
        P.sent_per_job = max(1, int(sent_per_job))
        P.epoch, P.seed = int(epoch), self.seed
        P.grid_warps = int(grid_warps or self.default_hogwild_warps(bool(negative_sharing)))
        P.atomic_updates = int(atomic_updates)
        P.negative_sharing = int(negative_sharing)
        P.tuning = int(os.environ.get("N2V_SGNS_TUNING", "0")) if tuning is None else int(tuning)
        P.hot_rows = (self.hot_rows(P.grid_warps) if hot_rows is None else int(hot_rows)) if negative_sharing else 0
        check(lib().n2v_sgns_train(ptr(tokens), ptr(sent_off), C.c_int64(n_sent), C.c_int32(stride),
                                   C.c_int64(sent_id_base), ptr(self.vocab_of_id if vocab_of_id is None else vocab_of_id),
                                   ptr(self.keep_thr if self.sample > 0 else None), ptr(self.cum_table),
                                   ptr(self.bucket_lo), C.byref(P), ptr(self.syn0), ptr(self.syn1neg),
                                   ptr(self.pairs), stream()))


class PeerSgnsTrainer(SgnsTrainer):
    """SgnsTrainer whose syn0 / syn1neg are ONE logical pair of tables spread over `n_parts`
    allocations: vocabulary row i lives in part i % n_parts at local row i // n_parts
    (n2v_sgns_train_sharded). Two uses:
      * multi-GPU (one process per GPU, torch.distributed initialised): every rank allocates its own
        part, the parts are mapped into every process over NVLink (CUDA IPC), and all ranks train
        their own walks against the same tables -- gensim's shared-memory Hogwild across GPUs; no
        replicas, so nothing to average (dist.py explains why replicas do not survive sparse sync);
      * single process with `local_parts` > 1: all parts on this device (tests the addressing)."""

    def __init__(self, counts_by_id, *args, local_parts: int = 0, **kw):
        from . import dist as D
        self._rank, self._world = D.world()
        self._local_parts = int(local_parts)
        self.n_parts = self._local_parts if self._local_parts else self._world
        if self.n_parts not in (1, 2, 4, 8):
            raise ValueError("the tables can be spread over 1, 2, 4 or 8 parts")
        super().__init__(counts_by_id, *args, **kw)
        if self.dim % 4 or self.dim > 128:
            raise ValueError("sharded tables need dim to be a multiple of 4, <= 128")

    def reset_weights(self):
        from . import dist as D
        dev = self.counts.device
        W = self.n_parts
        mine = range(W) if self._local_parts else [self._rank]
        self.parts0, self.parts1 = {}, {}
        for k in mine:
            n_local = max((self.V - k + W - 1) // W, 1)
            self.parts0[k] = torch.empty((n_local, self.dim), dtype=torch.float32, device=dev)
            self.parts1[k] = torch.empty((n_local, self.dim), dtype=torch.float32, device=dev)
            check(lib().n2v_sgns_init_part(ptr(self.parts0[k]), ptr(self.parts1[k]), C.c_int32(self.V), C.c_int32(self.dim),
                                           C.c_uint64(self.seed), C.c_int32(k), C.c_int32(W), stream()))
        torch.cuda.synchronize()
        if self._local_parts:
            a0 = [self.parts0[k].data_ptr() for k in range(W)]
            a1 = [self.parts1[k].data_ptr() for k in range(W)]
        else:
            a0 = D.exchange_peer_pointers(self.parts0[self._rank])
            a1 = D.exchange_peer_pointers(self.parts1[self._rank])
        self._p0 = (C.c_void_p * W)(*a0)
        self._p1 = (C.c_void_p * W)(*a1)
        self.syn0 = self.syn1neg = None

    def train(self, tokens, sent_off, n_sent, stride, *, total_examples, example_base=0, sent_id_base=0,
              epoch=0, sent_per_job=125, grid_warps=None, atomic_updates=1, alpha=None, min_alpha=None,
              tuning=None, negative_sharing=1):
        if not negative_sharing or self.dim > 128 or self.negative != 5:
            raise NotImplementedError("sharded tables run the shared-negative kernel (dim <= 128, negative = 5)")
        P = SgnsParams()
        P.V, P.dim, P.window, P.negative = self.V, self.dim, self.window, self.negative
        P.bucket_bits, P.max_sentence_len = self.bucket_bits, 10000
        P.alpha0 = self.alpha if alpha is None else float(alpha)
        P.min_alpha = self.min_alpha if min_alpha is None else float(min_alpha)
        P.total_examples, P.example_base = int(total_examples), int(example_base)
        P.sent_per_job = max(1, int(sent_per_job))
        P.epoch, P.seed = int(epoch), self.seed
        P.grid_warps = int(grid_warps or self.default_hogwild_warps(True))
        P.atomic_updates, P.negative_sharing, P.tuning = 1, 1, 0
        P.hot_rows = self.hot_rows(P.grid_warps)
        check(lib().n2v_sgns_train_sharded(ptr(tokens), ptr(sent_off), C.c_int64(n_sent), C.c_int32(stride),
                                           C.c_int64(sent_id_base), ptr(self.vocab_of_id),
                                           ptr(self.keep_thr if self.sample > 0 else None), ptr(self.cum_table),
                                           ptr(self.bucket_lo), C.byref(P), self._p0, self._p1, C.c_int32(self.n_parts),
                                           ptr(self.pairs), stream()))

    def gather(self):
        """-> (syn0, syn1neg) float32[V, dim] on this device, rows in vocabulary order"""
        import torch.distributed as tdist
        W, dev = self.n_parts, self.counts.device
        out = []
        for parts in (self.parts0, self.parts1):
            full = torch.empty((self.V, self.dim), dtype=torch.float32, device=dev)
            if self._local_parts or W == 1:
                got = [parts[k] for k in sorted(parts)]
            else:
                n0 = (self.V + W - 1) // W
                mine = torch.zeros((n0, self.dim), dtype=torch.float32, device=dev)
                mine[: parts[self._rank].shape[0]] = parts[self._rank]
                got = [torch.empty_like(mine) for _ in range(W)]
                tdist.all_gather(got, mine)
            for k in range(W):
                rows = (self.V - k + W - 1) // W
                full[k::W] = got[k][:rows]
            out.append(full)
        return out[0], out[1]


class BlockSgnsTrainer(SgnsTrainer):
    """Block-partitioned SGNS: the multi-GPU trainer (csrc/n2v_sgns_block.cu).

    syn0 / syn1neg are cut into n_parts row sets (vocabulary row i -> part i % n_parts). Every
    train() call is one POOL of walks: its (centre, context) pairs are bucketed by (part of the
    centre, part of the context); bucket (k, b) touches only syn1neg part k and syn0 part b. GPU k
    keeps syn1neg part k, trains bucket (k, (k + e) % n) in sub-step e and hands the syn0 part it
    holds to GPU k - 1 between sub-steps (dist.ring_pass). No replicas, nothing to average, every
    row has one writer GPU at a time.
      * one process per GPU (torch.distributed initialised): n_parts = world size; every rank calls
        train() with ITS walks [n_sent, stride] of the pool; the pool is rank 0's walks, then rank
        1's, ... (all-gathered, 4 bytes per token), sentence ids sent_id_base + position in pool;
      * single process, local_parts = n: all parts on this device, buckets run in the order the n
        GPUs would run them (the exact emulation the parity and AUC tests use; n = 1 is a plain
        single-GPU trainer over a group stream).
    The law is SgnsTrainer's shared-negative law whatever n_parts is: one negative set per centre
    occurrence -- the same Philox draws, mapped into the centre's part -- and the sentence's job alpha
    (example_base / total_examples / sent_per_job as in SgnsTrainer.train). neg_group = G > 1 shares
    a set among the centres at G consecutive positions of a walk (faster at many parts, a different
    estimator: scripts/auc_block.py)."""

    def __init__(self, counts_by_id, *args, local_parts: int = 0, neg_group: int = 1, **kw):
        from . import dist as D
        self._rank, self._world = D.world()
        self._local_parts = int(local_parts)
        self.n_parts = self._local_parts if self._local_parts else self._world
        if self.n_parts not in (1, 2, 4, 8):
            raise ValueError("the tables can be cut into 1, 2, 4 or 8 parts")
        self.neg_group = int(neg_group)
        if not 1 <= self.neg_group <= 256:
            raise ValueError("neg_group must be in [1, 256]")
        self._buf = {}
        self.phase_events = None          # set to [] to record (phase, start event, end event) per train()
        super().__init__(counts_by_id, *args, **kw)
        if self.V < self.n_parts:
            raise ValueError("fewer vocabulary rows than parts")
        if self.dim % 4 or self.dim > 128:
            raise ValueError("block-partitioned tables need dim to be a multiple of 4, <= 128")
        if self.negative != 5:
            raise ValueError("block-partitioned tables run the shared-negative law with negative = 5")

    @property
    def _mine(self):
        return range(self.n_parts) if self._local_parts else [self._rank]

    def reset_weights(self):
        dev = self.counts.device
        W = self.n_parts
        rows = (self.V + W - 1) // W                    # same shape for every part: ring buffers interchange
        self.parts0, self.parts1 = {}, {}
        for k in self._mine:
            self.parts0[k] = torch.zeros((rows, self.dim), dtype=torch.float32, device=dev)
            self.parts1[k] = torch.zeros((rows, self.dim), dtype=torch.float32, device=dev)
            check(lib().n2v_sgns_init_part(ptr(self.parts0[k]), ptr(self.parts1[k]), C.c_int32(self.V), C.c_int32(self.dim),
                                           C.c_uint64(self.seed), C.c_int32(k), C.c_int32(W), stream()))
        self._spare = torch.empty_like(self.parts0[self._rank]) if (not self._local_parts and W > 1) else None
        self.syn0 = self.syn1neg = None

    def default_hogwild_warps(self, shared: bool = True) -> int:
        sms = int(lib().n2v_sm_count())
        return int(max(4, min(sms * 20, (self.V // self.n_parts) // 4)))

    def _params(self, epoch, grid_warps, *, total_examples=1, example_base=0, sent_per_job=1, alpha=None,
                min_alpha=None, hot_rows=None):
        P = SgnsParams()
        P.V, P.dim, P.window, P.negative = self.V, self.dim, self.window, self.negative
        P.bucket_bits, P.max_sentence_len = self.bucket_bits, 10000
        P.alpha0 = self.alpha if alpha is None else float(alpha)
        P.min_alpha = self.min_alpha if min_alpha is None else float(min_alpha)
        P.total_examples, P.example_base = int(total_examples), int(example_base)
        P.sent_per_job = max(1, int(sent_per_job))
        P.epoch, P.seed = int(epoch), self.seed
        P.grid_warps = int(grid_warps or self.default_hogwild_warps())
        P.atomic_updates, P.negative_sharing, P.tuning = 1, 1, 0
        P.hot_rows = self.hot_rows(P.grid_warps) if hot_rows is None else int(hot_rows)
        return P

    def make_groups(self, tokens, sent_off, n_sent, stride, sent_id_base, P, part, exact_bounds=True):
        """-> (words uint32[capacity], bounds | None): stream b of `part` = words[bounds[b]:bounds[b + 1]].
        exact_bounds=False skips the host read of the stream bounds when the buffer already has 1.5x
        the room the previous pool needed (the kernels then read the bounds on the device and clamp to
        the capacity; check_overflow() tells whether a pool ever did not fit)."""
        dev, W = self.counts.device, self.n_parts
        b = self._buf.setdefault(part, {})
        n_off = W * n_sent + 1
        if b.get("n_off", 0) < n_off:
            b["offsets"] = torch.empty(n_off, dtype=torch.int64, device=dev)
            b["ws"] = torch.empty(int(lib().n2v_sgns_groups_workspace_bytes(C.c_int64(n_sent), C.c_int32(W))),
                                  dtype=torch.uint8, device=dev)
            b["n_off"] = n_off
            b["overflow"] = torch.zeros(1, dtype=torch.int64, device=dev)
            b["need"] = None
        keep = ptr(self.keep_thr if self.sample > 0 else None)
        head = (ptr(tokens), ptr(sent_off), C.c_int64(n_sent), C.c_int32(stride), C.c_int64(sent_id_base),
                ptr(self.vocab_of_id), keep, C.byref(P), C.c_int32(part), C.c_int32(W))
        check(lib().n2v_sgns_groups_count(*head, ptr(b["offsets"]), ptr(b["ws"]), C.c_size_t(b["ws"].numel()), stream()))
        bounds = None
        lazy = (not exact_bounds and b.get("need") is not None and b.get("need_n") == n_sent
                and b.get("cap", -1) >= int(b["need"] * 1.5))
        if not lazy:
            bounds = [int(x) for x in b["offsets"][: n_off : n_sent].cpu().tolist()]     # one small D2H read
            total = bounds[-1]
            b["need"], b["need_n"] = total, n_sent
            want = int(total * (1.1 if exact_bounds else 1.6)) + 1024
            if b.get("cap", -1) < (total if exact_bounds else int(total * 1.5)):
                b["cap"] = want
                b["words"] = torch.empty(want, dtype=torch.int32, device=dev)
        check(lib().n2v_sgns_groups_fill(*head, ptr(self.cum_table), ptr(self.bucket_lo), C.c_int32(self.neg_group),
                                         ptr(b["offsets"]), ptr(b["words"]), C.c_int64(b["cap"]), ptr(b["overflow"]),
                                         stream()))
        return b["words"], bounds

    def train_bucket(self, part, bucket, syn0_part, P, n_sent, sent_id_base, bounds=None):
        b = self._buf[part]
        if bounds is not None:
            first, n = bounds[bucket], bounds[bucket + 1] - bounds[bucket]
            if n <= 0:
                return
            d0 = d1 = C.c_void_p(0)
        else:                      # bounds stay on the device: two entries of the offsets array
            first, n = 0, 0
            base = b["offsets"].data_ptr()
            d0, d1 = C.c_void_p(base + 8 * bucket * n_sent), C.c_void_p(base + 8 * (bucket + 1) * n_sent)
        check(lib().n2v_sgns_train_groups(ptr(b["words"]), C.c_int64(first), C.c_int64(n), d0, d1, C.c_int64(b["cap"]),
                                          C.byref(P), C.c_int32(self.neg_group), ptr(syn0_part), ptr(self.parts1[part]),
                                          C.c_int32(part), C.c_int32(self.n_parts), ptr(self.pairs), stream()))

    def train(self, tokens, sent_off, n_sent, stride, *, total_examples, example_base=0, sent_id_base=0,
              epoch=0, sent_per_job=125, grid_warps=None, alpha=None, min_alpha=None, exact_bounds=True, hot_rows=None,
              **_ignored):
        """One pool. Multi-GPU: every rank passes its own walks (fixed-stride buffer, same n_sent on
        every rank) and the POOL's example_base / sent_id_base (identical on all ranks)."""
        from . import dist as D
        W = self.n_parts
        multi = not self._local_parts and W > 1
        mark = self._mark
        t = mark()
        if multi:
            if sent_off is not None:
                raise NotImplementedError("multi-GPU pools are fixed-stride walk buffers")
            tokens = D.gather_pool(tokens[:n_sent])
            n_sent = n_sent * W
        if n_sent <= 0:
            return
        t = mark("gather", t)
        P = self._params(epoch, grid_warps, total_examples=total_examples, example_base=example_base,
                         sent_per_job=sent_per_job, alpha=alpha, min_alpha=min_alpha, hot_rows=hot_rows)
        streams = {k: self.make_groups(tokens, sent_off, n_sent, stride, sent_id_base, P, k, exact_bounds)
                   for k in self._mine}
        t = mark("pairs", t)
        held = None if self._local_parts else self.parts0[self._rank]
        for e in range(W):
            for k in self._mine:
                b = (k + e) % W
                self.train_bucket(k, b, self.parts0[b] if self._local_parts else held, P, n_sent, sent_id_base,
                                  streams[k][1])
            t = mark("train", t)
            if multi:                                    # the syn0 part moves on; after W passes it is home again
                held, self._spare = D.ring_pass(held, self._spare)
                t = mark("ring", t)
        if multi:
            self.parts0[self._rank] = held

    def _mark(self, phase=None, since=None):
        if self.phase_events is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if phase is not None:
            self.phase_events.append((phase, since, e))
        return e

    def check_overflow(self):
        for b in self._buf.values():
            if int(b["overflow"].item()):
                raise N2VError("group buffer overflow: a pool did not fit the stream buffer")

    def gather(self):
        """-> (syn0, syn1neg) float32[V, dim] on this device, rows in vocabulary order"""
        import torch.distributed as tdist
        W, dev = self.n_parts, self.counts.device
        out = []
        for parts in (self.parts0, self.parts1):
            full = torch.empty((self.V, self.dim), dtype=torch.float32, device=dev)
            if self._local_parts or W == 1:
                got = [parts[k] for k in sorted(parts)]
            else:
                got = [torch.empty_like(parts[self._rank]) for _ in range(W)]
                tdist.all_gather(got, parts[self._rank].contiguous())
            for k in range(W):
                rows = (self.V - k + W - 1) // W
                full[k::W] = got[k][:rows]
            out.append(full)
        return out[0], out[1]


class Word2Vec:
    """gensim.models.Word2Vec(sg=1, hs=0, negative=k) on the GPU. Signature and defaults are
    gensim 3.2.0's; the reference passes size, window, min_count=0, sg=1, workers, iter."""

    def __init__(self, sentences=None, size=100, alpha=0.025, window=5, min_count=5,
                 max_vocab_size=None, sample=1e-3, seed=1, workers=3, min_alpha=0.0001,
                 sg=0, hs=0, negative=5, cbow_mean=1, hashfxn=hash, iter=5, null_word=0,
                 trim_rule=None, sorted_vocab=1, batch_words=10000, compute_loss=False,
                 *, hogwild_warps=None, atomic_updates=None, shared_negatives=None):
        if not sg or hs or negative <= 0:
            raise NotImplementedError("only skip-gram with negative sampling (sg=1, hs=0, negative>0) "
                                      "is implemented: it is the only mode the reference uses")
        self.vector_size = self.layer1_size = int(size)
        self.alpha, self.min_alpha = float(alpha), float(min_alpha)
        self.window, self.min_count, self.sample = int(window), int(min_count), float(sample)
        self.seed, self.workers, self.negative = int(seed), int(workers), int(negative)
        self.iter, self.batch_words = int(iter), int(batch_words)
        self.sg, self.hs = 1, 0
        self.hogwild_warps = hogwild_warps
        # row updates: 1 = red.global.add.v4.f32 (no lost updates; measured to keep link-prediction
        # AUC within 0.003 of the CPU oracle at every Hogwild width), 0 = plain racy stores
        self.atomic_updates = (int(os.environ.get("N2V_SGNS_ATOMIC", "1")) if atomic_updates is None
                               else int(atomic_updates))
        # 1 (default) = one negative set per centre shared by its context pairs (2.4x faster, AUC
        # within 0.002 of the CPU oracle, scripts/auc_sweep.py); 0 = gensim's fresh set per pair
        self.shared_negatives = (int(os.environ.get("N2V_SGNS_SHARED", "1")) if shared_negatives is None
                                 else int(shared_negatives))
        self.wv = KeyedVectors(self.vector_size)
        self.corpus_count = 0
        self.train_count = 0
        self._pairs_done, self._pairs_pending, self._shard = 0, None, None
        self.syn1neg = None
        if sentences is not None:
            self.build_vocab(sentences)
            self.train(self._corpus, total_examples=self.corpus_count, epochs=self.iter)

    # gensim forwards these to .wv
    def __getitem__(self, w):
        return self.wv[w]

    def __contains__(self, w):
        return w in self.wv

    def similarity(self, a, b):
        return self.wv.similarity(a, b)

    # ---- corpus ingest -----------------------------------------------------------------------
    def _ingest(self, sentences):
        """-> (tokens int32 device [n_tok] or [n_sent, stride], sent_off device|None, stride,
        n_sent, words list indexed by token id)"""
        dev = require_cuda()
        func = str
        self._shard = None
        if not isinstance(sentences, WalkCorpus):
            # `[map(str, walk) for walk in walks]` (main.py:86) over a WalkCorpus: the rows are lazy, the map
            # objects expose them (walker.rows_of) -- train on the device corpus, no Python strings at all
            r = rows_of(sentences)
            if r is not None and (r[2] is None or r[2] is str):
                corpus, idx, func = r
                if idx == list(range(len(corpus))):
                    sentences = corpus
                else:
                    sel = torch.as_tensor(idx, dtype=torch.int64, device=corpus.walks.device)
                    sentences = WalkCorpus(corpus.walks[sel].contiguous(), corpus.lens[sel].contiguous(), corpus.labels,
                                           corpus.shard)
                    sentences.n_ids = corpus.n_ids
        if isinstance(sentences, WalkCorpus):     # already on the device: tokens are compact node ids
            self._shard = sentences.shard
            labels = sentences.labels
            key = (id(labels), func) if labels is not None else None
            if key is not None and getattr(self, "_words_key", None) == key:
                words = self._words_cache           # same graph as last time: the label list is reused
            else:
                n_ids = int(len(labels)) if labels is not None else (
                    int(sentences.n_ids) if getattr(sentences, "n_ids", None) else int(sentences.walks.max().item()) + 1)
                key = ("ids", n_ids, func) if labels is None else key
                if getattr(self, "_words_key", None) == key:
                    return sentences.walks, None, int(sentences.walks.shape[1]), int(sentences.walks.shape[0]), self._words_cache
                src = labels.tolist() if labels is not None else range(n_ids)
                words = [str(l) for l in src] if func is str else list(src)
                self._words_key, self._words_cache, self._words_labels = key, words, labels
            return sentences.walks, None, int(sentences.walks.shape[1]), int(sentences.walks.shape[0]), words
        if isinstance(sentences, LineSentence) and isinstance(sentences.source, (str, os.PathLike)) \
                and sentences.limit is None:
            fast = self._ingest_walk_file(sentences, dev)
            if fast is not None:
                return fast
        # generic iterable of iterables of tokens (materialised once: py3 `map` objects are one-shot,
        # gensim needs >= 2 passes -- SURVEY.md section 2)
        ids = {}
        toks, offs = [], [0]
        for sent in sentences:
            for wd in sent:
                i = ids.get(wd)
                if i is None:
                    i = ids[wd] = len(ids)
                toks.append(i)
            offs.append(len(toks))
        words = [None] * len(ids)
        for wd, i in ids.items():
            words[i] = wd
        tok = torch.as_tensor(np.asarray(toks, dtype=np.int32)).to(dev)
        off = torch.as_tensor(np.asarray(offs, dtype=np.int64)).to(dev)
        return tok, off, 0, len(offs) - 1, words

    def _ingest_walk_file(self, ls, dev):
        """LineSentence over a file of integer tokens (a walk file): tokenised on the device
        (n2v_parse_walks_*). Returns None when a token is not an integer (generic path then)."""
        L = lib()
        raw = np.fromfile(ls.source, dtype=np.uint8)
        n = int(raw.shape[0])
        if n == 0 or n > (1 << 31):
            return None
        text = torch.as_tensor(raw).to(dev)
        tf = torch.empty(n + 1, dtype=torch.int32, device=dev); lf = torch.empty(n + 1, dtype=torch.int32, device=dev)
        ti = torch.empty(n + 1, dtype=torch.int64, device=dev); li = torch.empty(n + 1, dtype=torch.int64, device=dev)
        ws_bytes = int(L.n2v_parse_workspace_bytes(C.c_int64(n)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(L.n2v_parse_walks_index(ptr(text), C.c_int64(n), ptr(tf), ptr(lf), ptr(ti), ptr(li), ptr(ws),
                                      C.c_size_t(ws_bytes), stream()))
        n_tok, n_lines = int(ti[-1].item()), int(li[-1].item())
        labels = torch.empty(max(n_tok, 1), dtype=torch.int64, device=dev)
        off = torch.empty(n_lines + 1, dtype=torch.int64, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        check(L.n2v_parse_walks_fill(ptr(text), C.c_int64(n), ptr(tf), ptr(lf), ptr(ti), ptr(li), ptr(labels),
                                     ptr(off), ptr(bad), stream()))
        if int(bad.item()):
            return None
        off[n_lines] = n_tok
        labels = labels[:n_tok]
        lens = off[1:] - off[:-1]
        if int(lens.max().item()) > ls.max_sentence_length:      # LineSentence splits long lines
            return None
        off = torch.cat([off[:1], off[1:][lens > 0]])           # empty lines yield no sentence
        # ids in first-seen order, exactly like the generic path
        uniq, inv = torch.unique(labels, return_inverse=True)
        first = torch.full((uniq.numel(),), n_tok, dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, inv, torch.arange(n_tok, device=dev), reduce="amin")
        order = torch.argsort(first)
        id_of_uniq = torch.empty_like(order)
        id_of_uniq[order] = torch.arange(order.numel(), device=dev)
        tok = id_of_uniq[inv].to(torch.int32).contiguous()
        words = [str(v) for v in uniq[order].tolist()]
        return tok, off.contiguous(), 0, int(off.numel() - 1), words

    def build_vocab(self, sentences, **_):
        """scan_vocab + scale_vocab + finalize_vocab (word2vec.py): count, drop < min_count, sort by
        count descending, sub-sampling thresholds, cum_table, reset_weights."""
        dev = require_cuda()
        tok, off, stride, n_sent, words = self._ingest(sentences)
        self._corpus = (tok, off, stride, n_sent)
        self._vocab_words = words                 # token id -> word of the corpus the vocabulary was built from
        self.corpus_count = n_sent
        n_ids = len(words)
        counts = torch.zeros(max(n_ids, 1), dtype=torch.int64, device=dev)
        check(lib().n2v_vocab_count(ptr(tok), C.c_int64(tok.numel()), C.c_int32(n_ids), ptr(counts), stream()))
        shard = getattr(self, "_shard", None)
        first_seen = None
        if off is None and stride > 0:            # a walk buffer: token ids are node ids, not first-seen ranks
            flat = tok.reshape(-1)
            pos = torch.nonzero(flat >= 0).reshape(-1)
            ids = flat[pos].to(torch.int64)
            if shard is not None:                 # positions in the whole corpus: this rank's share starts here
                pos = pos + int(shard[0]) * (-(-int(shard[2]) // int(shard[1]))) * stride
            first_seen = torch.full((max(n_ids, 1),), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
            first_seen.scatter_reduce_(0, ids, pos, reduce="amin")
            del flat, pos, ids
        if shard is not None:
            # one rank's share of a corpus simulated by `world` processes: global counts, tables cut into
            # `world` row sets, block-partitioned training over an NCCL ring (BlockSgnsTrainer)
            from . import dist as D
            if off is not None or self.vector_size > 128 or self.vector_size % 4 or self.negative != 5:
                raise NotImplementedError("multi-GPU training needs a walk corpus, size <= 128 (multiple of 4), negative = 5")
            D.sum_counts(counts)
            import torch.distributed as tdist
            tdist.all_reduce(first_seen, op=tdist.ReduceOp.MIN)
            self.corpus_count = int(shard[2])
            self.trainer = T = BlockSgnsTrainer(counts[:n_ids], dim=self.vector_size, window=self.window,
                                                negative=self.negative, sample=self.sample, seed=self.seed,
                                                alpha=self.alpha, min_alpha=self.min_alpha, min_count=self.min_count,
                                                batch_words=self.batch_words, first_seen=first_seen[:n_ids])
        else:
            self.trainer = T = SgnsTrainer(counts[:n_ids], dim=self.vector_size, window=self.window,
                                           negative=self.negative, sample=self.sample, seed=self.seed,
                                           alpha=self.alpha, min_alpha=self.min_alpha, min_count=self.min_count,
                                           batch_words=self.batch_words,
                                           first_seen=None if first_seen is None else first_seen[:n_ids])
        # host-side vocabulary objects (what emb.vocab / index2word expose)
        order_h, vc_h = T.order.cpu().numpy(), T.counts.cpu().numpy()
        kt_h = T.keep_thr.cpu().numpy().view(np.uint32)
        self.wv.index2word = [words[i] for i in order_h]
        self.wv.vocab = {w: Vocab(i, int(vc_h[i]), int(kt_h[i])) for i, w in enumerate(self.wv.index2word)}
        self.wv.vector_size = self.vector_size
        self.wv._syn0_dev, self.wv._syn0_host = T.syn0, None

    def _train_sharded(self, tok, n_sent, stride, epochs, start_alpha, end_alpha):
        """every rank's share of the corpus, pool by pool (pool = the same slice of every rank's share;
        shares are padded with empty walks to one length so that all ranks run the same pools)"""
        from . import dist as D
        T = self.trainer
        rank, world, total = self._shard
        per, pools = D.pool_plan(total, world, max(1024, int(os.environ.get("N2V_POOL_WALKS", str(1 << 19)))))
        if n_sent < per:
            tok = torch.cat([tok, torch.full((per - n_sent, stride), -1, dtype=tok.dtype, device=tok.device)])
        mean_len = max(1.0, T.raw_words / max(self.corpus_count, 1))
        before = T.pairs[0].clone()
        for ep in range(epochs):
            for p0, n in pools:
                T.train(tok[p0:p0 + n].contiguous(), None, n, stride, total_examples=max(1, per * world * epochs),
                        example_base=(ep * per + p0) * world, sent_id_base=(ep * per + p0) * world, epoch=ep,
                        sent_per_job=int(self.batch_words // mean_len), grid_warps=self.hogwild_warps,
                        alpha=start_alpha, min_alpha=end_alpha, exact_bounds=False)
        self._overflow_check = T.check_overflow           # read with the pair count (no device sync per call)
        import torch.distributed as tdist
        mine = T.pairs[0] - before
        tdist.all_reduce(mine)
        self._pairs_pending = mine if self._pairs_pending is None else self._pairs_pending + mine
        self.wv._syn0_dev, self.wv._syn0_host = None, None        # all-gathered when the table is first read
        self.wv._gather = lambda: T.gather()[0]

    # test/inspection handles
    @property
    def _keep_thr(self):
        return self.trainer.keep_thr

    @property
    def _cum_table(self):
        return self.trainer.cum_table

    @property
    def _bucket_lo(self):
        return self.trainer.bucket_lo

    @property
    def _bucket_bits(self):
        return self.trainer.bucket_bits

    @property
    def syn1neg_dev(self):
        return self.trainer.syn1neg

    def reset_weights(self):
        self.trainer.reset_weights()
        self.wv._syn0_dev, self.wv._syn0_host = self.trainer.syn0, None

    # ---- training ------------------------------------------------------------------------------
    def train(self, sentences=None, total_examples=None, total_words=None, epochs=None,
              start_alpha=None, end_alpha=None, **_):
        vmap = None
        if sentences is None or isinstance(sentences, tuple):
            tok, off, stride, n_sent = self._corpus if sentences is None else sentences
        else:
            # gensim's model.train(new_sentences, total_examples=, epochs=) on the existing vocabulary:
            # words the vocabulary does not hold are ignored, as gensim ignores them
            if getattr(self, "trainer", None) is None:
                raise RuntimeError("you must first build vocabulary before training the model")
            tok, off, stride, n_sent, words = self._ingest(sentences)
            if words is not self._vocab_words:
                v = self.wv.vocab
                vmap = torch.as_tensor(np.fromiter((v[w].index if w in v else -1 for w in words), dtype=np.int32,
                                                   count=len(words))).to(tok.device)
        T = self.trainer
        epochs = self.iter if epochs is None else int(epochs)
        if getattr(self, "_shard", None) is not None:
            self.train_count += 1
            return self._train_sharded(tok, n_sent, stride, epochs, start_alpha, end_alpha)
        n_ex = int(n_sent) if total_examples is None else int(total_examples)
        mean_len = max(1.0, T.raw_words / max(self.corpus_count, 1))
        self.train_count += 1
        before = T.pairs[0].clone()
        for ep in range(epochs):
            T.train(tok, off, n_sent, stride, total_examples=max(1, n_ex * epochs), example_base=ep * n_ex,
                    epoch=ep + 1000 * (self.train_count - 1), sent_per_job=int(self.batch_words // mean_len),
                    grid_warps=self.hogwild_warps, atomic_updates=self.atomic_updates, alpha=start_alpha,
                    min_alpha=end_alpha, vocab_of_id=vmap,
                    negative_sharing=self.shared_negatives if (self.vector_size <= 128 and self.negative == 5) else 0)
        self._pairs_pending = (T.pairs[0] - before) if getattr(self, "_pairs_pending", None) is None \
            else self._pairs_pending + (T.pairs[0] - before)
        self.wv._syn0_dev, self.wv._syn0_host = T.syn0, None

    @property
    def pairs_trained(self):
        """(centre, context) pairs trained so far; reading it waits for the device"""
        if getattr(self, "_pairs_pending", None) is not None:
            self._pairs_done += int(self._pairs_pending.item())
            self._pairs_pending = None
        if getattr(self, "_overflow_check", None) is not None:
            self._overflow_check()
            self._overflow_check = None
        return self._pairs_done

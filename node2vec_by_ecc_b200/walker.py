"""Drop-in for the reference's ``node2vec`` module (src/node2vec.py): same class, method and
function names, argument meaning, walk order and return shapes; the work runs in
libn2v_b200.so on the current CUDA device. No CPU fallback.

    import node2vec                      # node2vec_by_ecc_b200/dropin/node2vec.py re-exports this
    G = node2vec.Graph(nx_G, is_directed, p, q)
    G.preprocess_transition_probs()
    walks = G.simulate_walks(num_walks, walk_length)

Extra, optional knobs (keyword-only or environment, so existing calls stay valid):
  seed            Philox key (default: N2V_SEED env or 0); each simulate_* call advances a walk-id
                  base so repeated calls give fresh walks, like the reference's global numpy RNG.
  mode            "auto" | "alias" | "reject": alias = precomputed edge tables (bit-exact to the
                  reference under injected uniforms), reject = rejection sampling (same law, no
                  edge tables); auto picks alias when the tables fit N2V_TABLE_BUDGET_GB.
"""
from __future__ import annotations

import gc
import itertools
import os
from collections.abc import Sequence

import numpy as np
import torch

from .graph import AliasTables, DeviceGraph


class _RowIter:
    """iterator of one walk of a WalkCorpus. Nothing is copied from the device until the first
    __next__; `[map(str, walk) for walk in walks]` (main.py:86) therefore costs one small object per
    walk, and Word2Vec recognises such sentences (map.__reduce__ exposes this iterator) and trains
    on the device-resident corpus instead of 80 Python strings per walk."""
    __slots__ = ("corpus", "index", "_it")

    def __init__(self, corpus, index):
        self.corpus, self.index, self._it = corpus, index, None

    def __iter__(self):
        return self

    def __next__(self):
        if self._it is None:
            self._it = iter(self.corpus[self.index])
        return next(self._it)

    @property
    def untouched(self):
        return self._it is None


class WalkRow(Sequence):
    """One walk of a WalkCorpus, as iteration over the corpus yields it: behaves like the list of
    node labels the reference's simulate_walks returns (len, indexing, iteration, == with a list),
    materialised from the device buffer only when its tokens are actually read."""
    __slots__ = ("corpus", "index")

    def __init__(self, corpus, index):
        self.corpus, self.index = corpus, index

    def tolist(self):
        return self.corpus[self.index]

    def __len__(self):
        return int(self.corpus._h()[1][self.index])

    def __getitem__(self, k):
        return self.tolist()[k]

    def __iter__(self):
        return _RowIter(self.corpus, self.index)

    def __eq__(self, other):
        if isinstance(other, WalkRow):
            other = other.tolist()
        return self.tolist() == other

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return repr(self.tolist())

    def __add__(self, other):
        return self.tolist() + list(other)

    def __radd__(self, other):
        return list(other) + self.tolist()


class WalkCorpus(Sequence):
    """The list-of-lists simulate_walks returns, kept on the device: int32[n_walks, L] compact
    ids padded with -1. Behaves like a list of lists of ORIGINAL node labels (lazy host copy on
    first Python access); Word2Vec consumes it without leaving the GPU."""

    def __init__(self, walks: torch.Tensor, lens: torch.Tensor, labels, shard=None):
        self.walks, self.lens, self.labels = walks, lens, labels
        # (rank, world, total walks) when this object holds one rank's contiguous share of a corpus that
        # was simulated by `world` processes (Graph(..., distributed=True)); Word2Vec then trains with
        # the block-partitioned multi-GPU trainer
        self.shard = shard
        self.n_ids = None                 # number of node ids when there are no labels (tokens are compact ids)
        self._host = None

    # -- device side
    @property
    def walk_length(self):
        return self.walks.shape[1]

    def num_steps(self) -> int:
        return int((self.lens.to(torch.int64) - 1).clamp_(min=0).sum().item())

    def extend(self, other):
        if not isinstance(other, WalkCorpus):
            raise TypeError("WalkCorpus.extend needs another WalkCorpus")
        if other.walks.shape[1] != self.walks.shape[1]:
            raise ValueError("walk_length differs")
        self.walks = torch.cat([self.walks, other.walks])
        self.lens = torch.cat([self.lens, other.lens])
        self._host = None

    def __add__(self, other):
        out = WalkCorpus(self.walks, self.lens, self.labels)
        out.extend(other)
        return out

    def format_walks(self, first=0, last=None) -> torch.Tensor:
        """Walks [first, last) as the bytes of the walk file (uint8 device tensor), formatted on the
        device (n2v_format_walks_*); integer labels only."""
        import ctypes as C
        from ._lib import check, lib, ptr, stream
        last = len(self) if last is None else last
        w, l = self.walks[first:last].contiguous(), self.lens[first:last].contiguous()
        n, L = int(w.shape[0]), int(w.shape[1])
        dev = w.device
        lab = None
        if self.labels is not None:
            if self.labels.dtype == object or not np.issubdtype(self.labels.dtype, np.integer):
                raise TypeError("device formatting needs integer node labels")
            lab = torch.as_tensor(self.labels.astype(np.int64)).to(dev)
        Lb = lib()
        ws_bytes = int(Lb.n2v_format_workspace_bytes(C.c_int64(n), C.c_int32(L)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        off = torch.empty(n * L + 1, dtype=torch.int64, device=dev)
        check(Lb.n2v_format_walks_offsets(ptr(w), ptr(l), C.c_int64(n), C.c_int32(L), ptr(lab), ptr(off), ptr(ws),
                                          C.c_size_t(ws_bytes), stream()))
        total = int(off[-1].item())
        out = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
        check(Lb.n2v_format_walks_write(ptr(w), ptr(l), C.c_int64(n), C.c_int32(L), ptr(lab), ptr(off), ptr(out),
                                        stream()))
        return out[:total]

    def save_walks(self, path, mode="w", chunk_walks=1 << 22):
        """The walk file of main_link.py:237-239,544-546: one walk per line, tokens (original
        labels) joined by single spaces -- what LineSentence / `-walk-path` read back. Integer
        labels are formatted on the device in chunks of `chunk_walks`; other labels on the host."""
        lab = self.labels
        if self.walks.is_cuda and (lab is None or (lab.dtype != object and np.issubdtype(lab.dtype, np.integer))):
            with open(path, mode + "b") as f:
                for a in range(0, len(self), chunk_walks):
                    f.write(self.format_walks(a, min(len(self), a + chunk_walks)).cpu().numpy().tobytes())
            return
        w, l = self._h()
        with open(path, mode) as f:
            for i in range(w.shape[0]):
                row = w[i, :l[i]]
                f.write(" ".join(map(str, row.tolist() if lab is None else lab[row].tolist())) + "\n")

    # -- list-of-lists view
    def _h(self):
        if self._host is None:
            self._host = (self.walks.cpu().numpy(), self.lens.cpu().numpy())
        return self._host

    def __len__(self):
        return self.walks.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        w, l = self._h()
        if i < 0:
            i += len(self)
        row = w[i, :l[i]]
        if self.labels is None:
            return row.tolist()
        return self.labels[row].tolist()

    def __iter__(self):
        # One small WalkRow per walk. A consumer that keeps them all (`[map(str, walk) for walk in
        # walks]`, main.py:86) would make CPython's cyclic collector rescan a growing heap of live
        # objects every few hundred allocations (measured: 4x the loop's own time at 5e5 walks), so
        # collection is paused while the corpus is being iterated and restored when the loop ends.
        was = gc.isenabled()
        gc.disable()
        try:
            yield from map(WalkRow, itertools.repeat(self), range(len(self)))
        finally:
            if was:
                gc.enable()


def rows_of(sentences):
    """If `sentences` is a list of walks of ONE WalkCorpus -- WalkRow objects, or untouched
    map(str, row) / map(<f>, row) iterators over them, in any order -- returns (corpus, row indices,
    the mapped function or None); otherwise None. Nothing is consumed."""
    if not isinstance(sentences, (list, tuple)) or not sentences:
        return None
    corpus, func, idx = None, None, []
    for k, sent in enumerate(sentences):
        f = None
        if type(sent) is map:
            try:
                _, args = sent.__reduce__()
            except Exception:
                return None
            if len(args) != 2:
                return None
            f, sent = args
        if type(sent) is WalkRow:
            c, i = sent.corpus, sent.index
        elif type(sent) is _RowIter and sent.untouched:
            c, i = sent.corpus, sent.index
        else:
            return None
        if k == 0:
            corpus, func = c, f
        elif c is not corpus or f is not func:
            return None
        idx.append(i)
    return corpus, idx, func


def alias_setup(probs):
    """node2vec.py:240-269 on the device for one distribution -> (J int64[K], q float64[K])."""
    pr = np.asarray(probs, dtype=np.float64)
    K = pr.shape[0]
    if K == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float64)
    # a star graph whose centre row carries `probs` as weights, built without normalisation
    row_ptr = np.concatenate([[0], [K], np.full(K, K)]).astype(np.int64)
    g = DeviceGraph.from_csr(row_ptr, np.arange(1, K + 1, dtype=np.int32), pr, symmetric=False)
    t = g.build_node_tables(keep_raw=True, raw_probs=True)
    return t.node_J.cpu().numpy().astype(np.int64), t.node_q.cpu().numpy()


def alias_draw(J, q):
    """node2vec.py:271-281 (host helper, numpy global RNG like the reference)."""
    K = len(J)
    kk = int(np.floor(np.random.rand() * K))
    if np.random.rand() < q[kk]:
        return kk
    return J[kk]


class _Identity:
    """label -> compact id map of a graph whose labels ARE its compact ids"""

    def __init__(self, n):
        self.n = n

    def __getitem__(self, k):
        k = int(k)
        if not 0 <= k < self.n:
            raise KeyError(k)
        return k


class _TableView:
    """dict-like read-only view of device alias tables in the reference's (J, q) form."""

    def __init__(self, graph: "Graph", edges: bool):
        self._g, self._edges = graph, edges

    def _raw(self):
        return self._g._raw_tables()

    def __getitem__(self, key):
        dg, idx = self._g._dg, self._g._index
        t = self._raw()
        if not self._edges:
            v = idx[key]
            a, b = int(dg.row_ptr[v]), int(dg.row_ptr[v + 1])
            return t.node_J[a:b].cpu().numpy().astype(np.int64), t.node_q[a:b].cpu().numpy()
        u, v = idx[key[0]], idx[key[1]]
        a, b = int(dg.row_ptr[u]), int(dg.row_ptr[u + 1])
        pos = int(torch.searchsorted(dg.col[a:b].contiguous(), torch.tensor([v], dtype=torch.int32, device=dg.device)))
        if pos >= b - a or int(dg.col[a + pos]) != v:
            raise KeyError(key)
        e = a + pos
        o, o2 = int(t.etab_ptr[e]), int(t.etab_ptr[e + 1])
        return t.edge_J[o:o2].cpu().numpy().astype(np.int64), t.edge_q[o:o2].cpu().numpy()

    def __len__(self):
        return self._g._dg.nnz if self._edges else self._g._dg.n


class Graph:
    """Same constructor and methods as the reference class (node2vec.py:5-237)."""

    def __init__(self, nx_G, is_directed, p, q, popwalk="none", *, seed=None, mode="auto", distributed=None):
        # distributed (or N2V_DISTRIBUTED=1) with torch.distributed initialised: every process builds the
        # same graph on its own GPU and simulate_walks returns THIS rank's contiguous share of the walks
        # (the partitioning main_link.py:263-264 applies across its pool; global walk ids, so the corpus
        # does not depend on the number of GPUs)
        self.distributed = (os.environ.get("N2V_DISTRIBUTED", "0") == "1") if distributed is None else bool(distributed)
        self.G = nx_G
        self.is_directed = is_directed
        self.p = p
        self.q = q
        self.popwalk = popwalk
        self.seed = int(os.environ.get("N2V_SEED", "0")) if seed is None else int(seed)
        self.mode = os.environ.get("N2V_WALK_MODE", mode)
        self._reset()

    def _reset(self):
        self._dg_obj = None
        self._dg_for = None
        self._tables = None
        self._tables_raw = None
        self._prep = None                 # which preprocess_* was called: None | "plain" | "pop" (picklable)
        self._otf_key = None
        self._walk_id_base = 0

    # Graph instances are pickled to pool workers by main_link.py:277 -- after preprocess_* ran in the
    # parent (:216-226). Device handles stay behind; the worker rebuilds the tables it needs lazily
    # from `_prep`, and takes a walk-id range of its own (the reference's workers all inherit one numpy
    # RNG state; here equal ids would mean equal Philox streams).
    def __getstate__(self):
        st = dict(self.__dict__)
        for k in ("_dg_obj", "_dg_for", "_tables", "_tables_raw", "alias_nodes", "alias_edges"):
            st.pop(k, None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._dg_obj = self._dg_for = self._tables = self._tables_raw = None
        self._otf_key = None
        self._walk_id_base += (os.getpid() & 0xFFFFF) << 40
        if self._prep is not None:
            self.alias_nodes = _TableView(self, edges=False)
            self.alias_edges = _TableView(self, edges=True)

    # ---- device graph, rebuilt when the caller swaps/mutates self.G (main_link.py:592) ----------
    @property
    def _dg(self) -> DeviceGraph:
        if isinstance(self.G, DeviceGraph):
            # a graph that never existed as a networkx object (100 M edges do not fit one): the CSR built
            # on the device by DeviceGraph.from_coo is taken as is; node labels are its compact ids
            if self._dg_obj is not self.G:
                self._dg_obj, self._dg_for = self.G, None
                self._tables = self._tables_raw = None
                self._otf_key = None
                self._index_map = None
                if self.G.order is None:
                    self.G.order = torch.arange(self.G.n, dtype=torch.int32, device=self.G.device)
            return self._dg_obj
        key = (id(self.G), self.G.number_of_nodes(), self.G.number_of_edges())
        if self._dg_obj is None or self._dg_for != key:
            self._dg_obj = DeviceGraph.from_networkx(self.G)
            self._dg_for = key
            self._tables = self._tables_raw = None
            self._otf_key = None
            self._index_map = {l: i for i, l in enumerate(self._dg_obj.labels.tolist())}
        return self._dg_obj

    @property
    def _index(self):
        self._dg
        if self._index_map is None:                     # DeviceGraph without labels: ids are the labels
            return _Identity(self._dg.n)
        return self._index_map

    def _table_budget(self) -> int:
        gb = float(os.environ.get("N2V_TABLE_BUDGET_GB", "0"))
        if gb > 0:
            return int(gb * 2 ** 30)
        free, _ = torch.cuda.mem_get_info()
        return int(free * 0.4)

    def _use_alias(self) -> bool:
        if self.mode == "alias":
            return True
        if self.mode == "reject":
            return False
        return self._dg.edge_table_bytes() * 2.5 <= self._table_budget()   # slots + build scratch

    def _build(self, popwalk_nodes: bool, edges: bool, keep_raw=False) -> AliasTables:
        dg = self._dg
        if edges:
            return dg.build_alias_tables(float(self.p), float(self.q), popwalk=popwalk_nodes, keep_raw=keep_raw)
        return dg.build_node_tables(popwalk=popwalk_nodes, keep_raw=keep_raw)

    def _raw_tables(self) -> AliasTables:
        if self._tables_raw is None:
            self._tables_raw = self._build(self._prep == "pop", edges=True, keep_raw=True)
        return self._tables_raw

    def _prepared(self) -> AliasTables:
        """the tables of the preprocess_* call on record (rebuilt after unpickling / a swap of self.G)"""
        self._dg
        if self._tables is None or self._otf_key is not None:
            if self._prep is None:
                raise AttributeError("'Graph' object has no attribute 'alias_nodes' "
                                     "(call preprocess_transition_probs() first)")
            self._tables = self._build(self._prep == "pop", edges=self._use_alias())
            self._otf_key = None
        return self._tables

    # ---- reference API ---------------------------------------------------------------------------
    def preprocess_transition_probs(self):
        """node2vec.py:176-204. Node tables always; edge tables when they fit (else the walks use
        the rejection sampler, which needs none)."""
        self._prep, self._tables, self._tables_raw = "plain", None, None
        self._prepared()
        self.alias_nodes = _TableView(self, edges=False)
        self.alias_edges = _TableView(self, edges=True)
        return

    def preprocess_transition_probs_popularity(self):
        """node2vec.py:206-237: popularity-normalised node tables (first step), plain edge law
        (:228-232) -- as edge tables when they fit, else by rejection."""
        self._prep, self._tables, self._tables_raw = "pop", None, None
        self._prepared()
        self.alias_nodes = _TableView(self, edges=False)
        self.alias_edges = _TableView(self, edges=True)
        return

    def _starts(self, num_walks, nodes):
        dg = self._dg
        if isinstance(nodes, (torch.Tensor, np.ndarray)):   # compact ids / integer labels in bulk
            ids = torch.as_tensor(nodes)
            if dg.labels is not None:
                lab = torch.as_tensor(np.asarray(dg.labels, dtype=np.int64))
                pos = torch.searchsorted(lab, ids.to(torch.int64).cpu())
                if bool((pos >= lab.numel()).any()) or not torch.equal(lab[pos.clamp_max(lab.numel() - 1)], ids.to(torch.int64).cpu()):
                    raise KeyError("start node not in the graph")
                ids = pos
            order = ids.to(device=dg.device, dtype=torch.int32, non_blocking=True)
        elif not nodes:                                 # `if not nodes` (node2vec.py:87)
            order = dg.order
        else:
            idx = self._index
            order = torch.as_tensor(np.fromiter((idx[x] for x in nodes), dtype=np.int32, count=len(nodes)),
                                    device=dg.device)
        return order.repeat(int(num_walks))             # walk_iter-major, then nodes (:89-93)

    def _simulate(self, num_walks, walk_length, nodes, tables: AliasTables, verbose):
        dg = self._dg
        starts = self._starts(num_walks, nodes)
        if verbose:
            for it in range(int(num_walks)):
                print(str(it + 1), '/', str(num_walks))
        base = self._walk_id_base
        self._walk_id_base += int(starts.shape[0])
        shard = None
        if self.distributed:
            from . import dist as D
            rank, world = D.world()
            if world > 1:
                total = int(starts.shape[0])
                lo, hi = D.shard_range(total, rank, world)
                starts, base, shard = starts[lo:hi].contiguous(), base + lo, (rank, world, total)
        if tables.edge_slots is not None:
            walks, lens = dg.walk_alias(tables, starts, int(walk_length), self.seed, base)
        else:
            # rejection mode. The candidate law of steps >= 2 is the PLAIN weight row (a node table on
            # weighted graphs, uniform otherwise); popularity node tables only ever drive the first step
            # (node2vec.py:69-70 with :213-218), unless the whole edge law is get_alias_edge_pop.
            pop_edges = bool(getattr(tables, "pop_edges", False))
            first = tables if tables.popwalk else None
            plain = None
            if not pop_edges and dg.w is not None:
                plain = tables if not tables.popwalk else dg.plain_node_tables()
            walks, lens = dg.walk_reject(float(self.p), float(self.q), starts, int(walk_length), self.seed,
                                         base, node_tables=plain, first_tables=first, pop_edges=pop_edges)
        corpus = WalkCorpus(walks, lens, dg.labels, shard)
        corpus.n_ids = dg.n
        return corpus

    def simulate_walks(self, num_walks, walk_length, nodes=None, verbose=False):
        """node2vec.py:81-95."""
        return self._simulate(num_walks, walk_length, nodes, self._prepared(), verbose)

    def simulate_walks_on_the_fly(self, num_walks, walk_length, nodes=None, verbose=False):
        """node2vec.py:97-111: same walks without a prior preprocess call. popwalk "pop" follows
        get_alias_nodes_cur / get_alias_edge_pop (:13-32,:154-174), whose edge law ignores q."""
        pop = self.popwalk == "pop"
        key = ("otf", float(self.p), float(self.q), pop)
        self._dg
        if self._prep == "plain" and not pop and self._tables is not None and self._otf_key is None:
            return self._simulate(num_walks, walk_length, nodes, self._tables, verbose)   # same law, tables at hand
        if self._tables is None or getattr(self, "_otf_key", None) != key:
            dg = self._dg
            if pop and self._use_alias():
                self._tables = dg.build_alias_tables(float(self.p), float(self.q), popwalk=True, pop_edges=True)
            elif pop:             # get_alias_edge_pop by rejection: popularity node tables + the w/pop candidate law
                self._tables = dg.build_node_tables(popwalk=True)
                self._tables.pop_edges = True
            else:
                self._tables = self._build(False, edges=self._use_alias())
            self._tables_raw = None
            self._otf_key = key
        return self._simulate(num_walks, walk_length, nodes, self._tables, verbose)

    def node2vec_walk(self, walk_length, start_node):
        """node2vec.py:55-79: one walk (one launch; use simulate_walks for throughput)."""
        return self.simulate_walks(1, walk_length, nodes=[start_node])[0]

    def node2vec_walk_on_the_fly(self, walk_length, start_node):
        return self.simulate_walks_on_the_fly(1, walk_length, nodes=[start_node])[0]

    def get_alias_edge(self, src, dst):
        """node2vec.py:133-152 -> (J, q) of one arc."""
        return _TableView(self, edges=True)[(src, dst)]

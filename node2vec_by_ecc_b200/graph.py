"""Device-resident CSR graph, alias tables and walk launchers (host side of the C ABI).

Data layout in HBM (all torch-owned, contiguous):
  row_ptr  int64[N+1]          col  int32[nnz] ascending per row          w  float64[nnz] | None
  node_slots  {int32 alias, uint32 thr}[nnz]   (viewed as int32[nnz, 2])
  etab_ptr int64[nnz+1]        edge_slots {int32, uint32}[sum_e deg(col[e])]
  walks    int32[n_walks, L]  (-1 padded)      lens int32[n_walks]
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from ._lib import check, lib, ptr, require_cuda, stream


@dataclass
class AliasTables:
    node_slots: torch.Tensor
    etab_ptr: torch.Tensor | None = None
    edge_slots: torch.Tensor | None = None
    # the reference's raw (J, q) (node2vec.py:240-269), kept only when asked for (tests)
    node_J: torch.Tensor | None = None
    node_q: torch.Tensor | None = None
    edge_J: torch.Tensor | None = None
    edge_q: torch.Tensor | None = None
    p: float = 1.0
    q: float = 1.0
    popwalk: bool = False
    pop_edges: bool = False                   # edge law is get_alias_edge_pop (node2vec.py:154-174)
    packed_rows: torch.Tensor | None = None   # row_ptr << 24 | deg per node
    arc_recs: torch.Tensor | None = None      # int64[nnz, 4]: one 32-byte record per arc


@dataclass
class DeviceGraph:
    row_ptr: torch.Tensor
    col: torch.Tensor
    w: torch.Tensor | None
    symmetric: bool
    labels: np.ndarray | None = None          # original node labels, sorted (compact id -> label)
    order: torch.Tensor | None = None         # compact ids in list(G.nodes()) order (walk start order)
    is_item: torch.Tensor | None = None       # uint8[N]: label starts with '9999999' (popularity walks)
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def n(self) -> int:
        return self.row_ptr.shape[0] - 1

    @property
    def nnz(self) -> int:
        return self.col.shape[0]

    @property
    def device(self):
        return self.row_ptr.device

    # ---- construction ----------------------------------------------------------------------
    @classmethod
    def from_coo(cls, src, dst, w, n_nodes: int, undirected: bool, labels=None, order=None,
                 is_item=None) -> "DeviceGraph":
        """Build the CSR on the device (n2v_csr_from_coo): sort, collapse duplicates (last wins,
        as networkx add_edge), add reverse arcs when undirected."""
        dev = require_cuda()
        src = torch.as_tensor(src, dtype=torch.int32).to(dev).contiguous()
        dst = torch.as_tensor(dst, dtype=torch.int32).to(dev).contiguous()
        wt = None if w is None else torch.as_tensor(w, dtype=torch.float64).to(dev).contiguous()
        m = int(src.shape[0])
        cap = max((2 * m) if undirected else m, 1)
        L = lib()
        ws_bytes = int(L.n2v_csr_workspace_bytes(C.c_int64(m), C.c_int32(n_nodes), C.c_int(int(undirected))))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        row_ptr = torch.empty(n_nodes + 1, dtype=torch.int64, device=dev)
        col = torch.empty(cap, dtype=torch.int32, device=dev)
        w_out = None if wt is None else torch.empty(cap, dtype=torch.float64, device=dev)
        nnz_d = torch.zeros(1, dtype=torch.int64, device=dev)
        check(L.n2v_csr_from_coo(ptr(src), ptr(dst), ptr(wt), C.c_int64(m), C.c_int32(n_nodes),
                                 C.c_int(int(undirected)), ptr(ws), C.c_size_t(ws_bytes), ptr(row_ptr),
                                 ptr(col), ptr(w_out), ptr(nnz_d), stream()))
        nnz = int(nnz_d.item())
        # (a zero-length view keeps a non-NULL data pointer for the C ABI)
        col = col[:nnz].clone() if 0 < nnz < cap else col[:nnz]
        if w_out is not None:
            w_out = w_out[:nnz].clone() if 0 < nnz < cap else w_out[:nnz]
        if order is not None:
            order = torch.as_tensor(order, dtype=torch.int32).to(dev).contiguous()
        if is_item is not None:
            is_item = torch.as_tensor(is_item, dtype=torch.uint8).to(dev).contiguous()
        return cls(row_ptr, col, w_out, bool(undirected), labels, order, is_item)

    @classmethod
    def from_csr(cls, row_ptr, col, w=None, symmetric=True, labels=None, order=None, is_item=None):
        dev = require_cuda()
        rp = torch.as_tensor(row_ptr, dtype=torch.int64).to(dev).contiguous()
        c = torch.as_tensor(col, dtype=torch.int32).to(dev).contiguous()
        wt = None if w is None else torch.as_tensor(w, dtype=torch.float64).to(dev).contiguous()
        if order is not None:
            order = torch.as_tensor(order, dtype=torch.int32).to(dev).contiguous()
        if is_item is not None:
            is_item = torch.as_tensor(is_item, dtype=torch.uint8).to(dev).contiguous()
        return cls(rp, c, wt, bool(symmetric), labels, order, is_item)

    @classmethod
    def from_networkx(cls, G, is_directed=None) -> "DeviceGraph":
        """Ingest the networkx graph the reference's read_graph hands to node2vec.Graph
        (main.py:66-80). Compact ids follow sorted(labels) so that CSR row order equals the
        reference's sorted(G.neighbors(v)); `order` keeps list(G.nodes()) for the walk order."""
        nodes = list(G.nodes())
        labels = sorted(nodes)
        idx = {l: i for i, l in enumerate(labels)}
        directed = G.is_directed()
        m = G.number_of_edges()
        src = np.empty(m, dtype=np.int32)
        dst = np.empty(m, dtype=np.int32)
        wts = np.empty(m, dtype=np.float64)
        unit = True
        for i, (u, v, wt) in enumerate(G.edges(data="weight", default=1)):
            src[i] = idx[u]; dst[i] = idx[v]; wts[i] = wt
            unit &= (wt == 1)
        order = np.fromiter((idx[x] for x in nodes), dtype=np.int32, count=len(nodes))
        is_item = np.fromiter((str(l).startswith("9999999") for l in labels), dtype=np.uint8, count=len(labels))
        try:
            lab = np.asarray(labels)
            if lab.dtype == object:
                lab = np.asarray(labels, dtype=object)
        except Exception:
            lab = np.asarray(labels, dtype=object)
        return cls.from_coo(src, dst, None if unit else wts, len(labels), undirected=not directed,
                            labels=lab, order=order, is_item=is_item if is_item.any() else None)

    # ---- alias tables -------------------------------------------------------------------------
    def etab_offsets(self) -> torch.Tensor:
        if "etab" not in self._cache:
            L = lib()
            nnz = self.nnz
            ws_bytes = int(L.n2v_etab_workspace_bytes(C.c_int64(nnz)))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            etab = torch.empty(nnz + 1, dtype=torch.int64, device=self.device)
            check(L.n2v_etab_offsets(ptr(self.row_ptr), ptr(self.col), C.c_int32(self.n), C.c_int64(nnz),
                                     ptr(etab), ptr(ws), C.c_size_t(ws_bytes), stream()))
            self._cache["etab"] = etab
        return self._cache["etab"]

    def sum_deg_sq(self) -> int:
        """Total edge-table entries, sum_e deg(col[e]) (== sum deg^2 when undirected)."""
        if "sds" not in self._cache:
            self._cache["sds"] = int(self.etab_offsets()[-1].item())
        return self._cache["sds"]

    def edge_table_bytes(self) -> int:
        return 8 * self.sum_deg_sq()

    def build_node_tables(self, popwalk=False, keep_raw=False, raw_probs=False) -> AliasTables:
        L = lib()
        nnz, dev = self.nnz, self.device
        slots = torch.empty((max(nnz, 1), 2), dtype=torch.int32, device=dev)
        J = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        q = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
        check(L.n2v_alias_build_nodes(ptr(self.row_ptr), ptr(self.col), ptr(self.w), C.c_int32(self.n),
                                      ptr(self.is_item), C.c_int(int(bool(popwalk)) | (2 if raw_probs else 0)), ptr(slots), ptr(J), ptr(q),
                                      stream()))
        t = AliasTables(node_slots=slots, popwalk=bool(popwalk))
        if keep_raw:
            t.node_J, t.node_q = J[:nnz], q[:nnz]
        return t

    def build_alias_tables(self, p: float, q: float, popwalk=False, keep_raw=False,
                           chunk_entries: int = 1 << 27, pop_edges=False) -> AliasTables:
        """preprocess_transition_probs (node2vec.py:176-204) on the device. popwalk: popularity node
        tables (:213-218); pop_edges: get_alias_edge_pop edge tables (:154-174, the on-the-fly law)."""
        L = lib()
        dev = self.device
        t = self.build_node_tables(popwalk=popwalk, keep_raw=keep_raw)
        t.p, t.q, t.pop_edges = float(p), float(q), bool(pop_edges)
        etab = self.etab_offsets()
        total = self.sum_deg_sq()
        t.etab_ptr = etab
        t.edge_slots = torch.empty((max(total, 1), 2), dtype=torch.int32, device=dev)
        nnz = self.nnz
        if nnz == 0:
            return t
        if keep_raw or total <= chunk_entries:
            bounds = [0, nnz]
        else:  # bound the scratch: split arcs where the table offsets cross multiples of chunk_entries
            marks = torch.arange(chunk_entries, total, chunk_entries, device=dev, dtype=torch.int64)
            cuts = torch.searchsorted(etab, marks, right=True) - 1
            bounds = sorted(set([0, nnz] + [int(c) for c in cuts.tolist()]))
        wJ = wq = None
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b <= a:
                continue
            ne = int((etab[b] - etab[a]).item()) if len(bounds) > 2 else total
            if wJ is None or wJ.shape[0] < ne:
                wJ = torch.empty(max(ne, 1), dtype=torch.int32, device=dev)
                wq = torch.empty(max(ne, 1), dtype=torch.float64, device=dev)
            check(L.n2v_alias_build_edges(ptr(self.row_ptr), ptr(self.col), ptr(self.w), C.c_int32(self.n),
                                          C.c_double(p), C.c_double(q), C.c_int(int(self.symmetric)), C.c_int(int(pop_edges)),
                                          ptr(etab), C.c_int64(a), C.c_int64(b), ptr(t.edge_slots),
                                          ptr(wJ), ptr(wq), stream()))
        if keep_raw:
            t.edge_J, t.edge_q = wJ[:total], wq[:total]
        # packed arc records for n2v_walk_alias_packed (32 B per arc, small next to the tables)
        L_ = lib()
        packed = torch.empty(max(self.n, 1), dtype=torch.int64, device=dev)
        recs = torch.empty((nnz, 4), dtype=torch.int64, device=dev)          # torch allocations are 512-B aligned
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        check(L_.n2v_pack_rows(ptr(self.row_ptr), C.c_int32(self.n), ptr(packed), ptr(flag), stream()))
        check(L_.n2v_pack_arcs(ptr(self.row_ptr), ptr(self.col), ptr(etab), C.c_int64(nnz), ptr(recs), ptr(flag),
                               stream()))
        if int(flag.item()) == 0:
            t.packed_rows, t.arc_recs = packed, recs
        return t

    # ---- walks ------------------------------------------------------------------------------------
    def walk_alias(self, tables: AliasTables, starts: torch.Tensor, L_: int, seed: int,
                   walk_id_base: int = 0, out=None, packed: bool = True):
        """simulate_walks (node2vec.py:81-95): -> (walks int32[n, L] padded -1, lens int32[n]).
        packed=True walks over the 32-byte arc records (n2v_walk_alias_packed) when the tables carry
        them, False over row_ptr / etab_ptr / col (n2v_walk_alias); the output is identical."""
        starts = torch.as_tensor(starts, dtype=torch.int32).to(self.device).contiguous()
        n = int(starts.shape[0])
        walks, lens = out if out is not None else (
            torch.empty((n, L_), dtype=torch.int32, device=self.device),
            torch.empty(n, dtype=torch.int32, device=self.device))
        if packed and tables.arc_recs is not None:
            check(lib().n2v_walk_alias_packed(ptr(tables.packed_rows), ptr(tables.node_slots), ptr(tables.arc_recs),
                                              ptr(tables.edge_slots), ptr(starts), C.c_int64(n), C.c_int32(L_),
                                              C.c_uint64(seed), C.c_uint64(walk_id_base), ptr(walks), ptr(lens),
                                              stream()))
            return walks, lens
        check(lib().n2v_walk_alias(ptr(self.row_ptr), ptr(self.col), ptr(tables.node_slots),
                                   ptr(tables.etab_ptr), ptr(tables.edge_slots), ptr(starts), C.c_int64(n),
                                   C.c_int32(L_), C.c_uint64(seed), C.c_uint64(walk_id_base), ptr(walks),
                                   ptr(lens), stream()))
        return walks, lens

    def reject_index(self):
        """(packed_rows uint64[N], edge_hash uint64[cap], cap) for n2v_walk_reject_indexed, built
        once per graph; None when a degree / offset does not fit the packed row word."""
        if "rix" not in self._cache:
            L = lib()
            dev = self.device
            packed = torch.empty(max(self.n, 1), dtype=torch.int64, device=dev)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            check(L.n2v_pack_rows(ptr(self.row_ptr), C.c_int32(self.n), ptr(packed), ptr(flag), stream()))
            if int(flag.item()):
                self._cache["rix"] = None
            else:
                cap = int(L.n2v_edge_hash_capacity(C.c_int64(self.nnz)))
                table = torch.empty(cap, dtype=torch.int64, device=dev)
                check(L.n2v_edge_hash_build(ptr(self.row_ptr), ptr(self.col), C.c_int32(self.n), C.c_int64(self.nnz),
                                            ptr(table), C.c_uint64(cap), stream()))
                strength = None
                if self.w is not None and self.symmetric:
                    rows = torch.repeat_interleave(torch.arange(self.n, device=dev), self.row_ptr[1:] - self.row_ptr[:-1])
                    strength = torch.zeros(max(self.n, 1), dtype=torch.float64, device=dev).index_add_(0, rows, self.w)
                self._cache["rix"] = (packed, table, cap, strength)
        return self._cache["rix"]

    def walk_reject(self, p: float, q: float, starts: torch.Tensor, L_: int, seed: int,
                    walk_id_base: int = 0, node_tables: AliasTables | None = None, counters=None,
                    out=None, indexed: bool = True, first_tables: AliasTables | None = None,
                    pop_edges: bool = False):
        """Same law as get_alias_edge (node2vec.py:142-150) by rejection sampling; no edge tables.
        indexed=True uses the hashed distance-1 test + per-lane state machine (n2v_walk_reject_indexed,
        16 bytes per arc of index), False the binary-search form (n2v_walk_reject, no extra memory).
        node_tables: the candidate law of steps >= 2 (plain weights; built here when missing on a
        weighted graph). first_tables: table of the first step when it differs (the popularity node
        tables, node2vec.py:213-218). pop_edges: steps >= 2 follow get_alias_edge_pop (:154-174);
        node_tables must then be built over w / len(G[nbr]) for every node (pop_all_node_tables)."""
        starts = torch.as_tensor(starts, dtype=torch.int32).to(self.device).contiguous()
        n = int(starts.shape[0])
        if pop_edges and node_tables is None:
            node_tables = self.pop_all_node_tables()
        if self.w is not None and node_tables is None:
            node_tables = self.build_node_tables()
        law = None
        if first_tables is not None or pop_edges:
            from ._lib import WalkLaw
            law = WalkLaw(C.c_void_p(first_tables.node_slots.data_ptr() if first_tables is not None else 0),
                          C.c_int32(int(bool(pop_edges))))
            law = C.byref(law)
        walks, lens = out if out is not None else (
            torch.empty((n, L_), dtype=torch.int32, device=self.device),
            torch.empty(n, dtype=torch.int32, device=self.device))
        rix = self.reject_index() if indexed else None
        if rix is not None:
            packed, table, cap, strength = rix
            check(lib().n2v_walk_reject_indexed_law(
                ptr(packed), ptr(self.col), C.c_int64(self.nnz), ptr(self.w), ptr(strength),
                ptr(node_tables.node_slots if node_tables is not None else None), law, ptr(table), C.c_uint64(cap),
                C.c_double(p), C.c_double(q), C.c_int(int(self.symmetric)), ptr(starts), C.c_int64(n),
                C.c_int32(L_), C.c_uint64(seed), C.c_uint64(walk_id_base), ptr(walks), ptr(lens), ptr(counters),
                stream()))
            return walks, lens
        check(lib().n2v_walk_reject_law(ptr(self.row_ptr), ptr(self.col), ptr(self.w),
                                        ptr(node_tables.node_slots if node_tables is not None else None), law,
                                        C.c_double(p), C.c_double(q), C.c_int(int(self.symmetric)), ptr(starts),
                                        C.c_int64(n), C.c_int32(L_), C.c_uint64(seed), C.c_uint64(walk_id_base),
                                        ptr(walks), ptr(lens), ptr(counters), stream()))
        return walks, lens

    def plain_node_tables(self) -> AliasTables:
        if "plain_nodes" not in self._cache:
            self._cache["plain_nodes"] = self.build_node_tables()
        return self._cache["plain_nodes"]

    def pop_all_node_tables(self) -> AliasTables:
        """node tables over w / len(G[nbr]) for EVERY node: the candidate law of get_alias_edge_pop
        (node2vec.py:154-174, which -- unlike the node law :213-218 -- has no item exception)."""
        if "popall" not in self._cache:
            item, self.is_item = self.is_item, None
            try:
                self._cache["popall"] = self.build_node_tables(popwalk=True)
            finally:
                self.is_item = item
        return self._cache["popall"]

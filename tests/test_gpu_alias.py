"""GPU: alias-table builder and CSR builder through the C ABI vs the reference's golden vectors
and the C oracle. Bar: J exact, q within 1e-6 (north-star); on every case here q is bit-equal."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import CASES, GOLDEN, load_case, random_graph

pytestmark = pytest.mark.gpu


def dev_graph(g, symmetric, is_item=None):
    from node2vec_by_ecc_b200 import DeviceGraph
    return DeviceGraph.from_csr(g.row_ptr, g.col, g.w, symmetric=symmetric, is_item=is_item)


def test_library_loaded_and_device():
    from node2vec_by_ecc_b200 import lib
    assert lib().n2v_version() >= 100
    assert lib().n2v_sm_count() > 0


def test_alias_setup_known_answers():
    from node2vec_by_ecc_b200 import alias_setup
    with open(os.path.join(GOLDEN, "alias_setup.json")) as f:
        cases = json.load(f)
    for c in cases:
        J, q = alias_setup(c["probs"])
        assert J.tolist() == c["J"]
        assert np.abs(q - np.asarray(c["q"])).max() <= 1e-6
        assert q.tolist() == c["q"]           # bit-equal: same float64 operation order


@pytest.mark.parametrize("name", CASES)
def test_tables_match_reference_golden(name):
    z, g = load_case(name)
    dg = dev_graph(g, symmetric=not bool(int(z["directed"])), is_item=z["is_item"])
    t = dg.build_alias_tables(float(z["p"]), float(z["q"]), popwalk=bool(int(z["popwalk"])), keep_raw=True)
    assert dg.sum_deg_sq() == z["eq"].shape[0]
    assert (t.etab_ptr.cpu().numpy() == z["etab_ptr"]).all()
    assert (t.node_J.cpu().numpy() == z["nJ"]).all()
    assert (t.edge_J.cpu().numpy() == z["eJ"]).all()
    assert np.abs(t.node_q.cpu().numpy() - z["nq"]).max() <= 1e-6
    assert np.abs(t.edge_q.cpu().numpy() - z["eq"]).max() <= 1e-6
    assert (t.node_q.cpu().numpy() == z["nq"]).all() and (t.edge_q.cpu().numpy() == z["eq"]).all()


def check_slots(slots, J, q):
    """packed {alias, thr} == (J, ceil(q*2^32)) with the always-accept convention"""
    s = slots.cpu().numpy()
    thr = np.ceil(q * 4294967296.0)
    sat = thr >= 4294967296.0
    want_thr = np.where(sat, 0xFFFFFFFF, thr).astype(np.uint64).astype(np.uint32)
    assert (s[:, 1].view(np.uint32) == want_thr).all()
    assert (s[~sat, 0] == J[~sat]).all()


@pytest.mark.parametrize("weighted,directed,p,q", [(False, False, 0.25, 4.0), (True, False, 0.5, 2.0),
                                                    (True, True, 4.0, 0.25), (True, False, 0.3, 3.0)])
def test_tables_match_oracle_random_graph(weighted, directed, p, q):
    _, g = random_graph(2000, 30000, seed=11, weighted=weighted, directed=directed, skew=1.0)
    to = oracle.preprocess(g, p, q)
    dg = dev_graph(g, symmetric=not directed)
    t = dg.build_alias_tables(p, q, keep_raw=True)
    assert (t.etab_ptr.cpu().numpy() == to.etab_ptr).all()
    assert (t.node_J.cpu().numpy() == to.nJ).all() and (t.edge_J.cpu().numpy() == to.eJ).all()
    assert np.abs(t.edge_q.cpu().numpy() - to.eq).max() <= 1e-6
    assert (t.edge_q.cpu().numpy() == to.eq).all() and (t.node_q.cpu().numpy() == to.nq).all()
    check_slots(t.edge_slots[:to.eJ.shape[0]], to.eJ, to.eq)
    check_slots(t.node_slots[:to.nJ.shape[0]], to.nJ, to.nq)


def test_chunked_build_equals_single_launch():
    _, g = random_graph(1500, 20000, seed=5, skew=0.5)
    dg = dev_graph(g, symmetric=True)
    a = dg.build_alias_tables(0.25, 4.0)
    b = dg.build_alias_tables(0.25, 4.0, chunk_entries=50000)
    assert torch.equal(a.edge_slots, b.edge_slots)


@pytest.mark.parametrize("undirected,weighted", [(True, False), (True, True), (False, True)])
def test_csr_from_coo_matches_oracle(undirected, weighted):
    from node2vec_by_ecc_b200 import DeviceGraph
    rng = np.random.RandomState(3)
    n, m = 500, 6000
    a, b = rng.randint(0, n, size=m), rng.randint(0, n, size=m)   # duplicates and self-loops included
    w = rng.rand(m) if weighted else None
    want = oracle.csr_from_coo(a, b, w, n, undirected=undirected)
    dg = DeviceGraph.from_coo(a, b, w, n, undirected=undirected)
    assert (dg.row_ptr.cpu().numpy() == want.row_ptr).all()
    assert (dg.col.cpu().numpy() == want.col).all()
    if weighted:
        assert (dg.w.cpu().numpy() == want.w).all()


def test_empty_and_isolated():
    from node2vec_by_ecc_b200 import DeviceGraph
    dg = DeviceGraph.from_coo(np.zeros(0, np.int32), np.zeros(0, np.int32), None, 5, undirected=True)
    assert dg.nnz == 0 and dg.sum_deg_sq() == 0
    t = dg.build_alias_tables(1.0, 1.0)
    walks, lens = dg.walk_alias(t, torch.arange(5, dtype=torch.int32), 6, seed=1)
    assert (lens.cpu().numpy() == 1).all()
    assert (walks.cpu().numpy()[:, 0] == np.arange(5)).all() and (walks.cpu().numpy()[:, 1:] == -1).all()


def test_alias_setup_many_random_vectors_vs_oracle():
    """generated probability vectors (uniform, one-hot, zeros, heavy-tailed, sizes
    1..3000) through ONE launch -- every row of a star-shaped CSR carries one vector, normalisation
    off -- against oracle.alias_setup == node2vec.py:240-269. J exact, q bit-equal."""
    from node2vec_by_ecc_b200 import DeviceGraph
    rng = np.random.RandomState(7)
    vecs = []
    for K in [1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 64, 100, 257, 1000, 3000]:
        for kind in range(5):
            if kind == 0:
                v = np.full(K, 1.0 / K)
            elif kind == 1:
                v = np.zeros(K); v[rng.randint(K)] = 1.0
            elif kind == 2:
                v = rng.rand(K) ** 6; v /= v.sum()
            elif kind == 3:
                v = rng.randint(0, 4, size=K).astype(np.float64); v = v / max(v.sum(), 1.0)   # zeros inside
            else:
                v = rng.dirichlet(np.full(K, 0.05))
            vecs.append(v)
    lens = np.asarray([len(v) for v in vecs])
    row_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n_rows = len(vecs)
    # rows 0..n_rows-1 carry the vectors, all pointing at leaf nodes with empty rows
    n = n_rows + int(lens.max())
    rp = np.concatenate([row_ptr, np.full(n - n_rows, row_ptr[-1])]).astype(np.int64)
    col = np.concatenate([np.arange(n_rows, n_rows + k, dtype=np.int32) for k in lens])
    dg = DeviceGraph.from_csr(rp, col, np.concatenate(vecs), symmetric=False)
    t = dg.build_node_tables(keep_raw=True, raw_probs=True)
    J, q = t.node_J.cpu().numpy(), t.node_q.cpu().numpy()
    for i, v in enumerate(vecs):
        wj, wq = oracle.alias_setup(v)
        a, b = row_ptr[i], row_ptr[i + 1]
        assert (J[a:b] == wj).all(), (i, len(v))
        assert (q[a:b] == wq).all(), (i, len(v))
    check_slots(t.node_slots[:row_ptr[-1]], J.astype(np.int64), q)

"""CPU: the C oracle (oracle/n2v_oracle.c) against vectors produced by the reference itself
(tests/golden/, written by oracle/make_golden.py from /root/reference/src/node2vec.py)."""
import glob
import json
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = oracle.CSR(z["row_ptr"], z["col"], z["w"] if int(z["weighted"]) else None)
    return z, g


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in oracle.philox4x32_10(ctr, key)] == want


def test_alias_setup_known_answers():
    with open(os.path.join(GOLDEN, "alias_setup.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 15
    for c in cases:
        J, q = oracle.alias_setup(c["probs"])
        assert J.tolist() == c["J"]
        assert q.tolist() == c["q"]          # bit-equal: same float64 operation order


def test_alias_setup_survey_vectors():
    # SURVEY.md section 8c, generated from node2vec.py:240-269
    J, q = oracle.alias_setup([0.1, 0.2, 0.7])
    assert J.tolist() == [2, 2, 0]
    assert q.tolist() == [0.30000000000000004, 0.6000000000000001, 0.9999999999999998]
    J, q = oracle.alias_setup([0.4, 0.1, 0.1, 0.4])
    assert J.tolist() == [0, 3, 3, 0] and q.tolist() == [1.0, 0.4, 0.4, 0.3999999999999999]
    J, q = oracle.alias_setup([1.0 / 49] * 49)
    assert (J == 0).all() and (q == 0.9999999999999999).all()


@pytest.mark.parametrize("name", CASES)
def test_tables_match_reference(name):
    z, g = load_case(name)
    t = oracle.preprocess(g, float(z["p"]), float(z["q"]), is_item=z["is_item"],
                          popwalk_nodes=bool(int(z["popwalk"])))
    assert (t.etab_ptr == z["etab_ptr"]).all()
    assert (t.nJ == z["nJ"]).all() and (t.eJ == z["eJ"]).all()
    assert (t.nq == z["nq"]).all() and (t.eq == z["eq"]).all()   # bit-equal
    assert oracle.sum_deg_sq(g) == z["eq"].shape[0]


@pytest.mark.parametrize("name", CASES)
def test_walks_match_reference(name):
    z, g = load_case(name)
    p, q, L, seed = float(z["p"]), float(z["q"]), int(z["L"]), int(z["seed"])
    pop = bool(int(z["popwalk"]))
    t = oracle.preprocess(g, p, q, is_item=z["is_item"], popwalk_nodes=pop)
    starts = np.tile(z["order"], int(z["R"]))
    wk, lens = oracle.walks_alias(g, t, starts, L, seed)
    assert (lens == z["lens"]).all() and (wk == z["walks"]).all()
    wk2, lens2 = oracle.walks_on_the_fly(g, p, q, starts, L, seed, is_item=z["is_item"], popwalk=pop)
    assert (lens2 == z["lens_otf"]).all() and (wk2 == z["walks_otf"]).all()
    sub = np.tile(z["sub_starts"], 2)
    wk3, lens3 = oracle.walks_alias(g, t, sub, L, seed, walk_id_base=1000)
    assert (lens3 == z["lens_sub"]).all() and (wk3 == z["walks_sub"]).all()


def test_dead_ends_present_in_directed_case():
    z, _ = load_case("dir_p025_q4")
    assert (z["lens"] < int(z["L"])).any() and (z["lens"] == int(z["L"])).any()


def test_csr_from_coo_matches_fixture():
    z, g = load_case("rndw_p05_q2")
    src = np.repeat(np.arange(g.n), np.diff(g.row_ptr))
    keep = src <= g.col                       # one orientation per undirected edge
    g2 = oracle.csr_from_coo(src[keep], g.col[keep], g.w[keep], g.n, undirected=True)
    assert (g2.row_ptr == g.row_ptr).all() and (g2.col == g.col).all() and (g2.w == g.w).all()


def test_csr_from_coo_is_networkx_last_write_wins():
    import networkx as nx
    rng = np.random.RandomState(3)
    n, m = 60, 900
    a, b, w = rng.randint(0, n, size=m), rng.randint(0, n, size=m), rng.rand(m)
    for undirected in (True, False):
        G = nx.Graph() if undirected else nx.DiGraph()
        G.add_nodes_from(range(n))
        for x, y, ww in zip(a.tolist(), b.tolist(), w.tolist()):
            G.add_edge(x, y, weight=ww)
        g = oracle.csr_from_coo(a, b, w, n, undirected=undirected)
        for v in range(n):
            nb = sorted(G.neighbors(v))
            assert g.col[g.row_ptr[v]:g.row_ptr[v + 1]].tolist() == nb
            assert g.w[g.row_ptr[v]:g.row_ptr[v + 1]].tolist() == [G[v][x]["weight"] for x in nb]

"""GPU: walks through the C ABI. Alias mode is bit-exact against (a) walks the reference itself
produced with Philox uniforms injected into np.random.rand (tests/golden) and (b) the C oracle on
larger seeded graphs; rejection mode passes chi-square against the exact transition rows."""
import numpy as np
import pytest
import torch

import oracle
from helpers import CASES, chi_square_p, load_case, random_graph

pytestmark = pytest.mark.gpu


def dev_graph(g, symmetric, is_item=None):
    from node2vec_by_ecc_b200 import DeviceGraph
    return DeviceGraph.from_csr(g.row_ptr, g.col, g.w, symmetric=symmetric, is_item=is_item)


@pytest.mark.parametrize("name", CASES)
def test_walks_bit_exact_vs_reference_golden(name):
    z, g = load_case(name)
    L, seed = int(z["L"]), int(z["seed"])
    dg = dev_graph(g, symmetric=not bool(int(z["directed"])), is_item=z["is_item"])
    t = dg.build_alias_tables(float(z["p"]), float(z["q"]), popwalk=bool(int(z["popwalk"])))
    starts = torch.as_tensor(np.tile(z["order"], int(z["R"])))
    for packed in (True, False):
        walks, lens = dg.walk_alias(t, starts, L, seed, packed=packed)
        assert (lens.cpu().numpy() == z["lens"]).all()
        assert (walks.cpu().numpy() == z["walks"]).all()
    # a contiguous shard of start nodes with a walk-id base (main_link.py:263-264 partitioning)
    sub = torch.as_tensor(np.tile(z["sub_starts"], 2))
    walks, lens = dg.walk_alias(t, sub, L, seed, walk_id_base=1000)
    assert (walks.cpu().numpy() == z["walks_sub"]).all() and (lens.cpu().numpy() == z["lens_sub"]).all()


def test_popularity_on_the_fly_walks_bit_exact_vs_reference_golden():
    """popwalk="pop": simulate_walks_on_the_fly follows get_alias_nodes_cur / get_alias_edge_pop
    (node2vec.py:13-32,:154-174) -- the golden `walks_otf` of the '9999999'-prefixed bipartite case"""
    z, g = load_case("bip_pop_p1_q1")
    assert not (z["walks"] == z["walks_otf"]).all()       # the two popularity laws differ
    dg = dev_graph(g, symmetric=True, is_item=z["is_item"])
    t = dg.build_alias_tables(float(z["p"]), float(z["q"]), popwalk=True, pop_edges=True)
    starts = torch.as_tensor(np.tile(z["order"], int(z["R"])))
    walks, lens = dg.walk_alias(t, starts, int(z["L"]), int(z["seed"]))
    assert (walks.cpu().numpy() == z["walks_otf"]).all() and (lens.cpu().numpy() == z["lens_otf"]).all()


@pytest.mark.parametrize("weighted,directed,p,q,L", [(False, False, 0.25, 4.0, 80), (True, False, 0.5, 2.0, 40),
                                                      (True, True, 4.0, 0.25, 33), (False, True, 1.0, 1.0, 7)])
def test_walks_bit_exact_vs_oracle(weighted, directed, p, q, L):
    _, g = random_graph(2000, 30000, seed=21, weighted=weighted, directed=directed, skew=1.0)
    to = oracle.preprocess(g, p, q)
    starts = np.tile(np.arange(g.n, dtype=np.int32), 3)
    want, want_len = oracle.walks_alias(g, to, starts, L, seed=77, walk_id_base=5)
    dg = dev_graph(g, symmetric=not directed)
    t = dg.build_alias_tables(p, q)
    assert t.arc_recs is not None
    for packed in (True, False):                      # both alias walkers: same tokens
        walks, lens = dg.walk_alias(t, torch.as_tensor(starts), L, seed=77, walk_id_base=5, packed=packed)
        assert (lens.cpu().numpy() == want_len).all()
        assert (walks.cpu().numpy() == want).all()
    if directed:
        assert (want_len < L).any()          # dead ends exercised


def test_sharding_is_invisible():
    """walk ids are global: any split of the start list gives the same corpus (multi-GPU rule)."""
    _, g = random_graph(3000, 40000, seed=8, skew=0.7)
    dg = dev_graph(g, symmetric=True)
    t = dg.build_alias_tables(0.25, 4.0)
    starts = torch.arange(g.n, dtype=torch.int32).repeat(2)
    full, _ = dg.walk_alias(t, starts, 40, seed=3)
    cuts = [0, 1111, 2500, 4096, starts.shape[0]]
    parts = [dg.walk_alias(t, starts[a:b], 40, seed=3, walk_id_base=a)[0] for a, b in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(full, torch.cat(parts))


def transition_counts(walks, lens, n):
    w = walks.cpu().numpy()
    ok = lens.cpu().numpy() >= 3
    w = w[ok]
    key = (w[:, 0].astype(np.int64) * n + w[:, 1]) * n + w[:, 2]
    u, c = np.unique(key, return_counts=True)
    return u, c


@pytest.mark.parametrize("weighted,directed,p,q", [(False, False, 0.25, 4.0), (False, False, 4.0, 0.25),
                                                    (False, False, 1.0, 1.0), (True, False, 0.25, 4.0),
                                                    (True, True, 0.5, 2.0), (False, True, 4.0, 0.5)])
@pytest.mark.parametrize("indexed", [True, False])
def test_rejection_walker_chi_square(weighted, directed, p, q, indexed):
    n = 300
    _, g = random_graph(n, 3000, seed=31, weighted=weighted, directed=directed, skew=1.0)
    dg = dev_graph(g, symmetric=not directed)
    deg = np.diff(g.row_ptr)
    hub = int(np.argmax(deg))
    # walks of 3 tokens: (start=a) -> b -> c ; c | (a, b) must follow get_alias_edge(a, b)
    starts_np = np.concatenate([np.full(400000, hub, dtype=np.int32),
                                np.repeat(np.arange(n, dtype=np.int32), 2000)])
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    walks, lens = dg.walk_reject(p, q, torch.as_tensor(starts_np), 3, seed=99, counters=counters, indexed=indexed)
    cnt = counters.cpu().numpy()
    assert cnt[0] == int((lens.cpu().numpy() - 1).sum()) and cnt[1] >= cnt[0]
    keys, c = transition_counts(walks, lens, n)
    a, b, nxt = keys // (n * n), (keys // n) % n, keys % n
    pair = a * n + b
    upair, tot = np.unique(pair, return_counts=False), None
    sums = {}
    for pk, cc in zip(pair, c):
        sums[pk] = sums.get(pk, 0) + cc
    tested, pvals = 0, []
    for pk in sorted(sums, key=lambda k: -sums[k])[:60]:
        aa, bb = int(pk // n), int(pk % n)
        probs = oracle.transition_row(g, p, q, aa, bb)
        row = g.col[g.row_ptr[bb]:g.row_ptr[bb + 1]]
        obs = np.zeros(len(row))
        sel = pair == pk
        pos = np.searchsorted(row, nxt[sel])
        assert (row[pos] == nxt[sel]).all()          # only real neighbours are ever produced
        obs[pos] = c[sel]
        pvals.append(chi_square_p(obs, probs))
        tested += 1
    assert tested >= 30
    pvals = np.asarray(pvals)
    # 60 independent tests: no catastrophic cell, and the p-values are not piled up near 0
    assert pvals.min() > 1e-5, pvals.min()
    assert (pvals < 0.01).sum() <= 4, np.sort(pvals)[:6]


def test_edge_hash_holds_exactly_the_arcs():
    _, g = random_graph(500, 6000, seed=17, directed=True, skew=0.8)
    dg = dev_graph(g, symmetric=False)
    packed, table, cap, _ = dg.reject_index()
    assert cap >= 2 * g.nnz and cap & (cap - 1) == 0
    t = table.cpu().numpy().view(np.uint64)
    keys = np.sort(t[t != np.uint64(0xFFFFFFFFFFFFFFFF)])
    src = np.repeat(np.arange(g.n, dtype=np.uint64), np.diff(g.row_ptr))
    want = np.sort((src << np.uint64(32)) | g.col.astype(np.uint64))
    assert np.array_equal(keys, want)
    pk = packed.cpu().numpy().view(np.uint64)
    assert np.array_equal(pk >> np.uint64(24), g.row_ptr[:-1].astype(np.uint64))
    assert np.array_equal(pk & np.uint64((1 << 24) - 1), np.diff(g.row_ptr).astype(np.uint64))


@pytest.mark.parametrize("L", [1, 2, 7, 8, 9, 33, 80])
def test_indexed_rejection_walk_shapes_and_dead_ends(L):
    """ragged lengths, -1 padding, every consecutive pair is an arc -- both rejection forms"""
    _, g = random_graph(400, 1500, seed=23, directed=True, skew=0.5)     # sparse: plenty of sinks
    dg = dev_graph(g, symmetric=False)
    starts = torch.arange(g.n, dtype=torch.int32).repeat(3)
    arcs = set(zip(np.repeat(np.arange(g.n), np.diff(g.row_ptr)).tolist(), g.col.tolist()))
    deg = np.diff(g.row_ptr)
    for indexed in (True, False):
        walks, lens = dg.walk_reject(0.5, 2.0, starts, L, seed=4, indexed=indexed)
        w, l = walks.cpu().numpy(), lens.cpu().numpy()
        assert w.shape == (starts.shape[0], L) and (w[:, 0] == starts.numpy()).all()
        for row, ln in zip(w[:200], l[:200]):
            assert 1 <= ln <= L and (row[ln:] == -1).all() and (row[:ln] >= 0).all()
            assert all((int(a), int(b)) in arcs for a, b in zip(row[:ln - 1], row[1:ln]))
            assert ln == L or deg[row[ln - 1]] == 0          # stops early only at a dead end


@pytest.mark.parametrize("weighted,directed", [(False, False), (False, True), (True, True)])
def test_both_rejection_forms_make_identical_walks(weighted, directed):
    """same Philox words, same decisions: the hashed/state-machine form reproduces the
    binary-search form token by token"""
    _, g = random_graph(2000, 30000, seed=41, weighted=weighted, directed=directed, skew=1.0)
    dg = dev_graph(g, symmetric=not directed)
    starts = torch.arange(g.n, dtype=torch.int32).repeat(2)
    c1 = torch.zeros(4, dtype=torch.int64, device="cuda"); c2 = torch.zeros_like(c1)
    a, la = dg.walk_reject(0.25, 4.0, starts, 40, seed=6, walk_id_base=77, counters=c1, indexed=False)
    b, lb = dg.walk_reject(0.25, 4.0, starts, 40, seed=6, walk_id_base=77, counters=c2, indexed=True)
    assert torch.equal(a, b) and torch.equal(la, lb) and torch.equal(c1, c2)


def test_rejection_first_step_follows_node_table():
    _, g = random_graph(200, 2500, seed=4, weighted=True, skew=0.5)
    dg = dev_graph(g, symmetric=True)
    v = int(np.argmax(np.diff(g.row_ptr)))
    walks, lens = dg.walk_reject(0.25, 4.0, torch.full((300000,), v, dtype=torch.int32), 2, seed=5)
    row = g.col[g.row_ptr[v]:g.row_ptr[v + 1]]
    probs = oracle.transition_row(g, 1.0, 1.0, -1, v)
    obs = np.bincount(np.searchsorted(row, walks.cpu().numpy()[:, 1]), minlength=len(row))
    assert chi_square_p(obs, probs) > 1e-4


def test_alias_and_rejection_agree_on_visit_frequencies():
    _, g = random_graph(1000, 15000, seed=13, skew=1.0)
    dg = dev_graph(g, symmetric=True)
    t = dg.build_alias_tables(0.25, 4.0)
    starts = torch.arange(g.n, dtype=torch.int32).repeat(20)
    def visit_law(walks):
        f = np.bincount(walks.cpu().numpy().ravel(), minlength=g.n).astype(np.float64)
        return f / f.sum()
    fa1 = visit_law(dg.walk_alias(t, starts, 40, seed=1)[0])
    fa2 = visit_law(dg.walk_alias(t, starts, 40, seed=3)[0])
    fr = visit_law(dg.walk_reject(0.25, 4.0, starts, 40, seed=2)[0])
    noise = np.abs(fa1 - fa2).sum()           # sampling noise of the statistic: alias vs alias
    assert np.abs(fa1 - fr).sum() < 1.5 * noise and np.abs(fa2 - fr).sum() < 1.5 * noise


def alias_table_probs(J, q):
    """the distribution an alias table (J, q) samples: p[k] = (q[k] + sum_{j: J[j]=k} (1 - q[j])) / K"""
    K = len(J)
    pr = np.minimum(q, 1.0).copy()
    np.add.at(pr, J, 1.0 - np.minimum(q, 1.0))
    return pr / K


@pytest.mark.parametrize("weighted,p", [(False, 0.25), (True, 0.5), (False, 4.0)])
@pytest.mark.parametrize("indexed", [True, False])
def test_popularity_rejection_walker_chi_square(weighted, p, indexed):
    """popwalk="pop" without edge tables (the graphs run_all_ue_pop.sh targets do not fit them):
    first step ~ get_alias_nodes_cur (node2vec.py:13-25, item rows plain), later steps ~
    get_alias_edge_pop (:154-174) -- chi-square against the oracle's exact popularity tables."""
    n = 300
    _, g = random_graph(n, 3000, seed=37, weighted=weighted, skew=1.0)
    is_item = (np.arange(n) % 3 == 0).astype(np.uint8)           # "9999999"-prefixed labels
    dg = dev_graph(g, symmetric=True, is_item=is_item)
    hub = int(np.argmax(np.diff(g.row_ptr)))
    starts_np = np.concatenate([np.full(400000, hub, dtype=np.int32), np.repeat(np.arange(n, dtype=np.int32), 2000)])
    first = dg.build_node_tables(popwalk=True)
    walks, lens = dg.walk_reject(p, 4.0, torch.as_tensor(starts_np), 3, seed=12, indexed=indexed,
                                 first_tables=first, pop_edges=True)
    w = walks.cpu().numpy()
    # first step: popularity node law (w / deg(nbr) unless the node is an item)
    to = oracle.preprocess(g, p, 4.0, is_item=is_item, popwalk_nodes=True)
    for v in (hub, 3, 4):
        a, b = g.row_ptr[v], g.row_ptr[v + 1]
        if b - a < 2:
            continue
        row = g.col[a:b]
        sel = w[:, 0] == v
        obs = np.bincount(np.searchsorted(row, w[sel, 1]), minlength=len(row))
        assert chi_square_p(obs, alias_table_probs(to.nJ[a:b], to.nq[a:b])) > 1e-4
    # later steps: get_alias_edge_pop rows
    keys, c = transition_counts(walks, lens, n)
    a_, b_, nxt = keys // (n * n), (keys // n) % n, keys % n
    pair = a_ * n + b_
    sums = {}
    for pk, cc in zip(pair, c):
        sums[pk] = sums.get(pk, 0) + cc
    pvals = []
    for pk in sorted(sums, key=lambda k: -sums[k])[:40]:
        aa, bb = int(pk // n), int(pk % n)
        J, qq = oracle.edge_table(g, p, 4.0, aa, bb, popwalk=True)
        row = g.col[g.row_ptr[bb]:g.row_ptr[bb + 1]]
        obs = np.zeros(len(row))
        sel = pair == pk
        obs[np.searchsorted(row, nxt[sel])] = c[sel]
        pvals.append(chi_square_p(obs, alias_table_probs(J, qq)))
    pvals = np.asarray(pvals)
    assert pvals.min() > 1e-5 and (pvals < 0.01).sum() <= 3, np.sort(pvals)[:6]


@pytest.mark.parametrize("weighted", [False, True])
def test_popularity_preprocess_rejection_first_step_then_plain_law(weighted):
    """preprocess_transition_probs_popularity + rejection mode (node2vec.py:206-237): the first step
    follows the popularity node table, later steps the PLAIN get_alias_edge law (:228-232)."""
    n = 300
    _, g = random_graph(n, 3000, seed=39, weighted=weighted, skew=1.0)
    is_item = (np.arange(n) % 2 == 0).astype(np.uint8)
    dg = dev_graph(g, symmetric=True, is_item=is_item)
    hub = int(np.argmax(np.diff(g.row_ptr)))
    hub = hub if not is_item[hub] else int(np.argsort(np.diff(g.row_ptr) * (1 - is_item))[-1])
    starts_np = np.concatenate([np.full(400000, hub, dtype=np.int32), np.repeat(np.arange(n, dtype=np.int32), 1000)])
    first = dg.build_node_tables(popwalk=True)
    walks, lens = dg.walk_reject(0.25, 4.0, torch.as_tensor(starts_np), 3, seed=13, first_tables=first)
    w = walks.cpu().numpy()
    to = oracle.preprocess(g, 0.25, 4.0, is_item=is_item, popwalk_nodes=True)
    a, b = g.row_ptr[hub], g.row_ptr[hub + 1]
    row = g.col[a:b]
    sel = w[:, 0] == hub
    obs = np.bincount(np.searchsorted(row, w[sel, 1]), minlength=len(row))
    assert chi_square_p(obs, alias_table_probs(to.nJ[a:b], to.nq[a:b])) > 1e-4
    plain = oracle.transition_row(g, 1.0, 1.0, -1, hub)
    assert np.abs(plain - alias_table_probs(to.nJ[a:b], to.nq[a:b])).max() > 1e-3     # the two laws differ here
    keys, c = transition_counts(walks, lens, n)
    a_, b_, nxt = keys // (n * n), (keys // n) % n, keys % n
    pair = a_ * n + b_
    sums = {}
    for pk, cc in zip(pair, c):
        sums[pk] = sums.get(pk, 0) + cc
    pvals = []
    for pk in sorted(sums, key=lambda k: -sums[k])[:30]:
        aa, bb = int(pk // n), int(pk % n)
        probs = oracle.transition_row(g, 0.25, 4.0, aa, bb)
        row = g.col[g.row_ptr[bb]:g.row_ptr[bb + 1]]
        obs = np.zeros(len(row))
        sel = pair == pk
        obs[np.searchsorted(row, nxt[sel])] = c[sel]
        pvals.append(chi_square_p(obs, probs))
    pvals = np.asarray(pvals)
    assert pvals.min() > 1e-5 and (pvals < 0.01).sum() <= 3, np.sort(pvals)[:6]

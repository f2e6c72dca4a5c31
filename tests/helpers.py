"""Shared test helpers: seeded synthetic graphs as oracle.CSR, golden loading."""
import glob
import os

import numpy as np

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = oracle.CSR(z["row_ptr"], z["col"], z["w"] if int(z["weighted"]) else None)
    return z, g


def random_graph(n, m, seed, weighted=False, directed=False, dyadic_weights=False, skew=0.0):
    """m random arcs/edges over n nodes (duplicates collapse); skew>0 gives a heavy-tailed degree
    distribution (endpoint ~ floor(n * u^(1+skew)))."""
    rng = np.random.RandomState(seed)
    a = np.floor(n * rng.rand(m) ** (1.0 + skew)).astype(np.int64)
    b = np.floor(n * rng.rand(m) ** (1.0 + skew)).astype(np.int64)
    keep = a != b
    a, b = a[keep], b[keep]
    w = None
    if weighted:
        w = (rng.randint(1, 9, size=a.shape[0]).astype(np.float64) / 4.0 if dyadic_weights
             else np.round(rng.rand(a.shape[0]) * 5 + 0.05, 3))
    return (a, b, w), oracle.csr_from_coo(a, b, w, n, undirected=not directed)


def chi_square_p(obs, probs):
    """p-value of Pearson's chi-square of counts `obs` against `probs` (cells with tiny
    expectation merged)."""
    from scipy import stats
    obs = np.asarray(obs, dtype=np.float64)
    exp = np.asarray(probs, dtype=np.float64) * obs.sum()
    order = np.argsort(exp)
    obs, exp = obs[order], exp[order]
    # merge the smallest cells until each has expectation >= 5
    k = 0
    while k < len(exp) - 1 and exp[:k + 1].sum() < 5:
        k += 1
    obs = np.concatenate([[obs[:k + 1].sum()], obs[k + 1:]])
    exp = np.concatenate([[exp[:k + 1].sum()], exp[k + 1:]])
    if len(exp) < 2:
        return 1.0
    chi2 = ((obs - exp) ** 2 / exp).sum()
    return float(stats.chi2.sf(chi2, len(exp) - 1))


# ---- link-prediction protocol of main_link.main (src/main_link.py:519-565), restated -----------
def chung_lu_graph(n, m, seed, gamma=0.75, max_deg=None, communities=1, mu_in=0.8):
    """Heavy-tailed simple undirected graph ("BlogCatalog-shaped"): endpoints
    drawn with weights ~ (i+10)^-gamma (capped so the largest expected degree is ~max_deg),
    self-loops/duplicates dropped until m distinct edges. communities > 1 plants a partition
    (node i in community i % communities; an edge stays inside the first endpoint's community
    with probability mu_in) so that held-out links are predictable from structure -- a pure
    Chung-Lu graph has none beyond degree, which cosine scores ignore."""
    rng = np.random.RandomState(seed)
    wts = (np.arange(n) + 10.0) ** (-gamma)
    if max_deg:
        wts = np.minimum(wts, wts.sum() * max_deg / (2.0 * m))
    cdf = np.cumsum(wts / wts.sum())
    comm_nodes = [np.arange(c, n, communities) for c in range(communities)]
    comm_cdf = [np.cumsum(wts[ix] / wts[ix].sum()) for ix in comm_nodes]
    edges = set()
    while len(edges) < m:
        k = int((m - len(edges)) * 1.3) + 16
        a = np.searchsorted(cdf, rng.rand(k))
        b = np.searchsorted(cdf, rng.rand(k))
        if communities > 1:
            inside = rng.rand(k) < mu_in
            u = rng.rand(k)
            for i in np.nonzero(inside)[0]:
                c = a[i] % communities
                b[i] = comm_nodes[c][min(np.searchsorted(comm_cdf[c], u[i]), len(comm_nodes[c]) - 1)]
        a = np.minimum(a, n - 1); b = np.minimum(b, n - 1)
        for x, y in zip(a.tolist(), b.tolist()):
            if x != y:
                edges.add((min(x, y), max(x, y)))
                if len(edges) >= m:
                    break
    return np.asarray(sorted(edges), dtype=np.int64)


def split_edges(edges, seed=123, test_ratio=0.5):
    """train_test_split(np.asarray(all_edges), test_size=0.5, random_state=123) (main_link.py:526)"""
    from sklearn.model_selection import train_test_split
    tr, te = train_test_split(np.asarray(edges), test_size=test_ratio, random_state=seed)
    return tr, te


def build_neg_samples(n, true_edges, count, seed):
    """build_neg_samples (main_link.py:191-204): random node pairs that are not edges"""
    rng = np.random.RandomState(seed)
    true = set(map(tuple, np.sort(np.asarray(true_edges), axis=1).tolist()))
    out = set()
    while len(out) < count:
        a, b = rng.randint(0, n, size=2)
        if a == b:
            continue
        e = (min(a, b), max(a, b))
        if e in true or e in out:
            continue
        out.add(e)
    return np.asarray(sorted(out), dtype=np.int64)


def roc_auc_cosine(emb, pos, neg):
    """get_roc_score with link_method 'cos' (main_link.py:43-49,:173-189); emb float32[N, d]
    indexed by node id (rows of nodes never seen may be zero -> score 0, as link_score's except)."""
    from sklearn.metrics import roc_auc_score
    nrm = np.linalg.norm(emb, axis=1, keepdims=True)
    e = emb / np.where(nrm > 0, nrm, 1.0)
    sp = (e[pos[:, 0]] * e[pos[:, 1]]).sum(1)
    sn = (e[neg[:, 0]] * e[neg[:, 1]]).sum(1)
    return float(roc_auc_score(np.concatenate([np.ones(len(sp)), np.zeros(len(sn))]),
                               np.concatenate([sp, sn])))

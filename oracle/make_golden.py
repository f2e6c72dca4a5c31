"""Generate tests/golden/*.npz|json by RUNNING THE REFERENCE ITSELF (build container only).

    python -m oracle.make_golden

Imports /root/reference/src/node2vec.py unmodified through oracle/ref_loader.py, runs its
alias_setup / preprocess_transition_probs[_popularity] / simulate_walks[_on_the_fly] on small
graphs, with Philox uniforms injected into np.random.rand for the walks, and stores inputs and
outputs in CSR arc order. The committed vectors are what pins oracle/n2v_oracle.c (not-gpu
tests) and, through it and directly, the CUDA path (gpu tests). The reference ships no tests or
golden files of its own; its only fixture is graph/karate.edgelist.
"""
from __future__ import annotations

import hashlib
import json
import os

import networkx as nx
import numpy as np

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KARATE = "/root/reference/graph/karate.edgelist"


def read_graph(path, weighted=False, directed=False):
    """main.py:66-80 restated (it reads module-global args, so it cannot be called directly)."""
    if weighted:
        G = nx.read_edgelist(path, nodetype=int, data=(("weight", float),), create_using=nx.DiGraph())
    else:
        G = nx.read_edgelist(path, nodetype=int, create_using=nx.DiGraph())
        for e in G.edges():
            G[e[0]][e[1]]["weight"] = 1
    if not directed:
        G = G.to_undirected()
    return G


def nx_to_csr(G):
    labels = sorted(G.nodes())
    idx = {l: i for i, l in enumerate(labels)}
    row_ptr = np.zeros(len(labels) + 1, dtype=np.int64)
    col, w = [], []
    for i, l in enumerate(labels):
        nb = sorted(G.neighbors(l))
        row_ptr[i + 1] = row_ptr[i] + len(nb)
        col.extend(idx[x] for x in nb)
        w.extend(float(G[l][x]["weight"]) for x in nb)
    return labels, idx, row_ptr, np.asarray(col, dtype=np.int32), np.asarray(w, dtype=np.float64)


def tables_in_arc_order(G, refG, labels, idx, row_ptr, col):
    nnz = int(row_ptr[-1])
    nJ = np.zeros(nnz, dtype=np.int64)
    nq = np.zeros(nnz, dtype=np.float64)
    for i, l in enumerate(labels):
        J, q = refG.alias_nodes[l]
        nJ[row_ptr[i]:row_ptr[i + 1]] = J
        nq[row_ptr[i]:row_ptr[i + 1]] = q
    etab = np.zeros(nnz + 1, dtype=np.int64)
    eJ, eq = [], []
    for i, l in enumerate(labels):
        for e in range(row_ptr[i], row_ptr[i + 1]):
            v = labels[col[e]]
            J, q = refG.alias_edges[(l, v)]
            etab[e + 1] = etab[e] + len(J)
            eJ.append(np.asarray(J, dtype=np.int64))
            eq.append(np.asarray(q, dtype=np.float64))
    eJ = np.concatenate(eJ) if eJ else np.zeros(0, dtype=np.int64)
    eq = np.concatenate(eq) if eq else np.zeros(0, dtype=np.float64)
    return nJ, nq, etab, eJ, eq


def pad_walks(walks, idx, L):
    out = np.full((len(walks), L), -1, dtype=np.int32)
    lens = np.zeros(len(walks), dtype=np.int32)
    for i, wk in enumerate(walks):
        lens[i] = len(wk)
        out[i, :len(wk)] = [idx[x] for x in wk]
    return out, lens


def graph_case(name, G, directed, p, q, R, L, seed, popwalk="none"):
    """Run the reference on G; write tests/golden/<name>.npz"""
    ref = ref_loader.load(naive_sum=True)
    labels, idx, row_ptr, col, w = nx_to_csr(G)
    refG = ref.Graph(G, directed, p, q, popwalk)
    if popwalk == "pop":
        refG.preprocess_transition_probs_popularity()
    else:
        refG.preprocess_transition_probs()
    nJ, nq, etab, eJ, eq = tables_in_arc_order(G, refG, labels, idx, row_ptr, col)
    order = [idx[x] for x in G.nodes()]
    with ref_loader.injected(refG, seed):
        walks = refG.simulate_walks(R, L)
    wk, lens = pad_walks(walks, idx, L)
    with ref_loader.injected(refG, seed):
        walks2 = refG.simulate_walks_on_the_fly(R, L)
    wk2, lens2 = pad_walks(walks2, idx, L)
    # a sub-list of start nodes with a walk-id base (the multi-process/multi-GPU shard shape,
    # main_link.py:263-264: contiguous chunks of list(G.nodes()))
    sub = list(G.nodes())[len(order) // 3: 2 * len(order) // 3]
    with ref_loader.injected(refG, seed, walk_id_base=1000):
        walks3 = refG.simulate_walks(2, L, nodes=sub)
    wk3, lens3 = pad_walks(walks3, idx, L)
    is_item = np.asarray([str(l).startswith("9999999") for l in labels], dtype=np.uint8)
    # same tables with the builtin (compensated) sum, to document the py2/py3 difference
    ref_b = ref_loader.load(naive_sum=False)
    refB = ref_b.Graph(G, directed, p, q, popwalk)
    (refB.preprocess_transition_probs_popularity if popwalk == "pop" else refB.preprocess_transition_probs)()
    _, nqb, _, eJb, eqb = tables_in_arc_order(G, refB, labels, idx, row_ptr, col)
    info = dict(max_abs_dq_builtin_sum=float(np.abs(eqb - eq).max()) if eq.size else 0.0,
                J_equal_builtin_sum=bool((eJb == eJ).all()))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        labels=np.asarray(labels, dtype=np.int64), row_ptr=row_ptr, col=col, w=w,
        weighted=np.asarray(int(not np.all(w == 1.0))), directed=np.asarray(int(directed)),
        p=np.asarray(float(p)), q=np.asarray(float(q)), R=np.asarray(R), L=np.asarray(L),
        seed=np.asarray(seed), order=np.asarray(order, dtype=np.int32), is_item=is_item,
        popwalk=np.asarray(int(popwalk == "pop")),
        nJ=nJ, nq=nq, etab_ptr=etab, eJ=eJ, eq=eq, walks=wk, lens=lens,
        walks_otf=wk2, lens_otf=lens2, sub_starts=np.asarray([idx[x] for x in sub], dtype=np.int32),
        walks_sub=wk3, lens_sub=lens3)
    same = bool((wk == wk2).all())
    print(f"{name}: N={len(labels)} nnz={int(row_ptr[-1])} sum_deg_sq={eq.size} walks={wk.shape} "
          f"otf==pre:{same} {info}")
    return info


def main():
    assert ref_loader.available(), "the reference is only present in the build container"
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load(naive_sum=True)
    rng = np.random.RandomState(20261018)

    # ---- alias_setup known answers (node2vec.py:240-269) ------------------------------
    cases = [[0.5, 0.5], [0.1, 0.2, 0.7], [0.4, 0.1, 0.1, 0.4], [0.05, 0.05, 0.9], [1.0]]
    for K in (3, 7, 49, 66, 100):
        cases.append([1.0 / K] * K)
    for K in (2, 5, 17, 64, 257, 1000):
        x = rng.rand(K) ** 3
        cases.append(list(x / x.sum()))
    for K in (8, 33):                       # dyadic: every intermediate is exact
        x = rng.randint(1, 9, size=K).astype(np.float64)
        x = x / 2.0 ** np.ceil(np.log2(x.sum()))
        x[-1] += 1.0 - x.sum()
        cases.append(list(x))
    out = []
    for probs in cases:
        J, q = ref.alias_setup(probs)
        out.append(dict(probs=[float(v) for v in probs], J=[int(v) for v in J],
                        q=[float(v) for v in q]))
    with open(os.path.join(OUT, "alias_setup.json"), "w") as f:
        json.dump(out, f)
    print("alias_setup cases:", len(out))

    # ---- oracle-loader regression anchors -----------------------
    G = read_graph(KARATE)
    sha = {}
    for p, q in ((1, 1), (0.25, 4)):
        refG = ref.Graph(G, False, p, q)
        refG.preprocess_transition_probs()
        np.random.seed(0)
        walks = refG.simulate_walks(10, 80)
        txt = "\n".join(" ".join(map(str, wk)) for wk in walks)
        sha[f"p{p}_q{q}"] = dict(sha256=hashlib.sha256(txt.encode()).hexdigest(),
                                 first=" ".join(map(str, walks[0][:12])))
    with open(os.path.join(OUT, "karate_mt19937_sha256.json"), "w") as f:
        json.dump(sha, f, indent=1)
    print(sha)

    # ---- graph cases ----------------------------------------------------------------
    graph_case("karate_p1_q1", G, False, 1.0, 1.0, 10, 80, seed=1)
    graph_case("karate_p025_q4", G, False, 0.25, 4.0, 10, 80, seed=2)
    graph_case("karate_p4_q025", G, False, 4.0, 0.25, 3, 40, seed=3)

    # undirected, float (non-dyadic) weights, non-contiguous labels, an isolated node
    H = nx.Graph()
    labs = sorted(rng.choice(5000, size=60, replace=False).tolist())
    H.add_nodes_from(rng.permutation(labs).tolist())
    while H.number_of_edges() < 300:
        a, b = rng.choice(labs[:-1], size=2, replace=False)
        H.add_edge(int(a), int(b), weight=float(np.round(rng.rand() * 4 + 0.1, 3)))
    graph_case("rndw_p05_q2", H, False, 0.5, 2.0, 4, 30, seed=4)
    graph_case("rndw_p03_q3", H, False, 0.3, 3.0, 2, 30, seed=5)

    # directed, dead ends (sinks), self-loops, integer weights 1..5
    D = nx.DiGraph()
    D.add_nodes_from(range(100, 150))
    while D.number_of_edges() < 220:
        a, b = rng.randint(100, 142, size=2)     # 142..149 never have out-arcs... as sources
        b = int(rng.randint(100, 150))
        D.add_edge(int(a), b, weight=int(rng.randint(1, 6)))
    D.add_edge(101, 101, weight=2)
    D.add_edge(117, 117, weight=1)
    graph_case("dir_p025_q4", D, True, 0.25, 4.0, 3, 25, seed=6)

    # user/item graph with the '9999999' item-id prefix, popularity walk (next-row f.1)
    B = nx.Graph()
    users = list(range(1, 31))
    items = [int("9999999" + str(i)) for i in range(1, 41)]
    for u in users:
        for it in rng.choice(items, size=rng.randint(2, 9), replace=False):
            B.add_edge(u, int(it), weight=int(rng.randint(1, 6)))
    graph_case("bip_pop_p1_q1", B, False, 1.0, 1.0, 2, 20, seed=7, popwalk="pop")
    graph_case("bip_p025_q4", B, False, 0.25, 4.0, 2, 20, seed=8)


if __name__ == "__main__":
    main()

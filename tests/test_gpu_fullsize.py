"""GPU: the hot path at BASELINE.json's full size (C4: R-MAT scale 22, 100 M undirected edges),
checked through size-independent properties instead of an oracle that could not finish:
CSR sortedness / symmetry / no duplicates, every walk step is an arc, sharding invariance, the two
rejection forms agree token by token, SGNS pair counts agree between the negative-sampling modes and
sit inside the window bounds, tables stay finite."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4():
    from node2vec_by_ecc_b200 import DeviceGraph, synth
    if torch.cuda.mem_get_info()[0] < 60 << 30:
        pytest.skip("needs ~60 GB of free device memory")
    lo, hi, n = synth.rmat_edges(22, 100_000_000, seed=1, device="cuda")
    assert lo.numel() == 100_000_000 and bool((lo < hi).all())
    dg = DeviceGraph.from_coo(lo, hi, None, n, undirected=True)
    del lo, hi
    torch.cuda.empty_cache()
    return dg


def arc_keys(dg):
    deg = dg.row_ptr[1:] - dg.row_ptr[:-1]
    src = torch.repeat_interleave(torch.arange(dg.n, device=dg.device), deg)
    return src * dg.n + dg.col.to(torch.int64)


def test_csr_properties(c4):
    dg = c4
    assert dg.nnz == 200_000_000 and int(dg.row_ptr[0]) == 0 and int(dg.row_ptr[-1]) == dg.nnz
    deg = dg.row_ptr[1:] - dg.row_ptr[:-1]
    assert int(deg.min()) >= 1                      # isolated vertices were dropped by the generator
    keys = arc_keys(dg)
    assert bool((keys[1:] > keys[:-1]).all())       # rows ascending, strictly: sorted and duplicate-free
    src, dst = keys // dg.n, keys % dg.n
    assert bool((src != dst).all())                 # no self-loops
    rev = torch.sort(dst * dg.n + src).values       # symmetric: the reversed arc set is the arc set
    assert torch.equal(rev, keys)
    assert dg.sum_deg_sq() == int((deg.to(torch.float64) ** 2).sum().item())


def test_walks_are_paths_and_sharding_is_invisible(c4):
    dg = c4
    B, L = 1 << 19, 80
    starts = torch.arange(B, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    walks, lens = dg.walk_reject(0.25, 4.0, starts, L, seed=1, walk_id_base=123, counters=cnt)
    assert bool((lens == L).all()) and bool((walks >= 0).all()) and bool((walks[:, 0] == starts).all())
    keys = arc_keys(dg)
    a, b = walks[:, :-1].reshape(-1).to(torch.int64), walks[:, 1:].reshape(-1).to(torch.int64)
    q = a * dg.n + b
    pos = torch.searchsorted(keys, q).clamp_(max=keys.numel() - 1)
    assert bool((keys[pos] == q).all())             # every step follows an arc
    c = cnt.cpu().numpy()
    assert c[0] == B * (L - 1) and c[1] >= c[0] and c[2] <= c[1]
    # any split of the start list gives the same corpus (global walk ids)
    h = B // 3
    w1, _ = dg.walk_reject(0.25, 4.0, starts[:h], L, seed=1, walk_id_base=123)
    w2, _ = dg.walk_reject(0.25, 4.0, starts[h:], L, seed=1, walk_id_base=123 + h)
    assert torch.equal(walks, torch.cat([w1, w2]))
    # binary-search form == hashed state-machine form
    sub = starts[: 1 << 16]
    wa, _ = dg.walk_reject(0.25, 4.0, sub, L, seed=1, walk_id_base=123, indexed=False)
    assert torch.equal(wa, walks[: 1 << 16])
    # p = q = 1: first-order walk, exactly one trial per step, no distance-1 tests
    cnt.zero_()
    dg.walk_reject(1.0, 1.0, sub, L, seed=2, counters=cnt)
    c = cnt.cpu().numpy()
    assert c[1] == c[0] and c[2] <= c[0] * 1e-6


def test_sgns_pair_counts_and_finiteness(c4):
    from node2vec_by_ecc_b200 import SgnsTrainer
    dg = c4
    B, L, W = 1 << 17, 80, 10
    starts = torch.arange(B, dtype=torch.int32, device="cuda")
    walks, lens = dg.walk_reject(0.25, 4.0, starts, L, seed=1)
    counts = torch.bincount(walks.reshape(-1).to(torch.int64), minlength=dg.n)
    res = {}
    for shared in (1, 0):
        tr = SgnsTrainer(counts, dim=128, window=W, negative=5, sample=0.0, seed=1)
        tr.sample = 0.0
        before = tr.syn0.clone()
        tr.train(walks, None, B, L, total_examples=B, sent_per_job=125, negative_sharing=shared)
        torch.cuda.synchronize()
        p = tr.pairs.cpu().numpy()
        res[shared] = int(p[0])
        assert bool(torch.isfinite(tr.syn0).all()) and bool(torch.isfinite(tr.syn1neg).all())
        assert float((tr.syn0 - before).abs().max()) > 0 and float(tr.syn1neg.abs().max()) > 0
        if shared:
            assert int(p[1]) == B * L               # every kept position is a centre with >= 1 context
    # window shrink is keyed by (sentence, position): both modes see the same windows
    assert res[0] == res[1]
    # per walk: between the fully shrunk window (2 per inner position) and L*(2W) - W(W+1) pairs
    assert B * (2 * L - 2) <= res[1] <= B * (2 * W * L - W * (W + 1))

"""GPU: out-of-bounds WRITE guards (compute-sanitizer is closed on this pool): every output buffer
is carved out of a larger allocation filled with a sentinel; after the kernels run the sentinels on
both sides must be intact. Odd sizes on purpose (partial sectors, non-multiples of 8/32)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import random_graph

pytestmark = pytest.mark.gpu
PAD = 1024


def guarded(shape, dtype, fill):
    n = int(np.prod(shape))
    big = torch.full((n + 2 * PAD,), fill, dtype=dtype, device="cuda")
    return big, big[PAD:PAD + n].view(*shape)


def intact(big, n, fill):
    return bool((big[:PAD] == fill).all()) and bool((big[PAD + n:] == fill).all())


@pytest.mark.parametrize("L", [1, 5, 8, 13, 80])
@pytest.mark.parametrize("n_walks", [1, 31, 257])
def test_walk_kernels_write_inside_their_buffers(L, n_walks):
    from node2vec_by_ecc_b200 import DeviceGraph
    _, g = random_graph(300, 1500, seed=2, directed=True, weighted=True, skew=0.7)
    dg = DeviceGraph.from_csr(g.row_ptr, g.col, g.w, symmetric=False)
    t = dg.build_alias_tables(0.5, 2.0)
    starts = torch.arange(n_walks, dtype=torch.int32) % g.n
    for kind in ("alias_packed", "alias", "reject_ix", "reject"):
        bw, walks = guarded((n_walks, L), torch.int32, -77)
        bl, lens = guarded((n_walks,), torch.int32, -77)
        if kind.startswith("alias"):
            dg.walk_alias(t, starts, L, 1, 3, out=(walks, lens), packed=kind == "alias_packed")
        else:
            dg.walk_reject(0.5, 2.0, starts, L, 1, 3, out=(walks, lens), indexed=kind == "reject_ix")
        torch.cuda.synchronize()
        assert intact(bw, n_walks * L, -77) and intact(bl, n_walks, -77), kind
        assert bool((walks != -77).all()) and bool((lens >= 1).all()) and bool((lens <= L).all()), kind


def test_alias_build_and_sgns_write_inside_their_buffers():
    from node2vec_by_ecc_b200 import DeviceGraph, SgnsTrainer
    from node2vec_by_ecc_b200._lib import check, lib, ptr, stream
    _, g = random_graph(200, 1300, seed=3, skew=0.6)
    dg = DeviceGraph.from_csr(g.row_ptr, g.col, None, symmetric=True)
    etab = dg.etab_offsets()
    tot, nnz = dg.sum_deg_sq(), dg.nnz
    bs, slots = guarded((tot, 2), torch.int32, -5)
    bj, wJ = guarded((tot,), torch.int32, -5)
    bq, wq = guarded((tot,), torch.float64, -5.0)
    check(lib().n2v_alias_build_edges(ptr(dg.row_ptr), ptr(dg.col), None, C.c_int32(dg.n), C.c_double(0.25),
                                      C.c_double(4.0), 1, 0, ptr(etab), C.c_int64(0), C.c_int64(nnz), ptr(slots),
                                      ptr(wJ), ptr(wq), stream()))
    bn, nslots = guarded((nnz, 2), torch.int32, -5)
    bnj, nJ = guarded((nnz,), torch.int32, -5)
    bnq, nq = guarded((nnz,), torch.float64, -5.0)
    check(lib().n2v_alias_build_nodes(ptr(dg.row_ptr), ptr(dg.col), None, C.c_int32(dg.n), None, 0, ptr(nslots),
                                      ptr(nJ), ptr(nq), stream()))
    torch.cuda.synchronize()
    assert intact(bs, tot * 2, -5) and intact(bj, tot, -5) and intact(bq, tot, -5.0)
    assert intact(bn, nnz * 2, -5) and intact(bnj, nnz, -5) and intact(bnq, nnz, -5.0)
    assert bool((wJ >= 0).all()) and bool((nJ >= 0).all())
    # SGNS: rows [V, V + pad) of both tables stay untouched in every kernel variant
    t = dg.build_alias_tables(0.25, 4.0)
    walks, lens = dg.walk_alias(t, torch.arange(dg.n, dtype=torch.int32).repeat(4), 37, 1)
    counts = torch.bincount(walks.reshape(-1).to(torch.int64), minlength=dg.n)
    for dim, neg, shared in ((128, 5, 1), (64, 5, 1), (128, 5, 0), (100, 3, 0), (256, 5, 0)):
        for atomic in (1, 0):
            tr = SgnsTrainer(counts, dim=dim, window=10, negative=neg, sample=1e-3, seed=1)
            b0, s0 = guarded((tr.V, dim), torch.float32, 7.0)
            b1, s1 = guarded((tr.V, dim), torch.float32, 7.0)
            s0.copy_(tr.syn0); s1.copy_(tr.syn1neg)
            tr.syn0, tr.syn1neg = s0, s1
            tr.train(walks, None, walks.shape[0], 37, total_examples=walks.shape[0], sent_per_job=100,
                     negative_sharing=shared, atomic_updates=atomic, grid_warps=64)
            torch.cuda.synchronize()
            assert intact(b0, tr.V * dim, 7.0) and intact(b1, tr.V * dim, 7.0), (dim, neg, shared, atomic)
            assert int(tr.pairs[0]) > 0 and bool(torch.isfinite(s0).all())


def test_argument_guards_of_the_block_and_packed_entry_points():
    """round-1 advisor findings: dim % 4 in the block kernel, window <= 96 in the pair expansion,
    arc records required from the first step of the packed alias walker"""
    import pytest
    from node2vec_by_ecc_b200 import BlockSgnsTrainer
    from node2vec_by_ecc_b200._lib import N2VError, SgnsParams, check, lib, ptr, stream
    counts = torch.arange(1, 65, dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        BlockSgnsTrainer(counts, dim=50, local_parts=2)
    tr = BlockSgnsTrainer(counts, dim=64, local_parts=2)
    P = tr._params(0, 4)
    P.dim = 50
    words = torch.zeros(16, dtype=torch.int32, device="cuda")
    with pytest.raises(N2VError, match="multiple of 4"):
        check(lib().n2v_sgns_train_groups(ptr(words), C.c_int64(0), C.c_int64(16), None, None, C.c_int64(16),
                                          C.byref(P), C.c_int32(1), ptr(tr.parts0[0]),
                                          ptr(tr.parts1[0]), C.c_int32(0), C.c_int32(2), ptr(tr.pairs), stream()))
    P = tr._params(0, 4)
    P.window = 113                     # used to spin forever on sentences with > 256 kept tokens
    tok = torch.zeros((2, 400), dtype=torch.int32, device="cuda")
    off = torch.zeros(2 * 2 + 1, dtype=torch.int64, device="cuda")
    ws = torch.empty(int(lib().n2v_sgns_groups_workspace_bytes(C.c_int64(2), C.c_int32(2))), dtype=torch.uint8, device="cuda")
    with pytest.raises(N2VError, match="window"):
        check(lib().n2v_sgns_groups_count(ptr(tok), None, C.c_int64(2), C.c_int32(400), C.c_int64(0), None, None, C.byref(P),
                                         C.c_int32(0), C.c_int32(2), ptr(off), ptr(ws), C.c_size_t(ws.numel()), stream()))
    dummy = torch.zeros(64, dtype=torch.int64, device="cuda")
    w = torch.zeros((4, 2), dtype=torch.int32, device="cuda")
    with pytest.raises(N2VError, match="arc records"):
        check(lib().n2v_walk_alias_packed(ptr(dummy), ptr(dummy), None, None, ptr(w), C.c_int64(4), C.c_int32(2),
                                          C.c_uint64(1), C.c_uint64(0), ptr(w), ptr(w), stream()))

"""CPU: the C-ABI library loads and exports every symbol include/n2v_b200.h declares (no compute
without a GPU: compute calls must fail loudly), host-side logic, and the N>1 sharding/averaging
logic under gloo with world_size 2."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from node2vec_by_ecc_b200 import SO_PATH, lib
    from node2vec_by_ecc_b200._lib import EXPORTS
    hdr = open(os.path.join(ROOT, "include", "n2v_b200.h")).read()
    declared = set(re.findall(r"\b(n2v_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    L = lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.n2v_version() >= 100
    assert os.path.dirname(SO_PATH).endswith("node2vec_by_ecc_b200")      # built in-tree


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from node2vec_by_ecc_b200 import DeviceGraph, Graph, N2VError, Word2Vec, lib
    import networkx as nx
    with pytest.raises(N2VError):
        DeviceGraph.from_coo([0], [1], None, 2, undirected=True)
    G = Graph(nx.path_graph(4), False, 1, 1)
    with pytest.raises(N2VError):
        G.preprocess_transition_probs()
    with pytest.raises(N2VError):
        Word2Vec([["a", "b"]], sg=1, min_count=0)
    # the C entry points themselves refuse without a device
    assert lib().n2v_sm_count() < 0
    rc = lib().n2v_walk_alias(*([ctypes.c_void_p(8)] * 6), ctypes.c_int64(1), ctypes.c_int32(4),
                              ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_void_p(8), ctypes.c_void_p(8),
                              ctypes.c_void_p(0))
    assert rc < 0 and b"CUDA" in lib().n2v_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "node2vec_by_ecc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), f
                assert "libn2v_oracle" not in src, f


def test_line_sentence_and_keyedvectors_host_logic(tmp_path):
    from node2vec_by_ecc_b200 import KeyedVectors, LineSentence, Vocab
    p = tmp_path / "w.txt"
    p.write_text("1 2 3\n\n4 5\n" + " ".join(str(i) for i in range(25)) + "\n")
    s = list(LineSentence(str(p), max_sentence_length=10))
    assert s[0] == ["1", "2", "3"] and s[1] == ["4", "5"]      # empty lines yield nothing (gensim)
    assert [len(x) for x in s[2:]] == [10, 10, 5]
    assert list(LineSentence(str(p), limit=1)) == [["1", "2", "3"]]
    kv = KeyedVectors(4)
    kv.index2word = ["a", "b"]
    kv.vocab = {"a": Vocab(0, 5), "b": Vocab(1, 3)}
    kv.syn0 = np.asarray([[1, 0, 0, 0], [1, 1, 0, 0]], dtype=np.float32)
    assert abs(kv.similarity("a", "b") - 2 ** -0.5) < 1e-6 and kv["b"].tolist() == [1, 1, 0, 0]
    assert kv[["a", "b"]].shape == (2, 4) and "a" in kv and "z" not in kv
    with pytest.raises(KeyError):
        kv["z"]
    out = tmp_path / "e.txt"
    kv.save_word2vec_format(str(out))
    assert out.read_text().splitlines()[0] == "2 4"


def test_walk_corpus_sequence_view():
    from node2vec_by_ecc_b200 import WalkCorpus
    labels = np.asarray([10, 20, 30, 40])
    w = torch.tensor([[0, 1, 2], [3, -1, -1]], dtype=torch.int32)
    c = WalkCorpus(w, torch.tensor([3, 1], dtype=torch.int32), labels)
    assert len(c) == 2 and c[0] == [10, 20, 30] and c[1] == [40] and c[-1] == [40]
    assert [list(map(str, x)) for x in c] == [["10", "20", "30"], ["40"]]
    c.extend(WalkCorpus(w[:1], torch.tensor([2], dtype=torch.int32), labels))
    assert len(c) == 3 and c[2] == [10, 20] and c.num_steps() == 3


def test_sync_interval_rule():
    from node2vec_by_ecc_b200.dist import sync_walks_per_rank
    assert sync_walks_per_rank(10 ** 6, 1, 400.0) > 10 ** 12          # one replica: never
    # total pairs per sync <= 100 V / world  (the emulated +-0.005 AUC band)
    for w in (2, 4, 8):
        walks = sync_walks_per_rank(10 ** 6, w, 396.0)
        assert walks * w * 396.0 <= 100 * 10 ** 6 / w * 1.01
    assert sync_walks_per_rank(1000, 8, 836.0) == 256                 # floor


def test_shard_range_matches_main_link_partition():
    from node2vec_by_ecc_b200.dist import shard_range, step_walk_ids
    import math
    for total in (0, 1, 7, 34, 100, 101):
        for world in (1, 2, 3, 6, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            # main_link.py:263-264: split_point = range(0, len, ceil(len/num_pool)) + [len]
            covered = [i for a, b in parts for i in range(a, b)]
            assert covered == list(range(total))
            if total:
                per = int(math.ceil(float(total) / world))
                assert all(b - a == per for a, b in parts if b < total)
    ids = sorted(step_walk_ids(s, r, 4, 8) for s in range(3) for r in range(4))
    assert ids == list(range(0, 96, 8))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from node2vec_by_ecc_b200.dist import ReplicaSync, average_tables, shard_range, sum_counts
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    a, b = shard_range(10, rank, world)
    counts = torch.zeros(10, dtype=torch.int64)
    counts[a:b] = torch.arange(a, b) + 1
    sum_counts(counts)
    t0 = torch.full((4, 8), float(rank + 1))
    t1 = torch.full((4, 8), float(10 * (rank + 1)))
    average_tables(t0, t1)
    # delta-sum: every rank starts from the same base and adds its own delta
    base = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    rep = base.clone()
    sync = ReplicaSync(rep)
    rep += float(rank + 1)                       # this replica's training moved every entry by rank+1
    sync.sync()
    ok1 = bool(torch.equal(rep, base + 3.0))     # base + (1 + 2)
    rep[rank] += 10.0                            # second interval: disjoint rows
    sync.sync()
    want = base + 3.0
    want[0] += 10.0; want[1] += 10.0
    ok2 = bool(torch.equal(rep, want)) and bool(torch.equal(sync.bases[0], want))
    q.put((rank, counts.tolist(), float(t0[0, 0]), float(t1[0, 0]), ok1 and ok2))
    dist.destroy_process_group()


def test_gloo_world2_counts_and_table_averaging():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, counts, a0, a1, delta_ok in res:
        assert counts == list(range(1, 11))          # disjoint shards summed
        assert a0 == 1.5 and a1 == 15.0              # (1+2)/2, (10+20)/2
        assert delta_ok                              # ReplicaSync: base + sum of deltas on every rank


def test_bench_reference_arm_contract_small():
    """bench.py --impl reference (the oracle port on host cores) prints ONE JSON line with the
    contract's keys; tiny R-MAT so that it runs in seconds on CPU."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "10",
                          "--edges", "5000", "--steps", "1", "--warmup", "0", "--ref-walks", "8"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "pairs/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def _gloo_block_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from node2vec_by_ecc_b200.dist import bucket_of, gather_pool, ring_pass
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pool = gather_pool(torch.full((3, 4), rank, dtype=torch.int32))
    ok = pool.shape == (3 * world, 4) and all(int(pool[3 * r + i, 0]) == r for r in range(world) for i in range(3))
    held, spare = torch.full((5, 2), float(rank)), torch.empty((5, 2))
    seen = []
    for e in range(world):                       # sub-step e: this rank must hold part (rank + e) % world
        seen.append(int(held[0, 0]))
        ok = ok and int(held[0, 0]) == bucket_of(rank, e, world)
        held, spare = ring_pass(held, spare)
    ok = ok and int(held[0, 0]) == rank          # home again after `world` passes
    q.put((rank, ok, seen))
    dist.destroy_process_group()


def _gloo_plan_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from node2vec_by_ecc_b200.dist import pool_plan, shard_range, sum_counts
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    total = 1003                                   # walks of the whole corpus; shares: 502 + 501
    lo, hi = shard_range(total, rank, world)
    per, pools = pool_plan(total, world, 200)
    # what Graph(distributed=True).simulate_walks + Word2Vec._train_sharded do with their share
    mine = torch.arange(lo, hi)                    # global walk ids of this rank's share
    padded = torch.cat([mine, torch.full((per - mine.numel(),), -1)])
    seen = []
    for p0, n in pools:
        part = padded[p0:p0 + n]
        got = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(got, part)                 # the pool: rank 0's slice, then rank 1's (gather_pool order)
        seen.append(torch.cat(got))
    counts = sum_counts(torch.bincount(mine % 7, minlength=7))
    q.put((rank, per, pools, torch.cat(seen).tolist(), counts.tolist()))
    dist.destroy_process_group()


def test_gloo_sharded_corpus_pool_plan():
    """the host arithmetic behind Graph(distributed=True) + Word2Vec on a sharded corpus: contiguous shares
    (main_link.py:263-264), padded to one length, the same pools on every rank, every walk in exactly one
    pool, vocabulary counts summed"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gloo_plan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, per0, pools0, seen0, c0), (_, per1, pools1, seen1, c1) = res
    assert per0 == per1 == 502 and pools0 == pools1 == [(0, 200), (200, 200), (400, 102)]
    assert seen0 == seen1                                             # every rank assembles the same pools
    real = [x for x in seen0 if x >= 0]
    assert sorted(real) == list(range(1003)) and len(seen0) == 2 * 502    # one empty walk pads the short share
    assert c0 == c1 == torch.bincount(torch.arange(1003) % 7, minlength=7).tolist()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_block_schedule_ring(world):
    """BlockSgnsTrainer's host plumbing: pool gathered in rank order; in every sub-step the ranks
    hold pairwise different syn0 parts (orthogonal buckets) and every rank sees every part once."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gloo_block_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, seen in res:
        assert ok and sorted(seen) == list(range(world))
    for e in range(world):
        assert sorted(r[2][e] for r in res) == list(range(world))


def parse_groups(words):
    """group stream -> int64[n_pairs, 4] rows {centre local row, context local row, sentence, position}"""
    out, p = [], 0
    while p < len(words):
        assert words[p] & 0x80000000
        c, s, pos, cnt = int(words[p] & 0x7FFFFFFF), int(words[p + 1]), int(words[p + 2] & 0xFFFF), int(words[p + 2] >> 16)
        assert cnt >= 1 and not (words[p + 1: p + 8 + cnt] & 0x80000000).any()
        out += [(c, int(x), s, pos) for x in words[p + 8: p + 8 + cnt]]
        p += 8 + cnt
    return np.asarray(out, dtype=np.int64).reshape(-1, 4)


def test_oracle_block_streams_cover_sentence_major_pairs():
    """oracle self-consistency: the group streams of all (centre part, context part) buckets together
    are exactly the pairs of the sentence-major law under the same Philox addressing, and with ONE
    part the block law IS the sentence-major shared-negative law (same draws, same order, same alpha)."""
    import oracle
    z = np.load(os.path.join(ROOT, "tests", "golden", "karate_p025_q4.npz"))
    w = z["walks"]
    voc = oracle.sgns_vocab(w, 34)
    tok = np.where(w >= 0, voc.id2index[np.maximum(w, 0)], -1).astype(np.int32).ravel()
    off = np.arange(w.shape[0] + 1, dtype=np.int64) * w.shape[1]
    V, dim = voc.V, 16
    s0, s1, pairs = oracle.sgns_train(tok, off, voc, dim=dim, window=10, negative=5, iters=1, workers=1, rng_mode=3, seed=4)
    whole = parse_groups(oracle.sgns_make_groups(tok, off, voc, 0, 1, window=10, seed=4)[0])
    assert len(whole) == pairs
    for n_parts in (2, 4, 8):
        got = []
        for k in range(n_parts):
            for b, st in enumerate(oracle.sgns_make_groups(tok, off, voc, k, n_parts, window=10, seed=4)):
                g = parse_groups(st)
                got.append(np.stack([g[:, 0] * n_parts + k, g[:, 1] * n_parts + b, g[:, 2], g[:, 3]], 1))
        got = np.concatenate(got)
        assert len(got) == pairs
        key = lambda a: np.sort(((a[:, 2] * 100 + a[:, 3]) * 64 + a[:, 0]) * 64 + a[:, 1])
        assert np.array_equal(key(got), key(whole))
    p0 = [oracle.sgns_init_syn0(V, dim, 4)]; p1 = [np.zeros((V, dim), np.float32)]
    n = oracle.sgns_block_pool(tok, off, voc, p0, p1, window=10, alpha=0.025, total_examples=w.shape[0],
                               sent_per_job=10000 // w.shape[1], seed=4)
    assert n == pairs
    assert np.array_equal(p0[0], s0) and np.array_equal(p1[0], s1)
    # several parts: every pair trained once; tables move and stay finite
    for n_parts in (2, 8):
        rows = (V + n_parts - 1) // n_parts
        full0 = oracle.sgns_init_syn0(V, dim, 4)
        q0 = [np.zeros((rows, dim), np.float32) for _ in range(n_parts)]
        q1 = [np.zeros((rows, dim), np.float32) for _ in range(n_parts)]
        for k in range(n_parts):
            q0[k][: len(full0[k::n_parts])] = full0[k::n_parts]
        n = oracle.sgns_block_pool(tok, off, voc, q0, q1, window=10, alpha=0.025, total_examples=w.shape[0],
                                   sent_per_job=10000 // w.shape[1], seed=4)
        assert n == pairs and all(np.isfinite(x).all() for x in q0) and max(np.abs(x).max() for x in q1) > 1e-3


def test_reference_main_py_binds_to_the_dropins_without_a_gpu():
    """src/main.py executed unmodified (runpy) with the drop-ins first on sys.path: argparse, read_graph and
    `import node2vec` / `from gensim.models import Word2Vec` resolve to this package, and without a CUDA
    device the first compute call fails loudly (no CPU fallback) -- the GPU flavour of this test is
    tests/test_gpu_api.py::test_reference_main_py_runs_unmodified_with_the_dropins."""
    import runpy
    import torch
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "main.py")):
        pytest.skip("reference tree not mounted here")
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU flavour")
    from node2vec_by_ecc_b200._lib import N2VError
    argv, cwd = sys.argv, os.getcwd()
    sys.path.insert(0, os.path.join(ROOT, "node2vec_by_ecc_b200", "dropin"))
    for m in ("node2vec", "gensim", "gensim.models", "gensim.models.word2vec"):
        sys.modules.pop(m, None)
    try:
        os.chdir(ref)
        sys.argv = ["main.py", "--input", "graph/karate.edgelist"]
        with pytest.raises(N2VError, match="no CUDA device"):
            runpy.run_path(os.path.join(ref, "src", "main.py"), run_name="__main__")
        import node2vec
        assert node2vec.Graph.__module__ == "node2vec_by_ecc_b200.walker"
    finally:
        os.chdir(cwd)
        sys.argv = argv
        sys.path.pop(0)
        for m in ("node2vec", "gensim", "gensim.models", "gensim.models.word2vec"):
            sys.modules.pop(m, None)


def test_committed_ncu_traffic_matches_the_kernel_sources():
    """profiles/ncu_traffic.json (what bench.py reports as roofline.traffic / frac_dram) was captured from
    exactly the kernel sources in the tree: bench.py refuses a stale entry, this test makes it visible"""
    import json
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        table = json.load(f)
    assert {"sgns_train_kernel_v3", "sgns_train_kernel_v2", "sgns_group_kernel", "walk_reject_indexed_kernel"} <= set(table)
    for kernel, ent in table.items():
        assert ent["source_sha16"] == bench.source_sha16(kernel), "re-capture %s (scripts/capture_traffic.py)" % kernel
        assert os.path.exists(os.path.join(ROOT, ent["profile"])), ent["profile"]
        assert ent["dram_bytes_per_launch"] > 0
    got, src = bench.ncu_traffic("sgns_train_kernel_v3", table["sgns_train_kernel_v3"]["workload"])
    assert got == table["sgns_train_kernel_v3"]["dram_bytes_per_launch"] and src.startswith("profiles/")
    assert bench.ncu_traffic("sgns_train_kernel_v3", "another workload")[0] is None

"""Block-partitioned SGNS (csrc/n2v_sgns_block.cu, word2vec.BlockSgnsTrainer) against the oracle's
restatement of the same schedule (oracle/sgns_oracle.c: sgns_oracle_make_pairs /
sgns_oracle_block_train): pair streams bit-exact, tables within float tolerance when the device runs
one warp (sequential order); all parts on one device == what n GPUs would compute."""
import numpy as np
import pytest
import torch

import oracle
from test_gpu_sgns import corpus_from_golden

pytestmark = pytest.mark.gpu


def oracle_vocab(tr, n_ids):
    order = tr.order.cpu().numpy()
    id2index = np.full(n_ids, -1, dtype=np.int32)
    id2index[order] = np.arange(len(order), dtype=np.int32)
    keep = tr.keep_thr.cpu().numpy().view(np.uint32).astype(np.uint64)
    keep = np.where(keep == 0xFFFFFFFF, np.uint64(1) << np.uint64(32), keep)
    return oracle.Vocab(tr.counts.cpu().numpy(), order.astype(np.int32), id2index, keep,
                        tr.cum_table.cpu().numpy().view(np.uint32).copy()), id2index


def oracle_tokens(walks_np, id2index):
    tok = np.where(walks_np >= 0, id2index[np.maximum(walks_np, 0)], -1).astype(np.int32).ravel()
    off = np.arange(walks_np.shape[0] + 1, dtype=np.int64) * walks_np.shape[1]
    return tok, off


def make_trainer(walks, n_ids, n_parts, dim=128, seed=4, run_pairs=16):
    from node2vec_by_ecc_b200 import BlockSgnsTrainer
    counts = torch.bincount(walks[walks >= 0].to(torch.int64), minlength=n_ids)
    return BlockSgnsTrainer(counts, dim=dim, window=10, negative=5, sample=1e-3, seed=seed, local_parts=n_parts,
                            run_pairs=run_pairs)


@pytest.mark.parametrize("n_parts", [1, 2, 4, 8])
def test_pair_streams_equal_oracle(n_parts):
    z, g, corpus = corpus_from_golden("karate_p025_q4")
    walks = corpus.walks
    tr = make_trainer(walks, g.n, n_parts)
    voc, id2index = oracle_vocab(tr, g.n)
    tok, off = oracle_tokens(z["walks"], id2index)
    P = tr._params(0, 1)
    total = 0
    for k in range(n_parts):
        pairs, bounds = tr.make_pairs(walks, None, walks.shape[0], walks.shape[1], 1000, P, k)
        want = oracle.sgns_make_pairs(tok, off, voc, k, n_parts, window=10, seed=4, epoch=0, sent_id_base=1000)
        got = pairs.cpu().numpy()
        for b in range(n_parts):
            assert bounds[b + 1] - bounds[b] == len(want[b])
            assert np.array_equal(got[bounds[b]:bounds[b + 1]], want[b])
        total += bounds[-1]
    tr.check_overflow()
    # the streams of all parts together are the pairs of the sentence-major trainer (same Philox law)
    from node2vec_by_ecc_b200 import SgnsTrainer
    ref = SgnsTrainer(torch.bincount(walks[walks >= 0].to(torch.int64), minlength=g.n), dim=32, window=10, negative=5,
                      sample=1e-3, seed=4)
    ref.train(walks, None, walks.shape[0], walks.shape[1], total_examples=walks.shape[0], sent_id_base=1000,
              negative_sharing=1)
    assert int(ref.pairs[0]) == total


def test_pair_streams_ragged_long_sentences():
    """sent_off corpus with sentences longer than the staging buffer (streamed in chunks)"""
    rng = np.random.default_rng(5)
    n_ids, lens = 300, rng.integers(0, 900, size=40)
    lens[3] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok_ids = rng.integers(0, n_ids, size=int(off[-1])).astype(np.int32)
    tok_ids[rng.random(tok_ids.shape[0]) < 0.02] = -1
    walks = torch.as_tensor(tok_ids).cuda()
    tr = make_trainer(walks, n_ids, 4, dim=32)
    voc, id2index = oracle_vocab(tr, n_ids)
    tok = np.where(tok_ids >= 0, id2index[np.maximum(tok_ids, 0)], -1).astype(np.int32)
    P = tr._params(2, 1)
    off_d = torch.as_tensor(off).cuda()
    for k in range(4):
        pairs, bounds = tr.make_pairs(walks, off_d, len(lens), 0, 7, P, k)
        want = oracle.sgns_make_pairs(tok, off, voc, k, 4, window=10, seed=4, epoch=2, sent_id_base=7)
        got = pairs.cpu().numpy()
        for b in range(4):
            assert np.array_equal(got[bounds[b]:bounds[b + 1]], want[b])


@pytest.mark.parametrize("n_parts,dim,run_pairs", [(1, 128, 16), (2, 128, 16), (4, 64, 7), (8, 128, 32)])
def test_block_schedule_sequential_equals_oracle(n_parts, dim, run_pairs):
    z, g, corpus = corpus_from_golden("karate_p025_q4")
    walks = corpus.walks
    tr = make_trainer(walks, g.n, n_parts, dim=dim, run_pairs=run_pairs)
    voc, id2index = oracle_vocab(tr, g.n)
    tok, off = oracle_tokens(z["walks"], id2index)
    W, V = n_parts, tr.V
    rows = (V + W - 1) // W
    full0 = oracle.sgns_init_syn0(V, dim, 4)
    parts0 = [np.zeros((rows, dim), np.float32) for _ in range(W)]
    parts1 = [np.zeros((rows, dim), np.float32) for _ in range(W)]
    for k in range(W):
        parts0[k][: len(full0[k::W])] = full0[k::W]
    half = walks.shape[0] // 2
    n_tot = walks.shape[0]
    pairs = 0
    for pool, (a, b) in enumerate([(0, half), (half, n_tot)]):        # two pools: alpha and the tag move on
        tr.train(walks[a:b], None, b - a, walks.shape[1], total_examples=n_tot, example_base=a, sent_id_base=a, grid_warps=1)
        alpha = tr.pool_alpha(a, n_tot)
        pairs += oracle.sgns_block_pool(tok[a * walks.shape[1]: b * walks.shape[1]], off[: b - a + 1], voc, parts0, parts1,
                                        window=10, alpha=alpha, run_pairs=run_pairs, seed=4, epoch=0, sent_id_base=a, pool=pool)
    tr.check_overflow()
    assert int(tr.pairs[0]) == pairs
    s0, s1 = tr.gather()
    want0 = np.zeros((V, dim), np.float32); want1 = np.zeros((V, dim), np.float32)
    for k in range(W):
        n_k = (V - k + W - 1) // W
        want0[k::W] = parts0[k][:n_k]; want1[k::W] = parts1[k][:n_k]
    assert np.abs(s0.cpu().numpy() - want0).max() < 2e-4
    assert np.abs(s1.cpu().numpy() - want1).max() < 2e-4
    assert np.abs(want1).max() > 1e-3


def test_block_wide_close_to_sequential():
    """full Hogwild width: same pairs, embeddings close to the one-warp run (Hogwild noise only)"""
    z, g, corpus = corpus_from_golden("karate_p025_q4")
    walks = corpus.walks
    res = []
    for gw in (1, None):
        tr = make_trainer(walks, g.n, 2, dim=128)
        tr.train(walks, None, walks.shape[0], walks.shape[1], total_examples=walks.shape[0], grid_warps=gw)
        s0, _ = tr.gather()
        assert torch.isfinite(s0).all()
        res.append((int(tr.pairs[0]), s0))
    assert res[0][0] == res[1][0]
    a, b = res[0][1], res[1][1]
    cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
    assert float(cos.mean()) > 0.9


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_one_device():
    """torchrun x 2 (NCCL all-gather + ring): bit-identical to all parts on one device"""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "dist_block_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["max_abs_diff_vs_one_device"] == [0.0, 0.0] and d["moved"] > 1e-3
    assert d["auc_mean"] > 0.78

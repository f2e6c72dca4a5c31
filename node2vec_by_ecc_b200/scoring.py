"""All-pairs similarity selections on the device, host side of
n2v_row_norms / n2v_sim_threshold (csrc/n2v_score.cu): the score matrix of link_prediction
(src/main_link.py:70-171) and of build_user_sim_matrx (:368-376) is computed tile by tile on the fp32
pipes and only the pairs that pass a threshold are ever written. Selections that need an order -- the
global top-k of links_score (:107-111), the per-user top share of get_add_edge_by_ratio (:378-393) --
take their threshold from a sample of the scores (n2v_cosine_pairs) and finish on the few emitted
candidates (a device sort of a short list); a threshold that turns out too high is lowered and the
pass repeated, so the result is exact."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, ptr, require_cuda, stream

NEG_INF = float("-inf")


def _dev_emb(wv, dev):
    m = wv._syn0_dev if getattr(wv, "_syn0_dev", None) is not None else torch.as_tensor(wv.syn0).to(dev)
    return m.float().contiguous()


def row_norms(emb, rows, centered=False):
    """-> (mean, inv_norm) float32[n] of emb[rows] (mean is 0 unless centered)"""
    n = int(rows.shape[0])
    mean = torch.empty(max(n, 1), dtype=torch.float32, device=emb.device)
    inv = torch.empty(max(n, 1), dtype=torch.float32, device=emb.device)
    check(lib().n2v_row_norms(ptr(emb), C.c_int32(emb.shape[1]), ptr(rows), C.c_int64(n), C.c_int(int(centered)),
                              ptr(mean), ptr(inv), stream()))
    return mean, inv


def sim_select(emb, rows_a, rows_b, *, thr=NEG_INF, thr_row=None, upper_only=False, skip_diagonal=False,
               exclude_keys=None, centered=False, capacity=1 << 22, norms=None):
    """One fused pass: every (r, c) with cos(emb[rows_a[r]], emb[rows_b[c]]) > threshold that survives
    the masks -> (a_pos int32, b_pos int32, score float32), unordered. Grows the buffer and repeats
    when more pairs pass than `capacity`."""
    dev = emb.device
    n_a, n_b = int(rows_a.shape[0]), int(rows_b.shape[0])
    if norms is None:
        ma, ia = row_norms(emb, rows_a, centered)
        mb, ib = (ma, ia) if rows_b is rows_a else row_norms(emb, rows_b, centered)
    else:
        ma, ia, mb, ib = norms
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    n_ex = 0 if exclude_keys is None else int(exclude_keys.shape[0])
    while True:
        oa = torch.empty(max(capacity, 1), dtype=torch.int32, device=dev)
        ob = torch.empty(max(capacity, 1), dtype=torch.int32, device=dev)
        os_ = torch.empty(max(capacity, 1), dtype=torch.float32, device=dev)
        count.zero_()
        check(lib().n2v_sim_threshold(ptr(emb), C.c_int32(emb.shape[1]), ptr(rows_a), C.c_int32(n_a), ptr(rows_b),
                                      C.c_int32(n_b), ptr(ma), ptr(ia), ptr(mb), ptr(ib), ptr(thr_row), C.c_float(thr),
                                      C.c_int(int(upper_only)), C.c_int(int(skip_diagonal)),
                                      ptr(exclude_keys) if n_ex else None, C.c_int64(n_ex), ptr(oa), ptr(ob), ptr(os_),
                                      C.c_int64(capacity), ptr(count), stream()))
        n = int(count.item())
        if n <= capacity:
            return oa[:n], ob[:n], os_[:n]
        capacity = int(n * 1.05) + 1024


def _sample_scores(emb, rows_a, rows_b, m, gen):
    dev = emb.device
    a = rows_a[torch.randint(0, rows_a.shape[0], (m,), device=dev, generator=gen)].contiguous()
    b = rows_b[torch.randint(0, rows_b.shape[0], (m,), device=dev, generator=gen)].contiguous()
    out = torch.empty(m, dtype=torch.float32, device=dev)
    check(lib().n2v_cosine_pairs(ptr(emb), C.c_int32(emb.shape[1]), ptr(a), ptr(b), C.c_int64(m), ptr(out), stream()))
    return out


def top_k_links(wv, words_a, words_b=None, k=10, exclude=()):
    """link_prediction's selection (main_link.py:70-171): cosine of every candidate pair -- words_a x
    words_b ("separated" user x item mode) or, with words_b None, every unordered pair i < j of words_a
    -- minus the `exclude` pairs (train edges, either orientation), global top-k by score.
    -> list of ((a, b), score), best first."""
    dev = require_cuda()
    words_a = list(words_a)
    same = words_b is None
    words_b = words_a if same else list(words_b)
    emb = _dev_emb(wv, dev)
    rows_a = torch.as_tensor([wv.vocab[w].index for w in words_a], dtype=torch.int32, device=dev)
    rows_b = rows_a if same else torch.as_tensor([wv.vocab[w].index for w in words_b], dtype=torch.int32, device=dev)
    n_a, n_b = len(words_a), len(words_b)
    pos_a = {w: i for i, w in enumerate(words_a)}
    pos_b = pos_a if same else {w: i for i, w in enumerate(words_b)}
    keys = set()
    for a, b in exclude:
        if a in pos_a and b in pos_b:
            keys.add(pos_a[a] * n_b + pos_b[b])
        if b in pos_a and a in pos_b:
            keys.add(pos_a[b] * n_b + pos_b[a])
    ex = torch.as_tensor(sorted(keys), dtype=torch.int64, device=dev) if keys else None
    total = n_a * (n_a - 1) // 2 if same else n_a * n_b
    if total <= 0 or k <= 0:
        return []
    norms = row_norms(emb, rows_a)
    norms = norms + (norms if same else row_norms(emb, rows_b))
    thr = NEG_INF
    if total > (1 << 21):           # threshold from a sample: expect ~max(8k, 64 sample points' worth) survivors
        m = 1 << 20
        smp = torch.sort(_sample_scores(emb, rows_a, rows_b, m, torch.Generator(device=dev).manual_seed(1)), descending=True).values
        j = int(max(64, min(m - 1, 8.0 * (k + len(keys)) / total * m)))
        thr = float(smp[j].item())
    while True:
        a, b, s = sim_select(emb, rows_a, rows_b, thr=thr, upper_only=same, exclude_keys=ex, norms=norms)
        if a.numel() >= min(k, total - len(keys)) or thr == NEG_INF:
            break
        thr = NEG_INF if thr <= -1.0 else thr - max(0.05, 0.5 * abs(thr))      # too few survivors: lower and repeat
    key = a.to(torch.int64) * n_b + b.to(torch.int64)
    order = torch.argsort(key)                                  # ties broken by pair position, then score descending
    s, key = s[order], key[order]
    order = torch.argsort(s, descending=True, stable=True)[:k]
    s, key = s[order].cpu().tolist(), key[order].cpu().tolist()
    return [((words_a[q // n_b], words_b[q % n_b]), sc) for q, sc in zip(key, s)]


def per_row_top_k(emb, rows, k, *, centered=False, sample_cols=1024):
    """the k most similar other rows of every row (diagonal scored 0 as main_link.py:386):
    -> (src, dst, score) sorted by (src, score descending), exactly k per row (k <= n)."""
    dev = emb.device
    n = int(rows.shape[0])
    k = min(int(k), n)
    if k <= 0:
        e = torch.zeros(0, dtype=torch.int64, device=dev)
        return e, e.clone(), torch.zeros(0, dtype=torch.float32, device=dev)
    norms = row_norms(emb, rows, centered)
    norms = norms + norms
    thr_row = torch.full((n,), NEG_INF, dtype=torch.float32, device=dev)
    if n > 4 * sample_cols and k < n // 8:
        # per-row threshold from `sample_cols` random columns: the j-th largest sampled score, with j chosen
        # so that about 2k + 32 columns are expected above it
        gen = torch.Generator(device=dev).manual_seed(2)
        cols = rows[torch.randint(0, n, (sample_cols,), device=dev, generator=gen)]
        j = int(max(1, min(sample_cols, (2.0 * k + 32.0) / n * sample_cols)))
        step = max(1, (1 << 24) // sample_cols)                 # n2v_cosine_pairs on <= 16 M sampled pairs at a time
        for r0 in range(0, n, step):                            # (plain cosine also when centred: it only has to be
            ra = rows[r0:r0 + step]                             #  close -- rows that come out short are redone below)
            a = ra.repeat_interleave(sample_cols).contiguous()
            b = cols.repeat(ra.shape[0]).contiguous()
            sc = torch.empty(a.shape[0], dtype=torch.float32, device=dev)
            check(lib().n2v_cosine_pairs(ptr(emb), C.c_int32(emb.shape[1]), ptr(a), ptr(b), C.c_int64(a.shape[0]), ptr(sc),
                                         stream()))
            thr_row[r0:r0 + step] = torch.topk(sc.view(-1, sample_cols), j, dim=1).values[:, -1]
        thr_row -= 1e-6
    done_src, done_dst, done_s = [], [], []
    todo = torch.arange(n, device=dev)
    while todo.numel():
        ra = rows[todo].contiguous()
        nm = (norms[0][todo].contiguous(), norms[1][todo].contiguous(), norms[2], norms[3])
        # skip_diagonal compares positions: the diagonal of a row subset is handled after the pass
        a, b, s = sim_select(emb, ra, rows, thr_row=thr_row[todo].contiguous(), norms=nm, centered=centered)
        src = todo[a.to(torch.int64)]
        dst = b.to(torch.int64)
        s = torch.where(src == dst, torch.zeros_like(s), s)
        keep = (src != dst) | (s > thr_row[src])
        src, dst, s = src[keep], dst[keep], s[keep]
        cnt = torch.bincount(src, minlength=n)
        short = (cnt[todo] < k)
        sel = ~short[torch.searchsorted(todo, src)]
        done_src.append(src[sel]); done_dst.append(dst[sel]); done_s.append(s[sel])
        todo = todo[short]
        thr_row[todo] = NEG_INF                                 # too few survivors in these rows: take everything
    src, dst, s = torch.cat(done_src), torch.cat(done_dst), torch.cat(done_s)
    # per row: score descending, ties by column (the reference's stable sort over user order, :391)
    o = torch.argsort(dst, stable=True)
    src, dst, s = src[o], dst[o], s[o]
    o = torch.argsort(s, descending=True, stable=True)
    src, dst, s = src[o], dst[o], s[o]
    o = torch.argsort(src, stable=True)
    src, dst, s = src[o], dst[o], s[o]
    start = torch.searchsorted(src, torch.arange(n, device=dev))
    rank = torch.arange(src.numel(), device=dev) - start[src]
    keep = rank < k
    return src[keep], dst[keep], s[keep]

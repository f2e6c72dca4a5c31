"""User-edge augmentation on the device: the user x user similarity and the per-user
edge selection of src/main_link.py:358-453 (get_similarity, build_user_sim_matrx, get_add_edge_by_ratio /
_by_step / _by_relu / _by_relu_ratio / _linear). The similarity tiles and the selection run fused in
n2v_sim_threshold (csrc/n2v_score.cu); the U x U matrix is never materialised (scoring.py).

    src, dst, w = user_edges(emb, user_nodes, mode="ratio", value=0.01)      # device tensors
    add_edge = as_tuples(user_nodes, src, dst, w)                            # the reference's list form
"""
from __future__ import annotations

import torch

from ._lib import require_cuda
from .scoring import NEG_INF, _dev_emb, per_row_top_k, sim_select


def user_edges(emb, user_nodes, mode="ratio", value=0.01, sim_method="cos", block_rows=None):
    """-> (src, dst, w): positions into user_nodes (int64, device) and float32 weights, in the
    reference's order (user by user; by score for the top-share modes, by user order otherwise).
    mode "ratio": the int(len(users) * value) most similar users of every user, weight 1 (:378-393);
    "relu-ratio": same selection, weight = similarity (:424-439); "step": similarity > value,
    weight 1 (:395-407); "relu": similarity > value, weight = similarity (:409-422);
    "linear": every pair, weight = similarity (:441-453). The self pair scores 0, as there (:386).
    sim_method "cos" | "pearson" (= cosine of the centred rows, :363)."""
    dev = require_cuda()
    if sim_method not in ("cos", "pearson"):
        raise NotImplementedError("sim_method %r (the reference's 'jsd' needs non-negative rows)" % sim_method)
    centered = sim_method == "pearson"
    E = _dev_emb(emb, dev)
    rows = torch.as_tensor([emb.vocab[str(w)].index for w in user_nodes], dtype=torch.int32, device=dev)
    n = len(user_nodes)
    if mode in ("ratio", "relu-ratio"):
        src, dst, s = per_row_top_k(E, rows, int(n * value), centered=centered)
        return src, dst, (s if mode == "relu-ratio" else torch.ones_like(s))
    if mode in ("step", "relu", "linear"):
        thr = NEG_INF if mode == "linear" else float(value)
        a, b, s = sim_select(E, rows, rows, thr=thr, skip_diagonal=True, centered=centered,
                             capacity=max(1 << 20, n * n if mode == "linear" else 0))
        o = torch.argsort(a.to(torch.int64) * n + b.to(torch.int64))
        a, b, s = a[o].to(torch.int64), b[o].to(torch.int64), s[o]
        return a, b, (torch.ones_like(s) if mode == "step" else s)
    raise ValueError("user-edges-mode value fault: " + str(mode))


def as_tuples(user_nodes, src, dst, w):
    """the reference's add_edge list: [(user, other_user, weight), ...]"""
    s, d, ww = src.cpu().numpy(), dst.cpu().numpy(), w.cpu().numpy()
    un = list(user_nodes)
    return [(un[a], un[b], float(c)) for a, b, c in zip(s.tolist(), d.tolist(), ww.tolist())]

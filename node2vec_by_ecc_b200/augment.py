"""User-edge augmentation on the device: the user x user similarity matrix and the
per-user edge selection of src/main_link.py:358-453 (get_similarity, build_user_sim_matrx,
get_add_edge_by_ratio / _by_step / _by_relu / _linear) as blocked cuBLAS SGEMMs + top-k / threshold
selections. The dense GEMM is a plain library GEMM (torch.matmul, fp32); everything else of the hot
path stays in libn2v_b200.so.

    src, dst, w = user_edges(emb, user_nodes, mode="ratio", value=0.01)      # device tensors
    add_edge = as_tuples(user_nodes, src, dst, w)                            # the reference's list form
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import require_cuda


def _rows(emb, words, dev):
    idx = torch.as_tensor([emb.vocab[str(w)].index for w in words], device=dev)
    m = emb._syn0_dev if getattr(emb, "_syn0_dev", None) is not None else torch.as_tensor(emb.syn0).to(dev)
    return m[idx].float()


def user_similarity_blocks(emb, user_nodes, sim_method="cos", block_rows=4096):
    """Yields (r0, S) with S = similarity of users [r0, r0+rows) against all users, diagonal 0
    (main_link.py:368-376 leaves it 0; the emb path sets user_user_sim_list[i] = 0, :386)."""
    dev = require_cuda()
    X = _rows(emb, user_nodes, dev)
    if sim_method == "pearson":                      # pearsonr == cosine of the centred rows
        X = X - X.mean(dim=1, keepdim=True)
    elif sim_method != "cos":
        raise NotImplementedError("sim_method %r (the reference's 'jsd' needs non-negative rows)" % sim_method)
    X = X / X.norm(dim=1, keepdim=True).clamp_min(1e-30)
    n = X.shape[0]
    for r0 in range(0, n, block_rows):
        S = X[r0:r0 + block_rows] @ X.T
        S[torch.arange(S.shape[0], device=dev), torch.arange(r0, r0 + S.shape[0], device=dev)] = 0.0
        yield r0, S


def user_edges(emb, user_nodes, mode="ratio", value=0.01, sim_method="cos", block_rows=4096):
    """-> (src, dst, w): positions into user_nodes (int64, device) and float32 weights.
    mode "ratio": the int(len(users) * value) most similar users of every user, weight 1 (:378-393);
    "relu-ratio": same selection, weight = similarity (:424-439); "step": similarity > value,
    weight 1 (:395-407); "relu": similarity > value, weight = similarity (:409-422);
    "linear": every pair, weight = similarity (:441-453)."""
    srcs, dsts, ws = [], [], []
    n = len(user_nodes)
    k = int(n * value) if mode in ("ratio", "relu-ratio") else 0
    for r0, S in user_similarity_blocks(emb, user_nodes, sim_method, block_rows):
        rows = S.shape[0]
        if mode in ("ratio", "relu-ratio"):
            if k == 0:
                continue
            val, col = torch.topk(S, k, dim=1)
            src = torch.arange(r0, r0 + rows, device=S.device)[:, None].expand(rows, k)
            srcs.append(src.reshape(-1)); dsts.append(col.reshape(-1))
            ws.append(val.reshape(-1) if mode == "relu-ratio" else torch.ones(rows * k, device=S.device))
        elif mode in ("step", "relu"):
            r, c = torch.nonzero(S > value, as_tuple=True)
            srcs.append(r + r0); dsts.append(c)
            ws.append(S[r, c] if mode == "relu" else torch.ones(r.numel(), device=S.device))
        elif mode == "linear":
            r, c = torch.meshgrid(torch.arange(rows, device=S.device), torch.arange(n, device=S.device), indexing="ij")
            srcs.append(r.reshape(-1) + r0); dsts.append(c.reshape(-1)); ws.append(S.reshape(-1))
        else:
            raise ValueError("user-edges-mode value fault: " + str(mode))
    if not srcs:
        e = torch.zeros(0, dtype=torch.int64, device=require_cuda())
        return e, e.clone(), torch.zeros(0, device=e.device)
    return torch.cat(srcs), torch.cat(dsts), torch.cat(ws).float()


def as_tuples(user_nodes, src, dst, w):
    """the reference's add_edge list: [(user, other_user, weight), ...]"""
    s, d, ww = src.cpu().numpy(), dst.cpu().numpy(), w.cpu().numpy()
    un = list(user_nodes)
    return [(un[a], un[b], float(c)) for a, b, c in zip(s.tolist(), d.tolist(), ww.tolist())]

"""Synthetic graphs of the shapes BASELINE.json names, generated with plain torch
ops from a counter-based hash (same edges on CPU and CUDA, any chunking): input generation only,
not part of the hot path.

  rmat_edges      C4/C5: R-MAT (a,b,c,d)=(0.57,0.19,0.19,0.05), symmetrised, de-duplicated,
                  self-loops removed, isolated vertices dropped, ids compacted
  planted_edges   C2: "BlogCatalog-shaped" heavy-tailed graph with planted communities
"""
from __future__ import annotations

import torch

_M64 = (1 << 64) - 1


def _wrap(c: int) -> int:
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def _mix64(z: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping multiply, logical shifts by masking)"""
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _wrap(0xBF58476D1CE4E5B9)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _wrap(0x94D049BB133111EB)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


def hash_uniform(seed: int, stream: int, idx: torch.Tensor) -> torch.Tensor:
    """float64 uniforms in [0,1) addressed by (seed, stream, idx)"""
    z = _mix64(idx * _wrap(0x9E3779B97F4A7C15) + _wrap(seed * 0xD1B54A32D192ED03 + stream * 0x2545F4914F6CDD1D + 1))
    return ((z >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / (1 << 53))


def rmat_edges(scale: int, n_edges: int, seed: int = 1, device="cuda", abcd=(0.57, 0.19, 0.19, 0.05),
               oversample: float = 1.12, chunk: int = 1 << 25):
    """-> (lo int32[M], hi int32[M], n_nodes): M <= n_edges distinct undirected edges lo < hi over
    compacted ids."""
    a, b, c, _ = abcd
    def raw(first, last):
        out = []
        for s in range(first, last, chunk):
            idx = torch.arange(s, min(s + chunk, last), dtype=torch.int64, device=device)
            src = torch.zeros_like(idx)
            dst = torch.zeros_like(idx)
            for lvl in range(scale):
                u = hash_uniform(seed, lvl, idx)
                sb = (u >= a + b).to(torch.int64)                       # quadrants c, d: source bit
                db = (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64)   # quadrants b, d
                src |= sb << lvl
                dst |= db << lvl
            lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
            keep = lo != hi
            out.append((lo[keep] << scale) | hi[keep])
            del idx, src, dst, lo, hi, keep
        return out

    done = int(n_edges * oversample)
    key = torch.unique(torch.cat(raw(0, done)))
    for _ in range(16):                 # R-MAT repeats edges: top up until n_edges distinct ones
        if key.numel() >= n_edges:
            break
        more = int((n_edges - key.numel()) * 1.5) + 1024
        key = torch.unique(torch.cat([key] + raw(done, done + more)))
        done += more
    if key.numel() > n_edges:      # keep a hash-random subset of exactly n_edges
        h = _mix64(key + _wrap(seed * 0x9E3779B97F4A7C15))
        key = key[torch.argsort(h)[:n_edges]]
        del h
    lo, hi = key >> scale, key & ((1 << scale) - 1)
    used = torch.unique(torch.cat([lo, hi]))
    lo = torch.searchsorted(used, lo).to(torch.int32)
    hi = torch.searchsorted(used, hi).to(torch.int32)
    return lo, hi, int(used.numel())


def planted_edges(n: int, m: int, seed: int = 42, gamma: float = 0.75, max_deg: int = 4000,
                  communities: int = 39, mu_in: float = 0.8, device="cuda"):
    """Heavy-tailed simple graph with a planted partition (node i in community i % communities):
    endpoint weights ~ (i+10)^-gamma capped at ~max_deg expected degree; an edge stays inside its
    first endpoint's community with probability mu_in. -> (lo, hi) int32, exactly m distinct edges
    (or fewer if the generator saturates)."""
    w = (torch.arange(n, dtype=torch.float64, device=device) + 10.0) ** (-gamma)
    w = torch.minimum(w, w.sum() * max_deg / (2.0 * m))
    cdf = torch.cumsum(w / w.sum(), 0)
    # per-community CDFs laid out community-major
    members = torch.argsort(torch.arange(n, device=device) % communities, stable=True)
    sizes = torch.bincount(torch.arange(n, device=device) % communities, minlength=communities)
    starts = torch.cumsum(sizes, 0) - sizes
    wc = w[members]
    cw = torch.cumsum(wc, 0)
    base = torch.cat([torch.zeros(1, dtype=torch.float64, device=device), cw])[starts]      # mass before community
    tot = torch.cat([cw, cw[-1:]])[starts + sizes - 1] - base
    keys = torch.zeros(0, dtype=torch.int64, device=device)
    rnd = 0
    while keys.numel() < m and rnd < 64:
        k = int((m - keys.numel()) * 1.3) + 1024
        idx = torch.arange(k, dtype=torch.int64, device=device) + rnd * (1 << 40)
        a = torch.searchsorted(cdf, hash_uniform(seed, 0, idx)).clamp_(max=n - 1)
        b = torch.searchsorted(cdf, hash_uniform(seed, 1, idx)).clamp_(max=n - 1)
        inside = hash_uniform(seed, 2, idx) < mu_in
        ca = a % communities
        target = base[ca] + hash_uniform(seed, 3, idx) * tot[ca]
        pos = torch.searchsorted(cw, target).clamp_(max=n - 1)
        pos = torch.minimum(torch.maximum(pos, starts[ca]), starts[ca] + sizes[ca] - 1)
        b = torch.where(inside, members[pos], b)
        lo, hi = torch.minimum(a, b), torch.maximum(a, b)
        ok = lo != hi
        new = torch.unique(lo[ok] * n + hi[ok])
        new = new[~torch.isin(new, keys)]
        if keys.numel() + new.numel() > m:
            h = _mix64(new + _wrap(seed + rnd))
            new = new[torch.argsort(h)[: m - keys.numel()]]
        keys = torch.cat([keys, new])
        rnd += 1
    keys = torch.sort(keys).values
    return (keys // n).to(torch.int32), (keys % n).to(torch.int32)


def csr_torch(lo: torch.Tensor, hi: torch.Tensor, n: int):
    """symmetric CSR from distinct undirected edges with torch ops only (used to hand the same
    graph to the CPU reference arm; the product path builds its CSR with n2v_csr_from_coo)."""
    src = torch.cat([lo, hi]).to(torch.int64)
    dst = torch.cat([hi, lo]).to(torch.int64)
    order = torch.argsort(src * n + dst)
    col = dst[order].to(torch.int32)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=lo.device)
    row_ptr[1:] = torch.cumsum(torch.bincount(src, minlength=n), 0)
    return row_ptr, col


def bipartite_edges(n_users: int = 200_000, n_items: int = 800_000, m: int = 20_000_000, seed: int = 7,
                    device="cuda"):
    """C3: user-item graph, user degree ~ lognormal (mean m/n_users), item choice
    ~ Zipf(1.0), integer weights 1..5. Users are ids [0, n_users), items [n_users, n_users+n_items)
    (the reference labels items int('9999999' + id)). -> (user int32[M], item int32[M],
    w float64[M], n_nodes) with M <= m distinct edges."""
    iu = torch.arange(n_users, dtype=torch.int64, device=device)
    # lognormal(sigma=1) user activity, normalised to m edges
    z = torch.sqrt(-2.0 * torch.log(hash_uniform(seed, 0, iu).clamp_min(1e-300))) * torch.cos(
        2.0 * torch.pi * hash_uniform(seed, 1, iu))
    act = torch.exp(z)
    ucdf = torch.cumsum(act / act.sum(), 0)
    icdf = torch.cumsum(1.0 / torch.arange(1, n_items + 1, dtype=torch.float64, device=device), 0)
    icdf = icdf / icdf[-1]
    keys = torch.zeros(0, dtype=torch.int64, device=device)
    rnd = 0
    while keys.numel() < m and rnd < 32:
        k = int((m - keys.numel()) * 1.15) + 1024
        idx = torch.arange(k, dtype=torch.int64, device=device) + rnd * (1 << 40)
        u = torch.searchsorted(ucdf, hash_uniform(seed, 2, idx)).clamp_(max=n_users - 1)
        it = torch.searchsorted(icdf, hash_uniform(seed, 3, idx)).clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, u * n_items + it]))
        rnd += 1
    if keys.numel() > m:
        keys = torch.sort(keys[torch.argsort(_mix64(keys + _wrap(seed)))[:m]]).values
    u, it = keys // n_items, keys % n_items
    w = (1 + torch.floor(hash_uniform(seed, 4, keys) * 5.0)).to(torch.float64)
    return u.to(torch.int32), (it + n_users).to(torch.int32), w, n_users + n_items

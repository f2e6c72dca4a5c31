"""ctypes binding of libn2v_b200.so (the C ABI in include/n2v_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
compute entry points raise. torch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libn2v_b200.so")
_lib = None

EXPORTS = [
    "n2v_last_error", "n2v_version", "n2v_sm_count", "n2v_csr_workspace_bytes", "n2v_csr_from_coo",
    "n2v_etab_workspace_bytes", "n2v_etab_offsets", "n2v_alias_build_nodes", "n2v_alias_build_edges",
    "n2v_walk_alias", "n2v_arc_record_bytes", "n2v_pack_arcs", "n2v_walk_alias_packed", "n2v_walk_reject", "n2v_pack_rows", "n2v_edge_hash_capacity", "n2v_edge_hash_build",
    "n2v_walk_reject_indexed", "n2v_walk_reject_law", "n2v_walk_reject_indexed_law", "n2v_vocab_count", "n2v_sgns_prepare_workspace_bytes",
    "n2v_sgns_prepare", "n2v_sgns_init", "n2v_sgns_train", "n2v_sgns_init_part", "n2v_sgns_train_sharded",
    "n2v_sgns_groups_workspace_bytes", "n2v_sgns_groups_count", "n2v_sgns_groups_fill", "n2v_sgns_train_groups", "n2v_cosine_pairs", "n2v_row_norms", "n2v_sim_threshold", "n2v_format_workspace_bytes", "n2v_format_walks_offsets", "n2v_format_walks_write", "n2v_parse_workspace_bytes", "n2v_parse_walks_index", "n2v_parse_walks_fill",
    "n2v_random_gather_bench",
]


class N2VError(RuntimeError):
    pass


class SgnsParams(C.Structure):
    """n2v_sgns_params_t"""
    _fields_ = [
        ("V", C.c_int32), ("dim", C.c_int32), ("window", C.c_int32), ("negative", C.c_int32),
        ("bucket_bits", C.c_int32), ("max_sentence_len", C.c_int32),
        ("alpha0", C.c_float), ("min_alpha", C.c_float),
        ("total_examples", C.c_int64), ("example_base", C.c_int64), ("sent_per_job", C.c_int64),
        ("epoch", C.c_uint32), ("seed", C.c_uint64),
        ("grid_warps", C.c_int32), ("atomic_updates", C.c_int32), ("negative_sharing", C.c_int32), ("tuning", C.c_int32),
        ("hot_rows", C.c_int32),
    ]


class WalkLaw(C.Structure):
    """n2v_walk_law_t"""
    _fields_ = [("first_slots", C.c_void_p), ("pop_edges", C.c_int32)]


def lib():
    """Load libn2v_b200.so; raises (never falls back) when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise N2VError(
            f"{SO_PATH} is missing: build it with `python -m node2vec_by_ecc_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    L.n2v_last_error.restype = C.c_char_p
    for name in ("n2v_csr_workspace_bytes", "n2v_etab_workspace_bytes", "n2v_sgns_prepare_workspace_bytes",
                 "n2v_arc_record_bytes", "n2v_format_workspace_bytes", "n2v_parse_workspace_bytes",
                 "n2v_sgns_groups_workspace_bytes"):
        getattr(L, name).restype = C.c_size_t
    L.n2v_edge_hash_capacity.restype = C.c_uint64
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise N2VError(f"libn2v_b200 error {rc}: {lib().n2v_last_error().decode()}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise N2VError("no CUDA device: node2vec_by_ecc_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


_dummy = {}


def ptr(t):
    """device pointer of a tensor (None -> NULL). Empty tensors get a valid dummy address: the C
    ABI treats NULL as "absent", zero sizes are passed explicitly."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "need a contiguous CUDA tensor"
    if t.numel() == 0:
        d = _dummy.get(t.device)
        if d is None:
            d = _dummy[t.device] = torch.zeros(64, dtype=torch.uint8, device=t.device)
        return C.c_void_p(d.data_ptr())
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)

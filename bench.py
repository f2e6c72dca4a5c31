#!/usr/bin/env python
"""bench.py -- the hot path (second-order walks -> skip-gram negative sampling) on synthetic
graphs of the shape BASELINE.json names. One JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on host cores

Workload (config.workload): C4 = R-MAT scale 22, (a,b,c,d)=(.57,.19,.19,.05), 100 M undirected
edges after symmetrise/dedup, rejection-sampling walker, p=0.25 q=4, R=10 walks x L=80 per node,
SGNS d=128 window=10 negative=5 sample=1e-3 (the main.py defaults). A STEP is one pass of the hot
path over one batch: `batch_walks` walks per GPU are simulated (n2v_walk_reject) and the batch is
trained on (n2v_sgns_train); with N > 1 the walk ids of a step are split across ranks (weak
scaling: batch per GPU fixed) and SGNS runs block-partitioned (--multi-gpu-sgns block, the
default): the tables are cut into N row sets, the step's walks of all GPUs form one pool, GPU k
trains the group stream (centre in part k, context in part (k + e) % N) in sub-step e and the syn0
parts travel round an NCCL ring -- no replicas (node2vec_by_ecc_b200.word2vec.BlockSgnsTrainer).
`replica` (full copies, delta-sum all-reduce every `sync_walks` walks) and `peer` (one table pair
in NVLink peer memory) are the measured alternatives.
value = (centre, context) pairs trained per second through the whole step, all ranks.
e2e = the same step through the reference-facing classes with host buffers: start nodes in pinned host memory
-> node2vec.Graph.simulate_walks -> walks copied back to pinned host memory -> gensim-style Word2Vec.train on
the returned corpus -> pair count read back (at N > 1: Graph(distributed=True), block-partitioned Word2Vec).
roofline: frac_model / frac_8d / frac_dram = the kernel time against its own algorithmic row bytes, SURVEY 8d's
7,168 B/pair, and the DRAM bytes ncu measured (profiles/ncu_traffic.json, refused when the kernel sources
changed since the capture). Block mode keeps the single-GPU negative law (one set per centre occurrence).
Inputs are larger than L2 (tables 2 x 1.35 GB, CSR 0.8 GB, arc hash 4.3 GB), no L2 flush needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_PAIR = 7168        # 2 x 512 B x (1 input + 1 positive + 5 negative rows)
BYTES_PER_ALIAS_STEP = 40
KERNEL_SOURCES = {     # the files a kernel is compiled from: a capture is valid only for these exact sources
    "sgns_train_kernel_v3": ["n2v_sgns.cu", "n2v_sgns_stage.cuh", "n2v_common.cuh"],
    "sgns_train_kernel_v2": ["n2v_sgns.cu", "n2v_sgns_stage.cuh", "n2v_common.cuh"],
    "sgns_group_kernel": ["n2v_sgns_block.cu", "n2v_sgns_stage.cuh", "n2v_common.cuh"],
    "walk_reject_indexed_kernel": ["n2v_walk2.cu", "n2v_reject.cuh", "n2v_common.cuh"],
    "walk_reject_kernel": ["n2v_walk.cu", "n2v_reject.cuh", "n2v_common.cuh"],
    "walk_alias_kernel": ["n2v_walk.cu", "n2v_common.cuh"],
}


def source_sha16(kernel):
    import hashlib
    h = hashlib.sha256()
    for f in KERNEL_SOURCES.get(kernel, []):
        with open(os.path.join(ROOT, "node2vec_by_ecc_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel, workload_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    of this kernel on this workload (profiles/ncu_traffic.json, written by scripts/capture_traffic.py).
    A capture taken from other kernel sources than the ones on disk is refused: -> (None, why)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(kernel)
    except Exception:
        ent = None
    if not ent:
        return None, "no ncu --set full capture of this kernel is committed"
    if ent.get("source_sha16") != source_sha16(kernel):
        return None, "stale: %s was captured from other kernel sources (%s)" % (ent.get("profile"), ent.get("source_sha16"))
    if ent.get("workload") != workload_key:
        return None, "capture %s is of another workload (%s)" % (ent.get("profile"), ent.get("workload"))
    return float(ent["dram_bytes_per_launch"]), ent.get("profile")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--edges", type=float, default=100e6)
    ap.add_argument("--batch-walks", type=int, default=1 << 19)
    ap.add_argument("--walk-length", type=int, default=80)
    ap.add_argument("--num-walks", type=int, default=10)
    ap.add_argument("--p", type=float, default=0.25)
    ap.add_argument("--q", type=float, default=4.0)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--window", type=int, default=10)
    ap.add_argument("--negative", type=int, default=5)
    ap.add_argument("--walk-mode", default="reject", choices=["reject", "alias"])
    ap.add_argument("--walk-indexed", type=int, default=1, help="1 = hashed distance-1 test + state machine")
    ap.add_argument("--hogwild-warps", type=int, default=0)
    ap.add_argument("--atomic", type=int, default=1)
    ap.add_argument("--shared-negatives", type=int, default=1)
    ap.add_argument("--sync-walks", type=int, default=0,
                    help="walks per rank between two delta-sum syncs of the replicated tables (N > 1); "
                         "0 = auto (total pairs per sync <= 100 V / N), -1 = once per step")
    ap.add_argument("--multi-gpu-sgns", default="block", choices=["block", "peer", "replica"],
                    help="N > 1: block = tables cut into N row sets, orthogonal (centre part, context part) pair "
                         "buckets per sub-step, syn0 parts passed round a ring (no replicas); peer = one table pair "
                         "sharded over the GPUs' HBM, trained over NVLink peer memory; replica = a full copy per GPU, "
                         "delta-sum all-reduce every --sync-walks")
    ap.add_argument("--neg-group", type=int, default=1,
                    help="block mode: token positions of a walk whose centres share one negative set (1 = one set per "
                         "centre occurrence, the single-GPU law)")
    ap.add_argument("--sgns-variant", default="default", choices=["default", "mma"],
                    help="mma = the tensor-core window-batch experiment (csrc/n2v_sgns_mma.cu; one GPU, shared negatives): "
                         "timed for DESIGN.md 3.4's choice, not a parity mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-walks", type=int, default=0, help="walks per reference-arm step (0 = auto)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def config_of(a, n_nodes=None, nnz=None):
    c = {"workload": f"C4 R-MAT scale {a.scale}, {int(a.edges)} undirected edges, {a.walk_mode} walker "
                     f"p={a.p} q={a.q}, R={a.num_walks} L={a.walk_length}; SGNS d={a.dim} window={a.window} "
                     f"negative={a.negative} sample=1e-3",
         "batch_walks_per_gpu": a.batch_walks, "generator": "node2vec_by_ecc_b200.synth.rmat_edges seed=1",
         "sgns_negatives": ("one set of 5 per %d consecutive centre positions of a walk, shared by their context pairs" % a.neg_group
                            if (a.shared_negatives and a.gpus > 1 and a.multi_gpu_sgns == "block" and a.neg_group > 1) else
                            "one set of 5 per centre, shared by its context pairs (other_negative_mode = fresh set per pair)"
                            if a.shared_negatives else "fresh set of 5 per (centre, context) pair, gensim's law"),
         "l2": "inputs larger than L2 (no flush)"}
    if n_nodes is not None:
        c["n_nodes"], c["nnz"] = n_nodes, nnz
    return c


# ==================================================================================================
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import ctypes as C
    from node2vec_by_ecc_b200 import DeviceGraph, SgnsTrainer, synth
    from node2vec_by_ecc_b200._lib import check, lib, ptr, stream

    t_setup = time.time()
    lo, hi, n = synth.rmat_edges(a.scale, int(a.edges), seed=1, device=dev)
    dg = DeviceGraph.from_coo(lo, hi, None, n, undirected=True)
    del lo, hi
    torch.cuda.empty_cache()
    B, L, R = a.batch_walks, a.walk_length, a.num_walks
    total_walks = R * n
    tables = None
    if a.walk_mode == "alias":
        tables = dg.build_alias_tables(a.p, a.q)
    walks = torch.empty((B, L), dtype=torch.int32, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    ar = torch.arange(B, dtype=torch.int64, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)

    def do_walk(starts, base):
        if tables is not None:
            dg.walk_alias(tables, starts, L, 1, base, out=(walks, lens))
        else:
            dg.walk_reject(a.p, a.q, starts, L, 1, base, counters=counters, out=(walks, lens),
                           indexed=bool(a.walk_indexed))

    # vocabulary (scan_vocab): one walk per node, this rank's contiguous shard, counts summed
    counts = torch.zeros(n, dtype=torch.int64, device=dev)
    per = (n + world - 1) // world
    for s in range(rank * per, min(n, (rank + 1) * per), B):
        e = min(s + B, n, (rank + 1) * per)
        st = torch.arange(s, e, dtype=torch.int32, device=dev)
        if e - s < B:
            st = torch.cat([st, st.new_zeros(B - (e - s))])
        do_walk(st, (1 << 40) + s)
        check(lib().n2v_vocab_count(ptr(walks), C.c_int64((e - s) * L), C.c_int32(n), ptr(counts), stream()))
    if world > 1:
        dist.all_reduce(counts)
    peer = world > 1 and a.multi_gpu_sgns == "peer" and bool(a.shared_negatives)
    block = world > 1 and a.multi_gpu_sgns == "block" and bool(a.shared_negatives)
    if block:     # tables cut into `world` row sets; orthogonal pair buckets; syn0 parts round a ring
        from node2vec_by_ecc_b200 import BlockSgnsTrainer
        trainer = BlockSgnsTrainer(counts, dim=a.dim, window=a.window, negative=a.negative, sample=1e-3, seed=1,
                                   neg_group=a.neg_group)
    elif peer:    # ONE table pair spread over the GPUs' HBM, trained by all of them over NVLink
        from node2vec_by_ecc_b200 import PeerSgnsTrainer
        trainer = PeerSgnsTrainer(counts, dim=a.dim, window=a.window, negative=a.negative, sample=1e-3, seed=1)
    else:
        trainer = SgnsTrainer(counts, dim=a.dim, window=a.window, negative=a.negative, sample=1e-3, seed=1)
    grid_warps = a.hogwild_warps or trainer.default_hogwild_warps(bool(a.shared_negatives))
    counters.zero_()
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern = {"walk": [], "sgns": []}

    mode = {"shared": int(a.shared_negatives)}
    from node2vec_by_ecc_b200.dist import ReplicaSync, sync_walks_per_rank
    replica_sync = None if (peer or block) else ReplicaSync(trainer.syn0, trainer.syn1neg)
    pairs_per_walk = (2 * a.window + 1) * L / 2.0            # ~ mean reduced window = (window + 1) / 2 each side
    if world == 1 or peer or block or a.sync_walks < 0:
        sync_walks = B
    elif a.sync_walks > 0:
        sync_walks = min(B, a.sync_walks)
    else:
        sync_walks = min(B, sync_walks_per_rank(trainer.V, world, pairs_per_walk))

    def step(i, record=False):
        g0 = (i * world + rank) * B                         # global id of this rank's first walk
        starts = ((g0 + ar) % n).to(torch.int32)
        e0, e1 = ev(), ev()
        e0.record(); do_walk(starts, g0); e1.record()
        # SGNS in sub-batches of `sync_walks`, tables combined (delta-sum) after each (N > 1, replica mode)
        for sa in range(0, B, sync_walks):
            sb_ = min(B, sa + sync_walks)
            e2, e3 = ev(), ev()
            e2.record()
            if block:     # one pool = this step's walks of all ranks (rank order = global walk id order)
                trainer.train(walks, None, B, L, total_examples=total_walks, example_base=(i * world * B) % total_walks,
                              sent_id_base=i * world * B, sent_per_job=10000 // L, grid_warps=a.hogwild_warps or None,
                              exact_bounds=False)
            else:
                trainer.train(walks[sa:sb_], None, sb_ - sa, L, total_examples=total_walks,
                              example_base=(g0 + sa) % total_walks, sent_id_base=g0 + sa, sent_per_job=10000 // L,
                              grid_warps=a.hogwild_warps or trainer.default_hogwild_warps(bool(mode["shared"])),
                              atomic_updates=a.atomic, negative_sharing=mode["shared"])
            e3.record()
            if replica_sync is not None:
                replica_sync.sync()
            if record:
                kern["sgns"].append((e2, e3))
        if record:
            kern["walk"].append((e0, e1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(first_step, record=False):
        p0 = trainer.pairs.clone()
        c0 = counters.clone()
        barrier()
        s, e = ev(), ev()
        s.record()
        for i in range(first_step, first_step + a.steps):
            step(i, record)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        pr = trainer.pairs - p0
        cn = counters - c0
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(pr); dist.all_reduce(cn)
        return float(ms.item()), int(pr[0].item()), cn.cpu().numpy(), int(pr[1].item())

    for i in range(a.warmup):
        step(i)
    sampler = ClockSampler(local)          # every rank watches its own GPU; rank 0 reports all of them
    if block:
        trainer.phase_events = []
    ms, pairs, cn, centres = timed(a.warmup, record=True)
    clocks = sampler.stop()
    if block:
        trainer.check_overflow()          # device-side stream bounds: a pool that did not fit its buffer would show here
    phases = None
    if block:     # rank 0's split of the SGNS phase: pool all-gather, pair expansion, bucket kernels, ring passes
        phases = {}
        for name, x, y in trainer.phase_events:
            phases[name] = phases.get(name, 0.0) + x.elapsed_time(y)
        trainer.phase_events = None
    walk_ms = sum(x.elapsed_time(y) for x, y in kern["walk"])
    sgns_ms = sum(x.elapsed_time(y) for x, y in kern["sgns"])
    by_rank = None
    if world > 1:      # outside the timed region: every rank's clocks and SGNS phase split, for imbalance / throttling
        by_rank = [None] * world
        dist.all_gather_object(by_rank, {"rank": rank, "clocks": clocks, "walk_ms_per_step": walk_ms / a.steps,
                                         "sgns_ms_per_step": sgns_ms / a.steps,
                                         "sgns_phases_ms_per_step": ({k: v / a.steps for k, v in phases.items()} if phases else None)})
        bad = sorted({r for x in by_rank for r in x["clocks"]["reasons"]})
        clocks = dict(clocks, reasons=bad, sm_mhz_min_over_ranks=min((x["clocks"]["sm_mhz"] or 0) for x in by_rank))
    my_pairs_rank = pairs / world
    steps_walked = int(cn[0]) if tables is None else None

    e2e = None
    if not a.no_e2e:
        # the same step through the reference-facing classes (the calls of src/main.py:92-101): start nodes
        # come from pinned host memory, node2vec.Graph.simulate_walks walks them, the walks go back to the
        # host (what main_link.py:544-546 writes to the walk file) and gensim-style Word2Vec.train learns
        # from the corpus object simulate_walks returned; the pair count is read back every step.
        from node2vec_by_ecc_b200 import Graph, Word2Vec
        G = Graph(dg, False, a.p, a.q, seed=1, mode=a.walk_mode, distributed=world > 1)
        G.preprocess_transition_probs()
        model = Word2Vec(size=a.dim, window=a.window, min_count=0, sg=1, workers=8, iter=1, negative=a.negative,
                         shared_negatives=a.shared_negatives, atomic_updates=a.atomic)
        model.build_vocab(G.simulate_walks(1, L))                  # scan_vocab over one walk per node
        T = model.trainer
        hs = torch.empty(B * world, dtype=torch.int32).pin_memory()
        hw = torch.empty((B, L), dtype=torch.int32).pin_memory()
        hl = torch.empty(B, dtype=torch.int32).pin_memory()
        hp = torch.empty(1, dtype=torch.int64).pin_memory()
        alpha_at = lambda i: max(1e-4, 0.025 - (0.025 - 1e-4) * ((i * world * B) % total_walks) / total_walks)

        def api_step(i):
            g0 = i * world * B
            np.copyto(hs.numpy(), ((g0 + np.arange(B * world, dtype=np.int64)) % n).astype(np.int32))
            corpus = G.simulate_walks(1, L, nodes=hs)              # H2D of the start nodes inside
            hw.copy_(corpus.walks, non_blocking=True); hl.copy_(corpus.lens, non_blocking=True)
            model.train(corpus, total_examples=len(corpus), epochs=1, start_alpha=alpha_at(i), end_alpha=alpha_at(i + 1))
            hp.copy_(T.pairs[:1], non_blocking=True)
            torch.cuda.current_stream().synchronize()              # the caller reads the result

        first = a.warmup + a.steps
        for i in range(first, first + 2):                          # warm the pinned path and the allocator
            api_step(i)
        p0 = T.pairs.clone()
        barrier()
        s_, e_ = ev(), ev()
        s_.record()
        for i in range(first + 2, first + 2 + a.steps):
            api_step(i)
        e_.record()
        barrier()
        ms_e = torch.tensor([s_.elapsed_time(e_)], dtype=torch.float64, device=dev)
        pr_e = (T.pairs - p0)[:1].clone()
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX); dist.all_reduce(pr_e)
        ms_e, pairs_e = float(ms_e.item()), int(pr_e.item())
        e2e = {"value": pairs_e / (ms_e / 1e3), "unit": "pairs/s",
               "h2d_bytes_per_step": B * world * 4 * world, "d2h_bytes_per_step": (B * L * 4 + B * 4 + 8) * world,
               "ms_per_step": ms_e / a.steps,
               "api": "node2vec.Graph(...).simulate_walks(1, L, nodes=<pinned host array>) -> walks copied to pinned host "
                      "memory -> Word2Vec.train(corpus, total_examples=, epochs=1, start_alpha=, end_alpha=) -> pair count "
                      "read back; trainer " + type(T).__name__}
        del model, G, T
        torch.cuda.empty_cache()

    other = None
    if not a.no_e2e and not peer and not block:      # the other negative-sampling mode, same steps, kernel-timed
        kern_main = kern
        kern = {"walk": [], "sgns": []}
        mode["shared"] = 1 - mode["shared"]
        step(a.warmup + 2 * a.steps + 2)
        ms_o, pairs_o, _, _ = timed(a.warmup + 2 * a.steps + 3, record=True)
        sg_o = sum(x.elapsed_time(y) for x, y in kern["sgns"])
        other = {"shared_negatives": mode["shared"], "value": pairs_o / (ms_o / 1e3), "unit": "pairs/s",
                 "sgns_pairs_per_s_kernel": pairs_o / (sg_o / 1e3), "ms_per_step": ms_o / a.steps,
                 "algorithmic_GBps_at_7168B_per_pair": pairs_o / world * BYTES_PER_PAIR / (sg_o / 1e3) / 1e9}
        mode["shared"] = 1 - mode["shared"]
        kern = kern_main

    out = None
    if rank == 0:
        peak, src = peaks()
        k_ms = sgns_ms
        wkey = "scale %d edges %d batch %d L %d world %d" % (a.scale, int(a.edges), B, L, world)
        if block:
            # group kernel: per pair the input row, per carried output row (one centre row per group + 5 per
            # negative set) one read + one reduction: 1,024 B each; `centres` = carried rows
            alg_bytes = (my_pairs_rank + centres / world) * 1024.0
            kshort = "sgns_group_kernel"
            kname = "sgns_group_kernel (block-partitioned tables, one negative set per centre occurrence)"
            k_ms = phases["train"]
        elif a.shared_negatives:
            # shared-negative kernel: per pair the input row (read + written, 1,024 B), per centre
            # the 6 carried output rows (read + written once, 6,144 B)
            alg_bytes = (my_pairs_rank * 1024.0 + centres / world * 6144.0)
            kshort = "sgns_train_kernel_v3"
            kname = "sgns_train_kernel_v3 (one negative set per centre occurrence)"
            if a.sgns_variant == "mma":
                kshort = kname = "sgns_train_kernel_mma (EXPERIMENT: window-batch semantics, TF32 tensor cores)"
        else:
            alg_bytes = my_pairs_rank * float(BYTES_PER_PAIR)
            kshort = "sgns_train_kernel_v2"
            kname = "sgns_train_kernel_v2 (per-pair negatives, gensim's law)"
        n_launch = max(1, len(kern["sgns"])) * (world if block else 1)
        sg_gbs = alg_bytes / (k_ms / 1e3) / 1e9
        traffic, traffic_src = ncu_traffic(kshort, wkey)
        gb_8d = my_pairs_rank * BYTES_PER_PAIR / (k_ms / 1e3) / 1e9
        shared_law = bool(a.shared_negatives)
        roof = {"kernel": kname, "bound": "hbm", "achieved": sg_gbs, "peak": peak, "unit": "GB/s",
                "frac": sg_gbs / peak,
                # the same kernel time against three byte counts:
                "frac_model": sg_gbs / peak,          # this kernel's own algorithmic bytes (rows it must move)
                "frac_8d": None if shared_law else gb_8d / peak,      # 7,168 B/pair, fresh negatives per pair
                "frac_8d_note": ("n/a: 7,168 B/pair assumes 5 fresh negative rows per pair; this kernel shares one set per "
                                 "centre occurrence, so it does not move those bytes (it would read %.2f)" % (gb_8d / peak)
                                 if shared_law else "SURVEY.md 8d byte model applies (gensim's per-pair law)"),
                "frac_dram": (traffic / (k_ms / n_launch / 1e3) / 1e9 / peak) if traffic else None,   # measured DRAM
                                                                              # bytes (ncu, same launch size) over the live kernel time
                "traffic": traffic, "traffic_source": traffic_src,
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)", "peak_source": src,
                "bound_note": "not HBM-stream-bound: ~75 % of the rows are served by the 126 MB L2, so frac_model (row bytes the "
                              "algorithm moves) exceeds frac_dram (bytes that reach HBM). What binds every SGNS kernel here is "
                              "the rate at which L2 + HBM serve random 512-byte rows and row reductions: ~11 G row operations/s "
                              "against 5.9 G reads/s or 4.7 G reductions/s from DRAM alone (DESIGN.md 3.4, "
                              "profiles/r02_a_row_microbench.jsonl); halving the FMA issue slots (FFMA2) moved it by +2 %",
                "algorithmic_bytes_per_launch": alg_bytes / n_launch,
                "algorithmic_bytes_per_pair": alg_bytes / my_pairs_rank, "pairs_per_launch": my_pairs_rank / n_launch,
                ("carried_rows_per_launch" if block else "centres_per_launch"): centres / world / n_launch,
                "ms_per_launch": k_ms / n_launch,
                "GBps_at_7168B_per_pair": gb_8d}
        if tables is None:
            S, T, P = (float(cn[0]) / world, float(cn[1]) / world, float(cn[3]) / world)
            wbytes = 20 * S + 4 * T + 4 * P
        else:
            S = float(B) * (L - 1) * a.steps
            T = P = 0.0
            wbytes = BYTES_PER_ALIAS_STEP * S
        w_gbs = wbytes / (walk_ms / 1e3) / 1e9
        w_traffic, w_traffic_src = ncu_traffic("walk_reject_indexed_kernel" if (a.walk_mode == "reject" and a.walk_indexed)
                                               else "walk_%s_kernel" % a.walk_mode, wkey)
        roof_walk = {"kernel": ("walk_reject_indexed_kernel" if (a.walk_mode == "reject" and a.walk_indexed) else "walk_%s_kernel" % a.walk_mode), "bound": "hbm", "achieved": w_gbs, "peak": peak,
                     "unit": "GB/s", "frac": w_gbs / peak, "traffic": w_traffic, "traffic_source": w_traffic_src,
                     "frac_dram": (w_traffic / (walk_ms / a.steps / 1e3) / 1e9 / peak) if w_traffic else None,
                     # against the random-access ceiling instead of the stream peak: a random 8-byte read costs HBM a
                     # 128-byte access (4 sectors); the gather microbenchmark sustains 1.36e11-1.63e11 DRAM sectors/s
                     # (profiles/r02_m_microbench_sectors.csv) -- 1.55e11 taken as the ceiling
                     "frac_random_access": (w_traffic / 32.0 / (walk_ms / a.steps / 1e3) / 1.55e11) if w_traffic else None,
                     "random_access_ceiling_dram_sectors_per_s": 1.55e11,
                     "algorithmic_bytes_per_launch": wbytes / a.steps,
                     "steps_per_s_kernel": S / (walk_ms / 1e3), "trials_per_step": (T / S) if S else None,
                     "probes_per_step": (P / S) if S else None, "ms_per_launch": walk_ms / a.steps}
        out = {
            "metric": "walk steps/s & SGNS pairs/s (value = SGNS pairs/s through the walk->SGNS step)",
            "value": pairs / (ms / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 rows / int32 ids", "data": "synthetic",
            "config": config_of(a, n, dg.nnz),
            "walk_steps_per_s": (float(cn[0]) if tables is None else S * world) / (walk_ms / 1e3),
            "sgns_pairs_per_s_kernel": pairs / (sgns_ms / 1e3),
            "roofline": roof, "roofline_walk": roof_walk, "e2e": e2e, "other_negative_mode": other,
            "gpu_launches": a.steps * ((1 + 2 + world) if block else (1 + (B + sync_walks - 1) // sync_walks)),
            "sgns_phases_ms_per_step": ({k: v / a.steps for k, v in phases.items()} if phases else None),
            "multi_gpu_sgns": None if world == 1 else (
                {"tables": "cut into N row sets (row i -> GPU i % N); a step's walks of all GPUs form one pool (all-gather, "
                           "4 B/token); GPU k expands the pairs whose centre is in part k, bucketed by the context's part, and "
                           "trains bucket (k, (k + e) % N) in sub-step e against syn1neg part k and the syn0 part it holds, "
                           "which then moves to GPU k - 1 (NCCL send/recv ring); no replicas, no averaging",
                 "neg_group": a.neg_group, "pool_walks": B * world} if block else
                {"tables": "one syn0/syn1neg pair, row i in GPU i % N's HBM, every GPU trains its own walks against all parts "
                           "over NVLink peer memory (red.global.add.v4.f32); no replicas, no sync"} if peer else
                {"tables": "replicated; delta-sum all-reduce of syn0 and syn1neg", "walks_per_gpu_per_sync": sync_walks,
                 "syncs_per_step": (B + sync_walks - 1) // sync_walks}),
            "clocks": clocks, "by_rank": by_rank, "hogwild_warps": grid_warps, "atomic_updates": a.atomic, "shared_negatives": a.shared_negatives, "setup_s": t_setup,
        }
        if not a.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(a, dg, trainer, walks)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ==================================================================================================
def _oracle_walk_rate(g, a, budget_s, threads, seed=5):
    """the reference's default walker for graphs whose tables do not fit: on-the-fly tables
    (node2vec.py:34-53; settings.py ON_THE_FLY_BOOL), C port, `threads` host threads."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    rng = np.random.RandomState(seed)
    deg = np.diff(g.row_ptr)
    done_steps, n_walks, t0 = 0, 0, time.time()
    all_walks = []
    with ThreadPoolExecutor(threads) as ex:
        while time.time() - t0 < budget_s:
            starts = rng.randint(0, g.n, size=threads).astype(np.int32)
            futs = [ex.submit(oracle.walks_on_the_fly, g, a.p, a.q, starts[i:i + 1], a.walk_length, 1, n_walks + i)
                    for i in range(threads)]
            for f in futs:
                w, l = f.result()
                done_steps += int(l[0]) - 1
                all_walks.append(w)
            n_walks += threads
    dt = time.time() - t0
    return done_steps / dt, done_steps, n_walks, dt, np.concatenate(all_walks), int(deg.max())


def cpu_baseline(a, dg, trainer, walks_dev):
    """The oracle port on this box's host cores, on a bounded sample of the same workload."""
    import numpy as np
    import oracle
    cores = os.cpu_count() or 1
    g = oracle.CSR(dg.row_ptr.cpu().numpy(), dg.col.cpu().numpy(), None)
    wrate, wsteps, nw, wdt, _, _ = _oracle_walk_rate(g, a, 8.0, cores)
    # SGNS: the device-generated walks of the last step, full-size tables, all cores
    counts_by_id = np.zeros(dg.n, dtype=np.int64)
    counts_by_id[trainer.order.cpu().numpy()] = trainer.counts.cpu().numpy()
    voc = oracle.vocab_from_counts(counts_by_id)
    rng = np.random.default_rng(1)
    syn0 = ((rng.random((voc.V, a.dim), dtype=np.float32) - 0.5) / a.dim).astype(np.float32)
    syn1 = np.zeros((voc.V, a.dim), dtype=np.float32)
    S = 2048
    rate, pairs, dt = 0.0, 0, 0.0
    for _ in range(3):
        w = walks_dev[:S].cpu().numpy()
        tok = np.where(w >= 0, voc.id2index[np.maximum(w, 0)], -1).astype(np.int32)
        off = np.arange(S + 1, dtype=np.int64) * a.walk_length
        t0 = time.time()
        _, _, pairs = oracle.sgns_train(tok, off, voc, dim=a.dim, window=a.window, negative=a.negative,
                                        workers=cores, rng_mode=0, seed=1, syn0=syn0, syn1neg=syn1)
        dt = time.time() - t0
        rate = pairs / dt
        if dt > 6.0 or S >= walks_dev.shape[0]:
            break
        S = int(min(walks_dev.shape[0], max(S * 2, S * 10.0 / max(dt, 1e-3))))
    return {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"SGNS: {S} device-generated walks ({pairs} pairs, {dt:.1f} s) against full [V={voc.V},{a.dim}] "
                      f"tables, oracle/sgns_oracle.c with {cores} threads; walker: {nw} on-the-fly walks "
                      f"({wsteps} steps, {wdt:.1f} s), oracle/n2v_oracle.c on {cores} threads",
            "walk_steps_per_s": wrate}


def run_reference(a):
    """The reference's own CPU implementation of the path (the oracle port: the reference is
    Python + gensim and cannot run here) with all host threads, whole path per step:
    on-the-fly walks (its default for graphs whose tables do not fit) -> SGNS on those walks."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    import oracle
    from node2vec_by_ecc_b200 import synth
    cores = os.cpu_count() or 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"      # input generation only (plain torch ops)
    lo, hi, n = synth.rmat_edges(a.scale, int(a.edges), seed=1, device=dev)
    rp, col = synth.csr_torch(lo, hi, n)
    g = oracle.CSR(rp.cpu().numpy(), col.cpu().numpy(), None)
    del lo, hi, rp, col
    deg = np.diff(g.row_ptr)
    voc = oracle.vocab_from_counts(deg)                       # visit counts ~ degree (stationary law)
    rng = np.random.default_rng(1)
    syn0 = ((rng.random((voc.V, a.dim), dtype=np.float32) - 0.5) / a.dim).astype(np.float32)
    syn1 = np.zeros((voc.V, a.dim), dtype=np.float32)
    from concurrent.futures import ThreadPoolExecutor
    nw = a.ref_walks or 32 * cores

    def one_step(i):
        starts = ((i * nw + np.arange(nw)) % n).astype(np.int32)
        with ThreadPoolExecutor(cores) as ex:
            futs = [ex.submit(oracle.walks_on_the_fly, g, a.p, a.q, starts[k:k + 1], a.walk_length, 1, i * nw + k)
                    for k in range(nw)]
            res = [f.result() for f in futs]
        w = np.concatenate([r[0] for r in res])
        steps = int(sum(int(r[1][0]) - 1 for r in res))
        t1 = time.time()
        tok = np.where(w >= 0, voc.id2index[np.maximum(w, 0)], -1).astype(np.int32)
        off = np.arange(nw + 1, dtype=np.int64) * a.walk_length
        _, _, pairs = oracle.sgns_train(tok, off, voc, dim=a.dim, window=a.window, negative=a.negative,
                                        workers=cores, rng_mode=0, seed=1, syn0=syn0, syn1neg=syn1)
        t2 = time.time()
        # the same walks once more under the GPU arm's default law (one negative set per centre occurrence):
        # like-for-like SGNS rate, reported beside the line's value (which is the reference's own per-pair law)
        _, _, pairs_s = oracle.sgns_train(tok, off, voc, dim=a.dim, window=a.window, negative=a.negative,
                                          workers=cores, rng_mode=2, seed=1, syn0=syn0, syn1neg=syn1)
        shared["pairs"] += pairs_s; shared["s"] += time.time() - t2
        return steps, pairs, t1, t2

    shared = {"pairs": 0, "s": 0.0}
    for i in range(a.warmup):
        one_step(i)
    shared = {"pairs": 0, "s": 0.0}
    steps = pairs = 0
    t_walk = dt = 0.0
    for i in range(a.warmup, a.warmup + a.steps):
        ts = time.time()
        s, p, t1, t2 = one_step(i)
        t_walk += t1 - ts
        dt += t2 - ts                      # the timed step = walks + the reference's own (per-pair) SGNS law
        steps += s; pairs += p
    val = pairs / dt
    sample = (f"{nw} walks per step ({steps} steps, {pairs} pairs in {dt:.1f} s; walk phase {t_walk:.1f} s), "
              f"on-the-fly walker + SGNS port, {cores} threads, full [V={voc.V},{a.dim}] tables")
    print(json.dumps({
        "impl": "reference",
        "metric": "walk steps/s & SGNS pairs/s (value = SGNS pairs/s through the walk->SGNS step)",
        "value": val, "unit": "pairs/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 rows / int32 ids", "data": "synthetic", "config": config_of(a, n, g.nnz),
        "walk_steps_per_s": steps / max(t_walk, 1e-9), "sgns_pairs_per_s_kernel": pairs / max(dt - t_walk, 1e-9),
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "negative_law": "fresh set of 5 per (centre, context) pair (gensim's law, the reference's); the GPU arm's default "
                        "shares one set per centre occurrence -- its per-pair kernel is in its line's other_negative_mode",
        "sgns_pairs_per_s_shared_negative_law": shared["pairs"] / max(shared["s"], 1e-9),
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    args = parse()
    if args.sgns_variant == "mma":
        os.environ["N2V_SGNS_TUNING"] = "16"
    # stdout carries exactly one JSON line: anything libraries print (NCCL banners ...) goes to stderr
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

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
trains the pair bucket (centre in part k, context in part (k + e) % N) in sub-step e and the syn0
parts travel round an NCCL ring -- no replicas (node2vec_by_ecc_b200.word2vec.BlockSgnsTrainer).
`replica` (full copies, delta-sum all-reduce every `sync_walks` walks) and `peer` (one table pair
in NVLink peer memory) are the measured alternatives.
value = (centre, context) pairs trained per second through the whole step, all ranks.
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the default bench step, from the
# committed `ncu --set full` captures (profiles/): filled in when a capture exists, else null
NCU_TRAFFIC = {     # bytes per launch at the default workload (2^19 walks per launch)
    "sgns_train_kernel_v3": 189.8e9,          # profiles/r01_j_sgns_v3_ncu_full.json (91.6 GB read + 98.2 GB written)
    "sgns_train_kernel_v2": None,
    "walk_reject_indexed_kernel": 28.43e9,    # profiles/r01_j_walk_reject_indexed_ncu_full.json
    "walk_reject_kernel": 41.26e9,            # profiles/r01_a_walk_reject_ncu_full.json
}


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

    def step(i, host_io=None, record=False):
        g0 = (i * world + rank) * B                         # global id of this rank's first walk
        if host_io is not None:                              # e2e: inputs from pinned host memory
            hs, hw, hl, hp, dst = host_io
            np.copyto(hs.numpy(), ((g0 + np.arange(B, dtype=np.int64)) % n).astype(np.int32))
            starts = dst.copy_(hs, non_blocking=True)
        else:
            starts = ((g0 + ar) % n).to(torch.int32)
        e0, e1 = ev(), ev()
        e0.record(); do_walk(starts, g0); e1.record()
        if host_io is not None:                              # simulate_walks returns to the host,
            hw.copy_(walks, non_blocking=True); hl.copy_(lens, non_blocking=True)   # learn_embeddings
            walks.copy_(hw, non_blocking=True)               # takes them back in
        # SGNS in sub-batches of `sync_walks`, tables combined (delta-sum) after each (N > 1)
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
        if host_io is not None:
            hp.copy_(trainer.pairs[:1], non_blocking=True)
            torch.cuda.current_stream().synchronize()        # the caller reads the result
        if record:
            kern["walk"].append((e0, e1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(first_step, host_io=None, record=False):
        p0 = trainer.pairs.clone()
        c0 = counters.clone()
        barrier()
        s, e = ev(), ev()
        s.record()
        for i in range(first_step, first_step + a.steps):
            step(i, host_io, record)
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
        hs = torch.empty(B, dtype=torch.int32).pin_memory()
        hw = torch.empty((B, L), dtype=torch.int32).pin_memory()
        hl = torch.empty(B, dtype=torch.int32).pin_memory()
        hp = torch.empty(1, dtype=torch.int64).pin_memory()
        dst = torch.empty(B, dtype=torch.int32, device=dev)
        io = (hs, hw, hl, hp, dst)
        step(a.warmup + a.steps, io)                                        # warm the pinned path
        ms_e, pairs_e, _, _ = timed(a.warmup + a.steps + 1, host_io=io)
        e2e = {"value": pairs_e / (ms_e / 1e3), "unit": "pairs/s",
               "h2d_bytes_per_step": (B * 4 + B * L * 4) * world, "d2h_bytes_per_step": (B * L * 4 + B * 4 + 8) * world,
               "ms_per_step": ms_e / a.steps,
               "api": "DeviceGraph.walk_reject -> host -> %s.train (pinned host buffers)" % type(trainer).__name__}

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
        if block:
            # block kernel: per pair the input row, per carried output row (centre changes + the 5
            # negatives of a run) one read + one reduction: 1,024 B each; `centres` = carried rows
            alg_bytes = (my_pairs_rank + centres / world) * 1024.0
            kname = "sgns_group_kernel (block-partitioned tables, neg_group %d)" % a.neg_group
            k_ms = phases["train"]
        elif a.shared_negatives:
            # shared-negative kernel: per pair the input row (read + written, 1,024 B), per centre
            # the 6 carried output rows (read + written once, 6,144 B)
            alg_bytes = (my_pairs_rank * 1024.0 + centres / world * 6144.0)
            kname = "sgns_train_kernel_v3 (shared negatives)"
        else:
            alg_bytes = my_pairs_rank * float(BYTES_PER_PAIR)
            kname = "sgns_train_kernel_v2 (per-pair negatives)"
        n_launch = max(1, len(kern["sgns"])) * (world if block else 1)
        sg_gbs = alg_bytes / (k_ms / 1e3) / 1e9
        roof = {"kernel": kname, "bound": "hbm", "achieved": sg_gbs, "peak": peak, "unit": "GB/s",
                "frac": sg_gbs / peak, "traffic": NCU_TRAFFIC.get(kname.split()[0]) if (a.scale == 22 and a.batch_walks == 1 << 19 and world == 1) else None,
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)", "peak_source": src,
                "algorithmic_bytes_per_launch": alg_bytes / n_launch,
                "algorithmic_bytes_per_pair": alg_bytes / my_pairs_rank, "pairs_per_launch": my_pairs_rank / n_launch,
                ("carried_rows_per_launch" if block else "centres_per_launch"): centres / world / n_launch,
                "ms_per_launch": k_ms / n_launch,
                "GBps_at_7168B_per_pair": my_pairs_rank * BYTES_PER_PAIR / (k_ms / 1e3) / 1e9}
        if tables is None:
            S, T, P = (float(cn[0]) / world, float(cn[1]) / world, float(cn[3]) / world)
            wbytes = 20 * S + 4 * T + 4 * P
        else:
            S = float(B) * (L - 1) * a.steps
            T = P = 0.0
            wbytes = BYTES_PER_ALIAS_STEP * S
        w_gbs = wbytes / (walk_ms / 1e3) / 1e9
        roof_walk = {"kernel": ("walk_reject_indexed_kernel" if (a.walk_mode == "reject" and a.walk_indexed) else "walk_%s_kernel" % a.walk_mode), "bound": "hbm", "achieved": w_gbs, "peak": peak,
                     "unit": "GB/s", "frac": w_gbs / peak, "traffic": NCU_TRAFFIC.get("walk_reject_indexed_kernel" if (a.walk_mode == "reject" and a.walk_indexed) else "walk_%s_kernel" % a.walk_mode) if (a.scale == 22 and a.batch_walks == 1 << 19 and world == 1) else None,
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
        return steps, pairs, t1

    for i in range(a.warmup):
        one_step(i)
    t0 = time.time()
    steps = pairs = 0
    t_walk = 0.0
    for i in range(a.warmup, a.warmup + a.steps):
        ts = time.time()
        s, p, t1 = one_step(i)
        t_walk += t1 - ts
        steps += s; pairs += p
    dt = time.time() - t0
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
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    args = parse()
    # stdout carries exactly one JSON line: anything libraries print (NCCL banners ...) goes to stderr
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

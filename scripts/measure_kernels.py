"""Kernel-level measurements beside bench.py (CUDA events, warm-up 3, inputs > L2 unless noted):
 * random-access HBM roofline denominators: 32 B sector gathers, 512 B row RMW
 * alias-table build (entries/s) and alias-mode walk (steps/s, 40 B/step) on graphs whose tables fit
 * C2-shaped graph: alias vs rejection walker
Prints one JSON line per measurement."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from node2vec_by_ecc_b200 import DeviceGraph, synth
from node2vec_by_ecc_b200._lib import check, lib, ptr, stream

dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- random-access roofline --------------------------------------------------------------------
buf = torch.zeros(16 << 30, dtype=torch.uint8, device=dev)       # 16 GiB >> 126 MB L2
sink = torch.zeros(1, dtype=torch.int64, device=dev)
n_acc = 1 << 28
ms = timeit(lambda: check(lib().n2v_random_gather_bench(ptr(buf), C.c_size_t(buf.numel()), C.c_int64(n_acc), 0, C.c_uint64(1), ptr(sink), stream())))
emit(what="random 32B-sector gather over 16 GiB", accesses=n_acc, ms=ms, Gsectors_per_s=n_acc / ms / 1e6,
     GBps_sectors=n_acc * 32 / ms / 1e6, frac_of_stream_peak=n_acc * 32 / ms / 1e6 / PEAK)
n_rows = 1 << 25
ms = timeit(lambda: check(lib().n2v_random_gather_bench(ptr(buf), C.c_size_t(buf.numel()), C.c_int64(n_rows), 1, C.c_uint64(1), ptr(sink), stream())))
emit(what="random 512B-row read-modify-write over 16 GiB", rows=n_rows, ms=ms, GBps_read_plus_write=n_rows * 1024 / ms / 1e6,
     frac_of_stream_peak=n_rows * 1024 / ms / 1e6 / PEAK)
del buf
torch.cuda.empty_cache()


# ---- alias build + alias walk ------------------------------------------------------------------
def alias_case(name, lo, hi, n, p, q, R, L):
    dg = DeviceGraph.from_coo(lo, hi, None, n, undirected=True)
    tot = dg.sum_deg_sq()
    if tot < 2e8:
        dg.build_alias_tables(p, q)                 # warm-up: module load, allocator
    torch.cuda.synchronize(); t0 = time.time()
    t = dg.build_alias_tables(p, q)
    torch.cuda.synchronize(); build_s = time.time() - t0
    starts = torch.arange(n, dtype=torch.int32, device=dev).repeat(R)
    walks = torch.empty((starts.shape[0], L), dtype=torch.int32, device=dev)
    lens = torch.empty(starts.shape[0], dtype=torch.int32, device=dev)
    ms_unpacked = timeit(lambda: dg.walk_alias(t, starts, L, 1, 0, out=(walks, lens), packed=False))
    ms = timeit(lambda: dg.walk_alias(t, starts, L, 1, 0, out=(walks, lens)))
    steps = int((lens.to(torch.int64) - 1).sum().item())
    emit(what="alias build + alias walk", graph=name, n=n, nnz=dg.nnz, edge_table_entries=tot, p=p, q=q,
         build_s=build_s, entries_per_s=(tot + dg.nnz) / build_s, slot_GBps=(tot + dg.nnz) * 8 / build_s / 1e9,
         walks=int(starts.shape[0]), L=L, walk_ms=ms, steps_per_s=steps / ms * 1e3,
         steps_per_s_unpacked_form=steps / ms_unpacked * 1e3,
         GBps_at_40B_per_step=steps * 40 / ms / 1e6, frac_of_stream_peak=steps * 40 / ms / 1e6 / PEAK)
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    ms = timeit(lambda: dg.walk_reject(p, q, starts, L, 1, 0, counters=cnt, out=(walks, lens)))
    cnt.zero_(); dg.walk_reject(p, q, starts, L, 1, 0, counters=cnt, out=(walks, lens)); c = cnt.cpu().numpy()
    by = 20 * c[0] + 4 * c[1] + 4 * c[3]
    emit(what="rejection walk", graph=name, walk_ms=ms, steps_per_s=float(c[0]) / ms * 1e3, trials_per_step=float(c[1]) / c[0],
         probes_per_step=float(c[3]) / c[0], GBps_algorithmic=by / ms / 1e6, frac_of_stream_peak=by / ms / 1e6 / PEAK)


lo, hi = synth.planted_edges(10000, 333000, seed=42, device=dev)
alias_case("C2 planted 10k/333k (L2-resident)", lo, hi, 10000, 0.25, 4.0, 100, 80)
lo, hi, n = synth.rmat_edges(20, 4_000_000, seed=1, device=dev)
alias_case("R-MAT scale 20, 4M edges", lo, hi, n, 0.25, 4.0, 10, 80)

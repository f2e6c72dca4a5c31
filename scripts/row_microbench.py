"""Random-access denominators of the SGNS row traffic and of the walker's sector gathers
(n2v_random_gather_bench): LSU forms (ld.global.cg.v4 / red.global.add.v4.f32 per lane) against the
bulk-async forms (cp.async.bulk / cp.reduce.async.bulk, one 512-byte row per instruction, TMA engine).
One JSON line per mode; BYTES=table size (default 2.7 GB = C4's two tables; also 16 GiB)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from node2vec_by_ecc_b200._lib import check, lib, ptr, stream

dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
MODES = {0: "32B sector gather, 4 loads in flight per thread", 8: "32B sector gather, 16 loads in flight per thread",
         2: "512B row read, ld.global.cg.v4 per lane", 5: "512B row read, cp.async.bulk (UBLKCP)",
         3: "512B row reduction, red.global.add.v4.f32 per lane", 6: "512B row reduction, cp.reduce.async.bulk (UBLKRED), constant source",
         7: "512B row reduction, cp.reduce.async.bulk, source rewritten + fence.proxy.async per row",
         1: "512B row read-modify-write (ld + st)"}
only = [int(x) for x in os.environ.get("MODES", "0,8,2,5,3,6,7,1").split(",")]
sizes = [int(float(x)) for x in os.environ.get("BYTES", "2.7e9,17.2e9").split(",")]
reps = int(os.environ.get("REPS", "3"))
sink = torch.zeros(1, dtype=torch.int64, device=dev)
for nbytes in sizes:
    buf = torch.zeros(nbytes // 512 * 512, dtype=torch.uint8, device=dev)
    for mode in only:
        n_acc = (1 << 27) if mode in (0, 8) else (1 << 25)

        def run():
            check(lib().n2v_random_gather_bench(ptr(buf), C.c_size_t(buf.numel()), C.c_int64(n_acc), mode, C.c_uint64(1), ptr(sink), stream()))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        unit = 32 if mode in (0, 8) else (1024 if mode == 1 else 512)
        print(json.dumps({"mode": mode, "what": MODES[mode], "table_GB": nbytes / 1e9, "accesses": n_acc, "ms": ms,
                          "G_accesses_per_s": n_acc / ms / 1e6, "GBps": n_acc * unit / ms / 1e6,
                          "frac_of_stream_peak": n_acc * unit / ms / 1e6 / PEAK}), flush=True)
    del buf
    torch.cuda.empty_cache()

"""Summarise an .ncu-rep (ncu --set full) into a small JSON: the metrics DESIGN.md / bench.py cite.
    python scripts/ncu_summary.py gpurun_out/prof_sgns.ncu-rep profiles/r01_sgns_ncu.json"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum",
        "dram__sectors_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
res = []
for vals in rows[2:]:
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": vals[i], "unit": units[i]}
    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(vals[i])
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1.0
    d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
    res.append(d)
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1)[:1500])

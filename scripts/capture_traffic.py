"""Fill profiles/ncu_traffic.json -- what bench.py's roofline.traffic / frac_dram read -- from `ncu --set
full` reports of bench.py's own launches. Every entry records the sha256 of the kernel's source files at
capture time; bench.py refuses an entry whose sources have changed since.
    python scripts/capture_traffic.py "<workload key>" <report.ncu-rep> [<report.ncu-rep> ...]
workload key = bench.py's wkey, e.g. "scale 22 edges 100000000 batch 524288 L 80 world 1"."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KERNEL_SOURCES, source_sha16

out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(out_path)) if os.path.exists(out_path) else {}
wkey = sys.argv[1]
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    by = {}
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        short = next((k for k in KERNEL_SOURCES if k in name), None)
        if short is None:
            continue
        f = lambda k: float(vals[hdr.index(k)].replace(",", ""))
        unit = lambda k: rows[1][hdr.index(k)]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        b = sum(f(k) * scale[unit(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
        ms = f("gpu__time_duration.sum") * tscale[unit("gpu__time_duration.sum")]
        by.setdefault(short, []).append((b, ms, name))
    for short, lst in by.items():
        table[short] = {"dram_bytes_per_launch": sum(x[0] for x in lst) / len(lst), "launches_captured": len(lst),
                        "ncu_ms_per_launch": sum(x[1] for x in lst) / len(lst), "kernel": lst[0][2][:120],
                        "source_sha16": source_sha16(short), "workload": wkey,
                        "profile": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_ncu_full.json")}
        print(short, table[short])
json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)

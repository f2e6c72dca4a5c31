"""C5 (BASELINE.json configs[4]): walk + SGNS on R-MAT scales 20..24, edge factor 24, one JSON line
per scale (bench.py's line, trimmed). Usage: python scripts/sweep_c5.py [scales...] (1 GPU), or under
torchrun for N GPUs (pass --gpus N through BENCH_ARGS)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scales = [int(x) for x in sys.argv[1:]] or [20, 21, 22, 23, 24]
extra = os.environ.get("BENCH_ARGS", "").split()
for sc in scales:
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--scale", str(sc), "--edges", str(24 * (1 << sc)),
           "--steps", "3", "--warmup", "2", "--no-cpu-baseline", "--no-e2e"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        print(json.dumps({"scale": sc, "error": out.stderr[-400:]}), flush=True)
        continue
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print(json.dumps({"scale": sc, "n_nodes": d["config"]["n_nodes"], "nnz": d["config"]["nnz"], "n_gpus": d["n_gpus"],
                      "pairs_per_s": d["value"], "ms_per_step": d["ms_per_step"], "walk_steps_per_s": d["walk_steps_per_s"],
                      "sgns_pairs_per_s_kernel": d["sgns_pairs_per_s_kernel"], "trials_per_step": d["roofline_walk"]["trials_per_step"],
                      "setup_s": d["setup_s"], "clocks": d["clocks"]}), flush=True)

"""Throughput of the fused all-pairs similarity selection (n2v_sim_threshold, csrc/n2v_score.cu): N x N
cosines of d-dimensional rows with a threshold that lets ~1e-5 of the pairs through. fp32 FMA pipes:
2 * N * N * d flop; peak = SMs * 128 lanes * 2 flop * clock.   N=50000 D=128 python scripts/sim_kernel_bench.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from node2vec_by_ecc_b200.scoring import row_norms, sim_select

n, d = int(os.environ.get("N", "50000")), int(os.environ.get("D", "128"))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
emb = torch.randn((n, d), device=dev, generator=g)
rows = torch.arange(n, dtype=torch.int32, device=dev)
nm = row_norms(emb, rows)
nm = nm + nm
thr = 4.3 / d ** 0.5          # cosines of random rows ~ N(0, 1/d): ~1e-5 of the pairs pass
for _ in range(2):
    a, b, s = sim_select(emb, rows, rows, thr=thr, upper_only=False, norms=nm)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a, b, s = sim_select(emb, rows, rows, thr=thr, norms=nm); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[1]
flop = 2.0 * n * n * d
sms = torch.cuda.get_device_properties(0).multi_processor_count
print(json.dumps({"n": n, "d": d, "ms": ms, "pairs_emitted": int(a.numel()), "TFLOP_per_s_fp32": flop / ms / 1e9,
                  "scores_per_s": n * n / ms * 1e3, "fp32_simt_peak_TFLOP_per_s_at_1.9GHz": sms * 128 * 2 * 1.9e9 / 1e12}))

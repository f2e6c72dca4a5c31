"""Small driver for ncu captures of the alias kernels: R-MAT scale 20 / 4 M edges (tables 45 GB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from node2vec_by_ecc_b200 import DeviceGraph, synth
dev = torch.device("cuda", 0)
lo, hi, n = synth.rmat_edges(20, 4_000_000, seed=1, device=dev)
dg = DeviceGraph.from_coo(lo, hi, None, n, undirected=True)
t = dg.build_alias_tables(0.25, 4.0)
starts = torch.arange(n, dtype=torch.int32, device=dev).repeat(10)
for _ in range(3):
    walks, lens = dg.walk_alias(t, starts, 80, 1, 0)
torch.cuda.synchronize()
print("ok", dg.sum_deg_sq(), int((lens.long() - 1).sum()))

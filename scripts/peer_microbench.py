"""Random 512-byte-row traffic to LOCAL vs PEER (NVLink) memory: the denominators of the sharded-table
SGNS. Run: torchrun --nproc-per-node 2 scripts/peer_microbench.py"""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from node2vec_by_ecc_b200._lib import check, lib, stream
from node2vec_by_ecc_b200.dist import exchange_peer_pointers

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
NB = 1 << 30                                    # 1 GiB of rows per GPU
buf = torch.zeros(NB // 4, dtype=torch.float32, device=dev)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
ptrs = exchange_peer_pointers(buf)
N = 1 << 24
names = {1: "row rmw (ld + st)", 2: "row read", 3: "row red.v4.f32", 4: "row red.f32 x4"}
res = {}
def run(target, mode, both):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if both or rank == 0:
        for rep in range(2):
            e0.record()
            check(lib().n2v_random_gather_bench(C.c_void_p(ptrs[target]), C.c_size_t(NB), C.c_int64(N), mode, C.c_uint64(rep + 1),
                                                C.c_void_p(sink.data_ptr()), stream()))
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        ms = float("inf")
    dist.barrier()
    return N / ms / 1e6, N * 512 / ms / 1e6      # G rows/s, GB/s
for mode in (2, 3, 4, 1):
    for where, both in (("local", False), ("peer", False), ("peer, both directions at once", True)):
        tgt = rank if where == "local" else (rank + 1) % world
        r = run(tgt, mode, both)
        if rank == 0:
            res["%s -> %s" % (names[mode], where)] = {"Grows_per_s": round(r[0], 3), "GBps": round(r[1], 1)}
            print(names[mode], where, r, flush=True)
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()

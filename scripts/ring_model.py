"""Timing model of the block-partitioned SGNS ring: N GPUs, N sub-steps; GPU k trains
bucket (k, (k + e) % N) in sub-step e, then the syn0 parts move one GPU down the ring. A bucket costs
`hot` ms when it holds syn0 part 0 (the part with the most frequent word), `base` ms otherwise
(profiles/r01_v_*: 25 vs 17 ms at the 8-GPU bucket size). Two pass protocols:
  rendezvous  a GPU continues when itself and both ring neighbours have finished the sub-step
              (one grouped NCCL send/recv pair, what dist.ring_pass does)
  decoupled   a GPU continues when itself and the GPU it receives from have finished
Both give N * hot: the hot part is trained by the N GPUs one after another, it IS the critical path.
    python scripts/ring_model.py [N base hot xfer]"""
import sys

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
base, hot, xfer = (float(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (17.0, 25.0, 0.4)


def step_ms(dur, decoupled):
    end = [0.0] * N
    for e in range(N):
        fin = [end[k] + dur(k, e) for k in range(N)]
        if decoupled:
            end = [max(fin[k], fin[(k + 1) % N] + xfer) for k in range(N)]
        else:
            end = [max(fin[k], fin[(k + 1) % N], fin[(k - 1) % N]) + xfer for k in range(N)]
    return max(end)


cases = {"hot part 0": lambda k, e: hot if (k + e) % N == 0 else base,
         "one slow GPU": lambda k, e: hot if k == N // 2 else base,
         "uniform": lambda k, e: base}
for name, d in cases.items():
    print("%-13s rendezvous %6.1f ms   decoupled %6.1f ms" % (name, step_ms(d, False), step_ms(d, True)))

cd $GRAFT_REPO_ROOT
python scripts/hot_rows_sweep.py 2>&1 | grep "hot default"
for C in 0.5 1 2; do
  N2V_SGNS_HOT_COPIES=$C timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('copies $C', 'value', d['value'], 'kernel', d['sgns_pairs_per_s_kernel'])"
done
N2V_SGNS_HOT_ROWS=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('hot 0', 'value', d['value'], 'kernel', d['sgns_pairs_per_s_kernel'])"
python - <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from node2vec_by_ecc_b200 import DeviceGraph, SgnsTrainer, synth
lo, hi, n = synth.rmat_edges(22, int(100e6), seed=1, device='cuda')
deg = torch.bincount(torch.cat([lo, hi]).long(), minlength=n)
tr = SgnsTrainer(deg, dim=128)
for c in ('0.3', '0.5', '1', '2', '8'):
    import os; os.environ['N2V_SGNS_HOT_COPIES'] = c
    h = tr.hot_rows(2960); print('C4 copies', c, 'hot rows', h, 'mass', float(tr._neg_prob[:h].sum()))
PY

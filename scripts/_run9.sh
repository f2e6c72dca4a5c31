cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for C in 0.5; do
  N2V_SGNS_HOT_COPIES=$C timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('copies $C', 'value', d['value'], 'kernel', d['sgns_pairs_per_s_kernel'])"
done
N2V_SGNS_HOT_ROWS=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('hot 0', 'value', d['value'], 'kernel', d['sgns_pairs_per_s_kernel'])"
PARTS=1,8 timeout 600 python scripts/block_throughput.py 2>&1 | cut -c1-300

cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_sgns.py -m gpu -q -k "tensor_core" 2>&1 | tail -5
bash scripts/_t.sh
for v in default mma; do timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --sgns-variant $v 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('$v', d['value'], d['sgns_pairs_per_s_kernel'], d['roofline']['kernel'][:40])"; done

cd $GRAFT_REPO_ROOT
python scripts/hot_rows_sweep.py 2>&1 | grep "hot default"
python -m pytest tests/test_gpu_auc.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r02_p_bench_n1.json
python -c "
import json; d = json.load(open('gpurun_out/r02_p_bench_n1.json')); print('default', 'value', d['value'], 'kernel', d['sgns_pairs_per_s_kernel'], 'e2e', d['e2e']['value'])"

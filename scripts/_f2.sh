cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_alias.py tests/test_gpu_walks.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -3
python scripts/measure_kernels.py 2>&1 | grep "alias build" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['graph'], 'build_s', round(d['build_s'], 4), 'entries/s', d['entries_per_s'])"

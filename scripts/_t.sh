cd $GRAFT_REPO_ROOT
MODES=2,3,1 BYTES=33.5e6,134e6,537e6 python scripts/row_microbench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['mode'], d['table_GB'], round(d['G_accesses_per_s'], 2), 'G rows/s')
"
python scripts/e2e_c2_api.py 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_r_bench_n1_driver_style.json 2>/dev/null ) 2>&1 | grep real

cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -3
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --sgns-variant mma"
$B > gpurun_out/plain_m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgns_train_kernel_mma -s 3 -c 1 -o gpurun_out/r02_s_sgns_mma $B > gpurun_out/ncu_m.log 2>&1
N2V_SGNS_TUNING=16 timeout 600 python scripts/auc_c2.py 2>&1 | tail -1 > gpurun_out/r02_s_auc_c2_mma.json
ls -la gpurun_out/r02_s*

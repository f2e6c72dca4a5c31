set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_m_bench_n1.json 2> gpurun_out/r02_m_bench_n1.err; tail -2 gpurun_out/r02_m_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_m_bench_reference.json 2> gpurun_out/r02_m_bench_reference.err; tail -2 gpurun_out/r02_m_bench_reference.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_m_launches.csv $B > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sgns_train_kernel_v3 -s 3 -c 1 -o gpurun_out/r02_m_sgns_v3 $B > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:walk_reject_indexed -s 9 -c 1 -o gpurun_out/r02_m_walk_reject_indexed $B > gpurun_out/ncu_c.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --shared-negatives 0"
$B2 > gpurun_out/plain_d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgns_train_kernel_v2 -s 3 -c 1 -o gpurun_out/r02_m_sgns_v2 $B2 > gpurun_out/ncu_d.log 2>&1
POOL=4194304 PARTS=8 timeout 900 python scripts/block_throughput.py > gpurun_out/r02_m_block_throughput_pool8.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgns_group_kernel -s 130 -c 1 -o gpurun_out/r02_m_sgns_group_parts8 env POOL=4194304 PARTS=8 python scripts/block_throughput.py > gpurun_out/ncu_e.log 2>&1
MODES=0,8,2,3 BYTES=17.2e9 REPS=1 python scripts/row_microbench.py > gpurun_out/plain_f.log 2>&1 &&
ncu --metrics dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sectors.sum,gpu__time_duration.sum --clock-control none -k regex:"gather_sector|row_half" --csv --log-file gpurun_out/r02_m_microbench_sectors.csv env MODES=0,8,2,3 BYTES=17.2e9 REPS=1 python scripts/row_microbench.py > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep

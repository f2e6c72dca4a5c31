set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_e_pytest.txt
cat gpurun_out/r02_e_pytest.txt | tail -5
PARTS=1,8 NEG_GROUPS=1 timeout 900 python scripts/block_throughput.py > gpurun_out/r02_e_block_throughput_minb5.txt 2>&1
PARTS=8 NEG_GROUPS=16,64 timeout 900 python scripts/block_throughput.py >> gpurun_out/r02_e_block_throughput_minb5.txt 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_e_bench_n1.json 2> gpurun_out/r02_e_bench_n1.err
# the group kernel at 128 registers / 4 blocks per SM
N2V_NVCC_FLAGS="-DN2V_BLK_MINB=4" python -m node2vec_by_ecc_b200.build --force > gpurun_out/build_minb4.log 2>&1
PARTS=1,8 NEG_GROUPS=1 timeout 900 python scripts/block_throughput.py > gpurun_out/r02_e_block_throughput_minb4.txt 2>&1
python -m node2vec_by_ecc_b200.build --force > /dev/null 2>&1
# ncu: two bucket launches of the 8-part step, two launches of the 1-part step
PARTS=8 NEG_GROUPS=1 timeout 600 python scripts/block_throughput.py > gpurun_out/plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgns_group_kernel -s 150 -c 2 -o gpurun_out/r02_e_group_kernel_parts8 env PARTS=8 NEG_GROUPS=1 python scripts/block_throughput.py > gpurun_out/ncu8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sgns_group_kernel -s 2 -c 1 -o gpurun_out/r02_e_group_kernel_parts1 env PARTS=1 NEG_GROUPS=1 python scripts/block_throughput.py > gpurun_out/ncu1.log 2>&1
ls -la gpurun_out/*.ncu-rep

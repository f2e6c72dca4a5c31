set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_d_pytest.txt
cat gpurun_out/r02_d_pytest.txt
SEEDS=3 timeout 900 python scripts/auc_block.py > gpurun_out/r02_d_auc_block_groups.txt 2>&1
PARTS=1,2,4,8 GROUPS=1 timeout 900 python scripts/block_throughput.py > gpurun_out/r02_d_block_throughput.txt 2>&1
PARTS=8 GROUPS=4,8 timeout 900 python scripts/block_throughput.py >> gpurun_out/r02_d_block_throughput.txt 2>&1

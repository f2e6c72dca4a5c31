set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_b_pytest.txt
# blocked distribution + negative groups of G token positions (tuning = G<<8 | 8 | 4 | 1)
for G in 1 2 4 8; do
  T=$(( (G<<8) | 13 ))
  N2V_BLK_TUNING=$T SEEDS=3 GRID="1,32,2048;2,32,4096;4,32,8192;8,32,16384;8,32,4096;1,32,2048,64;8,32,16384,64" timeout 900 python scripts/auc_block.py > gpurun_out/r02_b_auc_block_G$G.txt 2>&1
done
# blocked distribution, run-shared negatives (no groups)
N2V_BLK_TUNING=9 SEEDS=3 GRID="1,32,2048;2,32,4096;8,32,16384;1,8,2048;8,8,16384;8,4,16384" timeout 900 python scripts/auc_block.py > gpurun_out/r02_b_auc_block_blocked_runs.txt 2>&1
# strided distribution + groups
N2V_BLK_TUNING=$(( (1<<8) | 5 )) SEEDS=3 GRID="1,32,2048;8,32,16384" timeout 900 python scripts/auc_block.py > gpurun_out/r02_b_auc_block_G1_strided.txt 2>&1
cat gpurun_out/r02_b_pytest.txt

set -x
cd $GRAFT_REPO_ROOT
# C2: negative groups wider than the part count (sets shared by ~G/parts centres of a part)
SEEDS=3 GRID="8,16,4096;8,32,4096;8,64,4096;4,8,4096;4,16,4096;2,4,4096;2,8,4096;1,2,4096;1,4,4096" timeout 900 python scripts/auc_block.py > gpurun_out/r02_l_auc_block_wide_groups.txt 2>&1
PARTS=8 NEG_GROUPS=16,32 timeout 600 python scripts/block_throughput.py > gpurun_out/r02_l_block_throughput_groups.txt 2>&1
PARTS=1,2,4,8 NEG_GROUPS=1 timeout 1500 python scripts/auc_large_anchor.py > gpurun_out/r02_l_auc_large_anchor.jsonl 2> gpurun_out/r02_l_auc_large_anchor.err
PARTS=8 NEG_GROUPS=16,32 ORACLE=0 timeout 600 python scripts/auc_large_anchor.py > gpurun_out/r02_l_auc_large_groups.jsonl 2>> gpurun_out/r02_l_auc_large_anchor.err
tail -3 gpurun_out/r02_l_auc_large_anchor.err

set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/row_microbench.py > gpurun_out/r02_a_row_microbench.jsonl 2> gpurun_out/r02_a_row_microbench.err
echo "microbench rc=$?"
SEEDS=3 GRID="1,1,2048;2,1,4096;4,1,8192;8,1,16384;1,32,2048;2,32,4096;4,32,8192;8,32,16384;1,32,2048,128;8,32,16384,128;8,1,16384,128;2,1,4096,16;1,1,2048,16" timeout 900 python scripts/auc_block.py > gpurun_out/r02_a_auc_block_t0.txt 2>&1
echo "t0 rc=$?"
N2V_BLK_TUNING=5 SEEDS=3 GRID="1,32,2048;2,32,4096;4,32,8192;8,32,16384;1,32,512;8,32,4096;1,32,2048,128;8,32,16384,128;2,32,4096,16;1,32,2048,16" timeout 900 python scripts/auc_block.py > gpurun_out/r02_a_auc_block_t5.txt 2>&1
echo "t5 rc=$?"
N2V_BLK_TUNING=1 SEEDS=3 GRID="2,32,4096;8,32,16384" timeout 600 python scripts/auc_block.py > gpurun_out/r02_a_auc_block_t1.txt 2>&1
echo "t1 rc=$?"
tail -3 gpurun_out/r02_a_row_microbench.err

set -x
cd $GRAFT_REPO_ROOT
N2V_BLK_TUNING=$(( (1<<8) | 13 )) SEEDS=3 GRID="1,32,16384;1,32,8192;1,32,512;2,32,2048;2,32,1024;4,32,2048;4,32,1024;8,32,2048;8,32,1024;8,32,512" timeout 900 python scripts/auc_block.py > gpurun_out/r02_c_auc_block_G1_pools.txt 2>&1
N2V_BLK_TUNING=$(( (4<<8) | 13 )) SEEDS=3 GRID="4,32,1024;4,32,2048" timeout 900 python scripts/auc_block.py > gpurun_out/r02_c_auc_block_G4_pools.txt 2>&1
N2V_BLK_TUNING=$(( (8<<8) | 13 )) SEEDS=3 GRID="8,32,1024;8,32,2048" timeout 900 python scripts/auc_block.py > gpurun_out/r02_c_auc_block_G8_pools.txt 2>&1

#!/bin/bash
mkdir -p gpurun_out
(timeout -k 10 900 python -m pytest tests/test_gpu_tfidf.py -m gpu -q -x 2>&1 | tail -5) | tee gpurun_out/r2_blk3_tests.log
for cfg in "8 2048 16" "4 2048 16"; do
  set -- $cfg; echo "rows=$1 tile=$2 bits=$3"; PFZ_BLOCK_ROWS=$1 PFZ_BLOCK_ACC_BITS=$3 timeout -k 10 300 python tools/k2_sweep.py 100000 $2 block 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_blk3_sweep.log
timeout -k 10 300 python tools/profile_hash.py 1000000 100000 3 2>&1 | tail -1

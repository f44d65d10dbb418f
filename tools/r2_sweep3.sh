#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 1024 32" "8 2048 16" "8 4096 16" "16 1024 32" "16 2048 16" "8 1024 16" "8 512 32"; do
  set -- $cfg; echo "rows=$1 tile=$2 bits=$3"; PFZ_BLOCK_ROWS=$1 PFZ_BLOCK_ACC_BITS=$3 timeout -k 10 300 python tools/k2_sweep.py 100000 $2 block 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_sweep3.log

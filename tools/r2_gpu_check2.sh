#!/bin/bash
mkdir -p gpurun_out
(timeout -k 10 900 python -m pytest tests/test_gpu_tfidf.py -m gpu -x -q -k "block or ties or company_slice_topk or synthetic_two" 2>&1 | tail -8) | tee gpurun_out/r2_t5.log
for cfg in "8 1024 32" "8 2048 16" "8 1024 16" "16 2048 16" "8 4096 16" "16 1024 32"; do
  set -- $cfg; echo "rows=$1 tile=$2 bits=$3"; PFZ_BLOCK_ROWS=$1 PFZ_BLOCK_ACC_BITS=$3 timeout -k 10 300 python tools/k2_sweep.py 100000 $2 block 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_sweep2.log
timeout -k 10 300 python tools/e2e_profile.py 2>&1 | tail -4 | tee gpurun_out/r2_e2e_profile2.txt
for s in 8192 16384; do PFZ_HASH_SLOTS=$s timeout -k 10 300 python tools/profile_hash.py 1000000 100000 2 2>&1 | tail -2; done | tee gpurun_out/r2_hash_time.log
ncu --set full --clock-control none --import-source on -k regex:spcos_hash -s 1 -c 1 -o gpurun_out/r2_k2_hash_v1 -f python tools/profile_hash.py 1000000 20000 2 > gpurun_out/r2_ncu_hash.log 2>&1; tail -2 gpurun_out/r2_ncu_hash.log
(timeout -k 10 300 python -m pytest tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -3) | tee -a gpurun_out/r2_t5.log

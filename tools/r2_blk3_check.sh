#!/bin/bash
mkdir -p gpurun_out
(timeout -k 10 600 python -m pytest tests/test_gpu_tfidf.py -m gpu -q -x -k "block or ties or company_slice_topk or synthetic" 2>&1 | tail -3) | tee gpurun_out/r2_blk3_tests.log
for d in 1 0; do echo "draw=$d"; PFZ_BLOCK_DRAW=$d timeout -k 10 300 python tools/k2_sweep.py 100000 2048 block 2>&1 | tail -1; done 2>&1 | tee gpurun_out/r2_blk3_sweep.log

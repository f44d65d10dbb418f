#!/bin/bash
mkdir -p gpurun_out
(timeout -k 10 900 python -m pytest tests/test_gpu_tfidf.py -m gpu -q -x 2>&1 | tail -3) | tee gpurun_out/r2_blk3_tests.log
timeout -k 10 300 python tools/k2_sweep.py 100000 2048 block 2>&1 | tail -1 | cut -c1-125

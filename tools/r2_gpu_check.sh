#!/bin/bash
# Developer tool: one GPU call that exercises the round-2 kernels (tests per file, K2 sweep, bench); logs under gpurun_out/.
mkdir -p gpurun_out
for f in test_gpu_tfidf test_gpu_fuzz test_gpu_editdist test_gpu_dropin test_gpu_dense; do
  echo "=== $f"; (timeout -k 10 900 python -m pytest tests/$f.py -m gpu -x -q 2>&1 | tail -15) | tee gpurun_out/r2_$f.log
done
for cfg in "8 1024" "8 2048" "16 1024" "16 512" "8 512"; do
  set -- $cfg; echo "rows=$1 tile=$2"; PFZ_BLOCK_ROWS=$1 timeout -k 10 300 python tools/k2_sweep.py 100000 $2 block 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_sweep.log
(timeout -k 10 600 python bench.py --steps 5 --warmup 3) > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -3 gpurun_out/r2_bench3.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench3.json").read().strip().splitlines()[-1])
    print("step", d["ms_per_step"], d["k2"], "k1", d["k1_ms_avg"], "k2", d["k2_ms_avg"], d["step_ms_each"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("frame_tail"))
    for c in ("c3","c4","c5"):
        if c in d: print(c, d[c]["ms"], d[c].get("k2_variant"), d[c]["roofline"]["frac"])
except Exception as e: print("bench parse failed", e)
PY

#!/bin/bash
# Developer tool: one GPU call = full GPU test suite, bench line, ncu launch list, ncu --set full of the dominant kernel.
mkdir -p gpurun_out
(timeout -k 10 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25) | tee gpurun_out/r2_pytest_gpu.log
(timeout -k 10 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
    print("step", d["ms_per_step"], d["k2"], "k1", d["k1_ms_avg"], "k2", d["k2_ms_avg"], d["step_ms_each"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("frame_tail"))
    for c in ("c3","c4","c5"):
        if c in d: print(c, d[c]["ms"], d[c].get("k2_variant"), d[c]["roofline"]["frac"])
except Exception as e: print("bench parse failed", e)
PY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --skip c3,c4,c5,e2e > gpurun_out/r2_launches_bench.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:spcos_blk3_kernel -s 1 -c 1 -o gpurun_out/r2_k2_blk3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip c3,c4,c5,e2e > gpurun_out/r2_ncu_block.log 2>&1; tail -2 gpurun_out/r2_ncu_block.log

#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -k "not n50k" ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2_bench_flow.json 2> gpurun_out/r2_bench_flow.err; echo "bench rc=$?" >> gpurun_out/r2_bench_flow.err
( timeout 600 python tools/pool_bench.py ) > gpurun_out/r2_pool.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_bench_flow.err
python - <<'PY'
import json
for f in ("flow",):
    try:
        d=json.loads(open("gpurun_out/r2_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phases_ms"], d["repro"]["mismatches"])
    except Exception as e: print(f, "failed", e)
PY
grep "evals/s" gpurun_out/r2_pool.log

#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -x -q -m gpu -s -k "sampled" ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "sampled factor|passed|failed|Error" gpurun_out/r2_pytest_gpu.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["phases_ms"], d["parity"]["rel_err"], d["repro"]["mismatches"], d["e2e"]["value"])
print(json.dumps(d.get("north_star_n100k_1gpu"), indent=0))
print(d["cpu_baseline"]["value"])
PY
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/r2_bench_ref.err; cut -c1-700 gpurun_out/r2_bench_ref.json

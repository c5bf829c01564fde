#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2_bench_ref.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["phases_ms"], d["parity"]["rel_err"], d["repro"]["mismatches"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline"]["frac"], d["roofline"]["cholesky_phase"])
print("ref", r["value"], r["ms_per_step"], r["measured_full_eval"])
print("ratio e2e", d["e2e"]["value"]/r["value"])
PY

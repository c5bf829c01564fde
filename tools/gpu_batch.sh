#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-large ) > gpurun_out/r2_bench_n1s.json 2> gpurun_out/r2_bench_n1s.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1s.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1s.json").read().strip().splitlines()[-1])
print(d["value"], d["phases_ms"], d["parity"]["rel_err"], d["repro"]["mismatches"], d["e2e"]["value"])
print(json.dumps(d.get("reference_datasets"), indent=0))
PY

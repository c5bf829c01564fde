#!/bin/bash
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1s.json 2> gpurun_out/r2_bench_n1s.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1s.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n1s.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['cholesky_phase'])
"

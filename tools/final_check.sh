#!/bin/bash
# what the driver runs at round end, on one B200: GPU tests, smoke, both bench arms with its arguments
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/final_pytest.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/final_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['repro']['mismatches'], d['parity']['max_rel_err'], d['roofline']['frac'], d['north_star_n100k_1gpu']['chol_tflops'], d['north_star_n100k_1gpu']['factor_residual']['max_rel'])
"

#!/bin/bash
# flakiness hunt before the round ends: everything the driver runs, several times
mkdir -p gpurun_out
for r in 1 2 3; do ( timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_rep$r.log 2>&1; echo "pytest rep $r rc=$? $(tail -1 gpurun_out/r2_pytest_rep$r.log)"; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for r in 1 2; do python bench.py --gpus 1 --steps 20 --warmup 5 --no-large > gpurun_out/r2_bench_rep$r.json 2> gpurun_out/r2_bench_rep$r.err; echo "bench rep $r rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_rep$r.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['repro'], d['parity']['max_rel_err'])
" | cut -c1-200; done
timeout 300 tools/micro/bin/chol_race 4 12032 40 | tail -1
timeout 300 tools/micro/bin/chol_race_nopf 4 12032 40 | tail -1
timeout 600 tools/micro/bin/chol_race 1 50048 20 1 | tail -1

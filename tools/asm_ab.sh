#!/bin/bash
# A/B of assembly-kernel builds on one B200: timing at n = 50 000 (and the holes size), then the entry-parity tests.
# usage: bash tools/asm_ab.sh [lib.so ...]   ("-" = the shipped library)
mkdir -p gpurun_out
for lib in "${@:--}"; do
  [ "$lib" = "-" ] || [ -f "$lib" ] || continue
  python tools/asm_time.py $lib 50000 2>&1 | tail -1
  python tools/asm_time.py $lib 5570 2>&1 | tail -1
done | tee gpurun_out/asm_ab.log
( timeout 900 python -m pytest tests/test_gpu_cov.py tests/test_gpu_taper.py tests/test_gpu_n2ll.py tests/test_gpu_predict_sim.py -q -m gpu ) > gpurun_out/asm_ab_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/asm_ab_pytest.log)"

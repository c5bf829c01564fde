#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_asm_variants.log; : > $OUT
timeout 300 python tools/asm_time.py - 30000 >> $OUT 2>&1; timeout 300 python tools/asm_time.py - 5570 >> $OUT 2>&1; timeout 300 python tools/asm_time.py - 50000 >> $OUT 2>&1
grep ASM_TIME $OUT
( time timeout 1200 python -m pytest tests -x -q -m gpu -k "cov or taper or n2ll" ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log

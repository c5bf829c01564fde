#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_band.log; : > $OUT
for h in 16 24 32; do echo "=== band $h" >> $OUT; timeout 300 tools/micro/bin/gemm_time_band$h 32768 768 50048 4 >> $OUT 2>&1; timeout 300 tools/micro/bin/gemm_time_band$h 49152 768 12032 4 >> $OUT 2>&1; done
for o in 2 3 4 6; do echo "=== COCONS_CHOL_OUTER=$o" >> $OUT; COCONS_CHOL_OUTER=$o timeout 300 python tools/pool_bench.py 2>&1 | grep -E "in_flight=(1|8)" >> $OUT; done
cat $OUT

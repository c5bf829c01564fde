#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_tpc.log; : > $OUT
for t in 1 2 4 8; do echo "=== COCONS_GEMM_TPC=$t" >> $OUT; COCONS_GEMM_TPC=$t timeout 300 tools/micro/bin/gemm_time 32768 768 50048 3 >> $OUT 2>&1; done
echo "=== default" >> $OUT; timeout 300 tools/micro/bin/gemm_time 32768 768 50048 3 >> $OUT 2>&1
timeout 300 tools/micro/bin/gemm_time 16384 512 12032 3 >> $OUT 2>&1
timeout 300 tools/micro/bin/chol_race 4 12032 20 >> $OUT 2>&1
timeout 300 tools/micro/bin/chol_race_nopf 1 12032 40 >> $OUT 2>&1
COCONS_GEMM_TPC=4 timeout 300 tools/micro/bin/chol_race_nopf 4 12032 20 >> $OUT 2>&1
timeout 300 tools/micro/bin/chol_race 1 50048 6 1 >> $OUT 2>&1
grep -E "^===|GEMM_TIME|SUMMARY" $OUT
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log

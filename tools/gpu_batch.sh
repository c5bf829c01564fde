#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2_cs.log; : > $OUT
timeout 300 tools/micro/bin/gemm_time 32768 768 50048 4 >> $OUT 2>&1
timeout 300 tools/micro/bin/chol_race 4 12032 10 >> $OUT 2>&1
timeout 600 python tools/pool_bench.py 2>&1 | grep -E "in_flight=(1|8)" >> $OUT
grep -E "GEMM_TIME|SUMMARY|evals" $OUT
CMD="python bench.py --profile --steps 1 --warmup 0"
$CMD > gpurun_out/p_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_nt_tma_kernelILi64 -s 11 -c 1 -f -o gpurun_out/r2_gemm_cs $CMD > gpurun_out/p_ncu_g.log 2>&1
echo "gemm ncu rc=$?"; tail -1 gpurun_out/p_plain.log

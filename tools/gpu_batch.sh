#!/bin/bash
# one gpurun call: GPU tests, reproducibility soak, pooled evaluations, bench.py with the driver's arguments
mkdir -p gpurun_out
B=tools/micro/bin
( time python -m pytest tests -x -q -m gpu -k "not n50k" ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
OUT=gpurun_out/r2_soak.log; : > $OUT
run() { echo "=== $*" >> $OUT; ( time timeout 600 "$@" ) >> $OUT 2>&1; echo "rc=$?" >> $OUT; }
run $B/chol_race_nopf_rel0 1 12032 5
run $B/chol_race_nopf 1 12032 100
run $B/chol_race_nopf 4 12032 30
run $B/chol_race 4 12032 30
run $B/chol_race 4 20096 10
run $B/chol_race 1 50048 12 1
run $B/gemm_time 32768 768 50048 4
( COCONS_DEBUG_CHECKSUM=1 python tools/pool_stress.py 4 stripes 8; COCONS_DEBUG_CHECKSUM=1 python tools/pool_stress.py 4 holes 8; python tools/pool_bench.py ) > gpurun_out/r2_pool.log 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n1.err
tail -3 gpurun_out/r2_pytest_gpu.log; grep -E "^===|SUMMARY|GEMM_TIME" $OUT; grep -E "POOL_STRESS|evals/s" gpurun_out/r2_pool.log; tail -5 gpurun_out/r2_bench_n1.err; cut -c1-1500 gpurun_out/r2_bench_n1.json

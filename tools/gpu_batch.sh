#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/pool_bench.py 2>&1 | grep -E "in_flight=(1|8)"
COCONS_DEBUG_CHECKSUM=1 timeout 600 python tools/pool_stress.py 3 stripes 8 2>&1 | tail -1
CMD="python tools/asm_time.py - 50000"
$CMD > gpurun_out/p_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:assemble_lower -s 1 -c 1 -f -o gpurun_out/r2_asm3 $CMD > gpurun_out/p_ncu_a.log 2>&1
echo "asm ncu rc=$?"; tail -1 gpurun_out/p_plain.log

#!/bin/bash
# ncu evidence for profiles/: launch list of one evaluation + full captures of the three main kernels
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --warmup 0"
$CMD > gpurun_out/p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/p_ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_nt_tma_kernelILi64 -s 11 -c 1 -f -o gpurun_out/r2_gemm $CMD > gpurun_out/p_ncu_g.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fwd_solve_flow -c 1 -f -o gpurun_out/r2_solve $CMD > gpurun_out/p_ncu_s.log 2>&1
echo "solve rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:assemble_lower -c 1 -f -o gpurun_out/r2_asm $CMD > gpurun_out/p_ncu_a.log 2>&1
echo "asm rc=$?"
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/p_plain.log; wc -l gpurun_out/r2_launches.csv

#!/bin/bash
# round-2 batch 3: candidate fixes of the slot-release race (correctness under the no-prefetch amplifier, speed)
B=tools/micro/bin
OUT=gpurun_out/race_batch3.log
mkdir -p gpurun_out
: > $OUT
run() { echo "=== $*" >> $OUT; ( time timeout 600 "$@" ) >> $OUT 2>&1; echo "rc=$?" >> $OUT; }
for r in 0 1 2 3; do
  run $B/chol_race_np_rel$r 1 12032 20
  run $B/chol_race_np_rel$r 4 12032 10
done
for r in 1 2 3; do
  run $B/chol_race_rel$r 1 50048 6 1
  run $B/chol_race_rel$r 4 12032 10
done
for r in 0 1 2 3; do
  run $B/gemm_time_rel$r 32768 768 50048 4
done
grep -E "^===|SUMMARY|GEMM_TIME" $OUT

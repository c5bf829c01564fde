#!/bin/bash
mkdir -p gpurun_out
echo "=== old"; timeout 120 tools/micro/bin/potrf_check_old 2>&1 | grep -E "variant 2|blocked kernel"
echo "=== new"; timeout 120 tools/micro/bin/potrf_check 2>&1 | grep -E "variant 2|blocked kernel|non-PD"

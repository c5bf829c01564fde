#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu -k "predict_matches" ) > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log

#!/bin/bash
# Round-2 visit B (1 GPU): the new subsystems first (fail fast), then the whole GPU suite.
set -u
mkdir -p gpurun_out
export SVSB_XCHG_TIMEOUT_MS=8000
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_mutate.py -x -q > gpurun_out/r2b_new.log 2>&1; echo "new rc=$?"; tail -25 gpurun_out/r2b_new.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest_gpu.log

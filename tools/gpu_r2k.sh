#!/bin/bash
# recovery block launched after the first pass (all SMs to the first pass) against the reserved SM: cfg 3 warm (no failures),
# cfg 3 from the reference's start (two thirds fail), cfg 5 (29 % fail); then the tests that exercise recovery and launches
set -x
mkdir -p gpurun_out
for e in 0 1; do
  echo "== OBCA_B200_RESERVE_SM=$e"
  OBCA_B200_RESERVE_SM=$e timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -3
  OBCA_B200_RESERVE_SM=$e OBCA_QUICK_INIT=0 timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
  OBCA_B200_RESERVE_SM=$e timeout 300 python tools/gpu_quick.py 5 8192 2>&1 | tail -2
done 2>&1 | tee gpurun_out/ab_reserve.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300

#!/bin/bash
# lockstep groups: throughput at 3 / 2 / 1 instances per block, memcheck, tests
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_plain.log 2>&1; tail -4 gpurun_out/quick_plain.log
OBCA_GROUPS=1 timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
OBCA_GROUPS=2 timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
timeout 300 python tools/gpu_quick.py 5 8192 > gpurun_out/quick_cfg5.log 2>&1; tail -3 gpurun_out/quick_cfg5.log
timeout 300 python tools/gpu_quick.py 2 1024 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_quick.py 3 600 > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/gpu_quick.py 3 450 > gpurun_out/racecheck_3.log 2>&1; tail -3 gpurun_out/racecheck_3.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-400

#!/bin/bash
# bulk-copy staging: tests, memcheck on a small batch, throughput
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_quick.py 3 600 > gpurun_out/memcheck.log 2>&1; tail -6 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/gpu_quick.py 3 450 > gpurun_out/racecheck_3.log 2>&1; tail -4 gpurun_out/racecheck_3.log
timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_plain.log 2>&1; tail -4 gpurun_out/quick_plain.log
timeout 300 python tools/gpu_quick.py 5 8192 > gpurun_out/quick_cfg5.log 2>&1; tail -4 gpurun_out/quick_cfg5.log
timeout 300 python tools/phase_profile.py 3 8192 > gpurun_out/phase_cfg3.log 2>&1; head -5 gpurun_out/phase_cfg3.log

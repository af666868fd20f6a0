#!/bin/bash
# concurrent recovery block: tests, sanitizer (serialised kernels: the heartbeat timeout must end the polling block), throughput
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_plain.log 2>&1; tail -4 gpurun_out/quick_plain.log
timeout 300 python tools/gpu_quick.py 5 8192 > gpurun_out/quick_cfg5.log 2>&1; tail -3 gpurun_out/quick_cfg5.log
OBCA_QUICK_INIT=0 timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_quick.py 3 600 > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches.csv python tools/gpu_quick.py 3 8192 > /dev/null 2>&1; grep -c obca gpurun_out/launches.csv; awk -F'","' 'NR>2{print $5, $15}' gpurun_out/launches.csv | tail -9 | cut -c1-150

#!/bin/bash
# round 2, first GPU session: parity tests with the restoration phase, cfg 3 throughput with / without it, bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_plain.log 2>&1; tail -4 gpurun_out/quick_plain.log
OBCA_QUICK_INIT=34 timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_noresto.log 2>&1; tail -4 gpurun_out/quick_noresto.log
OBCA_QUICK_INIT=0 timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_zero.log 2>&1; tail -4 gpurun_out/quick_zero.log
timeout 300 python tools/gpu_quick.py 5 8192 > gpurun_out/quick_cfg5.log 2>&1; tail -4 gpurun_out/quick_cfg5.log
timeout 300 python tools/gpu_quick.py 2 1024 > gpurun_out/quick_cfg2.log 2>&1; tail -4 gpurun_out/quick_cfg2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err

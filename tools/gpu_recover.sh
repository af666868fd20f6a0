#!/bin/bash
# GPU session: full GPU suite, cfg3 bench line, kernel-only numbers with and without the recovery rules, closed loop.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-900 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_plain.log 2>&1; tail -3 gpurun_out/quick_plain.log
OBCA_QUICK_INIT=786 timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_recover.log 2>&1; tail -3 gpurun_out/quick_recover.log
timeout 300 python tools/bench_closed_loop.py 4096 both > gpurun_out/closed_loop.json 2> gpurun_out/closed_loop.err; cat gpurun_out/closed_loop.json; tail -3 gpurun_out/closed_loop.err

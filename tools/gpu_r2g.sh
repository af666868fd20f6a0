#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_plain.log 2>&1; tail -3 gpurun_out/quick_plain.log
timeout 300 python tools/gpu_quick.py 5 8192 > gpurun_out/quick_cfg5.log 2>&1; tail -3 gpurun_out/quick_cfg5.log
OBCA_QUICK_INIT=0 timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
timeout 300 python tools/gpu_quick.py 2 1024 2>&1 | tail -2
timeout 300 python tools/bench_closed_loop.py 4096 both > gpurun_out/closed_loop.json 2> gpurun_out/closed_loop.err; cut -c1-420 gpurun_out/closed_loop.json; tail -2 gpurun_out/closed_loop.err

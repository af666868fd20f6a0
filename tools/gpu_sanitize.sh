#!/bin/bash
# final checks: compute-sanitizer memcheck / racecheck of the current kernels (sweep behind the ABI call, ordering pre-kernels), results unchanged
mkdir -p gpurun_out
timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_quick.py 3 2000 > gpurun_out/memcheck.log 2>&1; tail -2 gpurun_out/memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/gpu_quick.py 3 450 > gpurun_out/racecheck_3.log 2>&1; tail -3 gpurun_out/racecheck_3.log

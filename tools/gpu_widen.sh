#!/bin/bash
# GPU session for the rows either side of the solve: new GPU tests first, then the whole GPU suite, the closed-loop
# bench (host-orchestrated vs device-resident), the planner bench and one bench line.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loop.py -m gpu -q -x > gpurun_out/pytest_loop.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_loop.log
tail -30 gpurun_out/pytest_loop.log
timeout 300 python tools/bench_closed_loop.py 4096 both > gpurun_out/closed_loop.json 2> gpurun_out/closed_loop.err; cat gpurun_out/closed_loop.json; tail -3 gpurun_out/closed_loop.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python tools/bench_planner.py > gpurun_out/planner.json 2>&1; cat gpurun_out/planner.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err

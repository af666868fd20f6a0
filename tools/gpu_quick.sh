#!/bin/bash
# Short GPU session: parity tests, phase profile, one bench line, ncu full capture of the solver kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_profile.py 3 8192 > gpurun_out/phase_cfg3.log 2>&1; cat gpurun_out/phase_cfg3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obca_solve -s 1 -c 1 -o gpurun_out/prof python tools/gpu_quick.py 3 8192 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log

#!/bin/bash
# branch-free Riccati sweep behind an ABI call: A/B against the direct call, correctness (fixtures + full-size cfg 3 parity), phase profile
set -x
mkdir -p gpurun_out
bash tools/gpu_ab.sh libobca_b200.so libobca_ab_direct.so 2>&1 | tee gpurun_out/ab_sweep.log
timeout 300 python tools/gpu_quick.py 5 8192 2>&1 | tail -2
timeout 300 python tools/gpu_quick.py 2 1024 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fixtures or (full_size and 3) or legacy or benchmark" > gpurun_out/pytest_sweep.log 2>&1; tail -5 gpurun_out/pytest_sweep.log | cut -c1-300

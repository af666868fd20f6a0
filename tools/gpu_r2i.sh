#!/bin/bash
# new legacy-mode parity tests + the full ncu capture of the first-pass kernel (launch order per solve: polling block, first pass, device recovery)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "legacy or small_time" > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obca_solve -s 7 -c 1 -o gpurun_out/prof -f python tools/gpu_quick.py 3 8192 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log

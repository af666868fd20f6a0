#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_lpt_try.py 1955 2>&1 | tee gpurun_out/lpt_try_1.log | tail -12
OBCA_B200_RESERVE_SM=1 timeout 600 python tools/gpu_lpt_try.py 1955 2>&1 | tee gpurun_out/lpt_try_1r.log | head -4
timeout 600 python tools/gpu_lpt_try.py 977 2>&1 | tee gpurun_out/lpt_try_2.log | tail -12

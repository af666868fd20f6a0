#!/bin/bash
# longest-first work order against the caller's order, over the batches of eight ranks; then the GPU tests
mkdir -p gpurun_out
for e in 1 0; do echo "== OBCA_B200_FIFO=$e"; OBCA_B200_FIFO=$e timeout 600 python tools/gpu_seeds.py 2>&1 | tail -10; done | tee gpurun_out/order_ab.log
OBCA_B200_FIFO=1 timeout 300 python tools/gpu_quick.py 5 8192 2>&1 | tail -2
timeout 300 python tools/gpu_quick.py 5 8192 2>&1 | tail -2
OBCA_B200_FIFO=1 OBCA_QUICK_INIT=0 timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
OBCA_QUICK_INIT=0 timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300

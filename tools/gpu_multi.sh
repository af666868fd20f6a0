#!/bin/bash
# 2-GPU session (gpurun --gpus 2): GPU suite, bench at N = 1 and 2 (NCCL gather), reference arm under torchrun,
# scenario-sharded closed loop.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
bash tools/gpu_scale.sh
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; cut -c1-300 gpurun_out/bench_2gpu_ref.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_closed_loop.py 4096 > gpurun_out/closed_loop_2gpu.json 2> gpurun_out/closed_loop_2gpu.err; cat gpurun_out/closed_loop_2gpu.json; tail -3 gpurun_out/closed_loop_2gpu.err

#!/bin/bash
# 2-GPU check of the sharded bench path (NCCL gather) + the closed-loop bench on one GPU.  Run with gpurun --gpus 2.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_closed_loop.py 4096 > gpurun_out/closed_loop.json 2> gpurun_out/closed_loop.err; cat gpurun_out/closed_loop.json; tail -3 gpurun_out/closed_loop.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; cut -c1-600 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err; cut -c1-200 gpurun_out/bench_2gpu_ref.json

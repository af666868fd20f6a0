#!/bin/bash
# Scaling check of the sharded bench path: bench.py at N = 1, 2, 4, 8 GPUs of one box (gpurun --gpus 8).
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_scale.txt
NG=$(nvidia-smi -L | wc -l)
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; cut -c1-220 gpurun_out/scale_1.json
for n in 2 4 8; do
  [ "$n" -le "$NG" ] || continue
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  cut -c1-220 gpurun_out/scale_$n.json; tail -2 gpurun_out/scale_$n.err
done

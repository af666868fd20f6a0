#!/bin/bash
# sharded bench path on 2 GPUs: cfg 3 and cfg 5, plus the world-size-2 CPU arm
set -x
mkdir -p gpurun_out
for cfg in 3 5; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29530 + cfg)) bench.py --gpus 2 --cfg $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale2_cfg$cfg.json 2> gpurun_out/scale2_cfg$cfg.err
  cut -c1-1500 gpurun_out/scale2_cfg$cfg.json; tail -3 gpurun_out/scale2_cfg$cfg.err
done
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; cut -c1-300 gpurun_out/scale_1.json

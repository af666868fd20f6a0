#!/bin/bash
# Scaling of the sharded bench path: bench.py at N = 1, 2, 4, 8 GPUs of one box (gpurun --gpus 8): cfg 3 (headline) and
# cfg 5 (BASELINE configs[4]: 65,536 FIXED_SET instances over 8 GPUs)
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_scale.txt
NG=$(nvidia-smi -L | wc -l)
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; cut -c1-220 gpurun_out/scale_1.json
for n in 2 4 8; do
  [ "$n" -le "$NG" ] || continue
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  grep "^{" gpurun_out/scale_$n.json | cut -c1-220; tail -2 gpurun_out/scale_$n.err
done
timeout 400 python bench.py --gpus 1 --cfg 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale5_1.json 2> gpurun_out/scale5_1.err; cut -c1-220 gpurun_out/scale5_1.json
for n in 8; do
  [ "$n" -le "$NG" ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --cfg 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale5_$n.json 2> gpurun_out/scale5_$n.err
  grep "^{" gpurun_out/scale5_$n.json | cut -c1-220; tail -2 gpurun_out/scale5_$n.err
done

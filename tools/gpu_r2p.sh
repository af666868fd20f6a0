#!/bin/bash
mkdir -p gpurun_out
for e in 0 1; do
OBCA_BENCH_ROOT_COPY=$e timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/diag_2_$e.json 2> gpurun_out/diag_2_$e.err
grep "^{" gpurun_out/diag_2_$e.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('root copy $e', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_rank'], d['success_rate'])" || tail -5 gpurun_out/diag_2_$e.err
done

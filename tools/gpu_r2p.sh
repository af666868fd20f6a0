#!/bin/bash
mkdir -p gpurun_out
for g in serial overlapped; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --gather $g > gpurun_out/diag_4_$g.json 2> gpurun_out/diag_4_$g.err
grep "^{" gpurun_out/diag_4_$g.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$g', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_rank'])" || tail -5 gpurun_out/diag_4_$g.err
done

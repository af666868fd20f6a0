#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_ref.json
for m in nvml off smi; do
OBCA_BENCH_SAMPLER=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sampler_$m.json 2>> gpurun_out/bench.err
python -c "import json; d=json.loads(open('gpurun_out/bench_sampler_$m.json').read()); print('$m', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"
done

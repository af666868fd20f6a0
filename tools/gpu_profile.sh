#!/bin/bash
# round 2 profile session (no tests): bench lines (cfg 3 / 5 / 2, reference start), CPU arm, ncu launch list, fp64 instruction
# counts, one full capture of the first-pass kernel, phase profile.  Run under gpurun.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_ref.json
timeout 600 python bench.py --cfg 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2>> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_cfg5.json
timeout 600 python bench.py --cfg 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2>> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_cfg2.json
timeout 600 python bench.py --start reference --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_refstart.json 2>> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_refstart.json
timeout 300 python tools/phase_profile.py 3 8192 > gpurun_out/phase_cfg3.log 2>&1; cat gpurun_out/phase_cfg3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:obca_solve -s 6 -c 3 --csv --log-file gpurun_out/fp64_counts_cfg3.csv python tools/gpu_quick.py 3 8192 > gpurun_out/fp64_quick.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obca_solve -s 7 -c 1 -o gpurun_out/prof -f python tools/gpu_quick.py 3 8192 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out

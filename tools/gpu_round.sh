#!/bin/bash
# One GPU session: parity tests, phase profile, bench, ncu launch list + full capture.  Run under gpurun.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_profile.py 3 8192 > gpurun_out/phase_cfg3.log 2>&1; cat gpurun_out/phase_cfg3.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obca_solve -s 1 -c 1 -o gpurun_out/prof python tools/gpu_quick.py 3 8192 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 300 python tools/bench_closed_loop.py 4096 both > gpurun_out/closed_loop.json 2> gpurun_out/closed_loop.err; cat gpurun_out/closed_loop.json
timeout 600 python tools/sweep_recovery.py 4096 > gpurun_out/sweep_recovery.jsonl 2>> gpurun_out/closed_loop.err; cat gpurun_out/sweep_recovery.jsonl
timeout 120 python tools/bench_planner.py > gpurun_out/planner.json 2>&1; cat gpurun_out/planner.json
OBCA_QUICK_INIT=786 timeout 300 python tools/gpu_quick.py 3 8192 > gpurun_out/quick_recover.log 2>&1; tail -2 gpurun_out/quick_recover.log
ls -la gpurun_out

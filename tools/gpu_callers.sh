mkdir -p gpurun_out
timeout 300 python tools/bench_closed_loop.py 4096 both > gpurun_out/closed_loop.json 2> gpurun_out/closed_loop.err; cut -c1-330 gpurun_out/closed_loop.json
timeout 120 python tools/bench_planner.py > gpurun_out/planner.json 2>&1; cat gpurun_out/planner.json | cut -c1-300

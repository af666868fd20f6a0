#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "two_moving or lockstep or device_loop or longest_first" > gpurun_out/pytest_two.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_two.log; tail -6 gpurun_out/pytest_two.log | cut -c1-300

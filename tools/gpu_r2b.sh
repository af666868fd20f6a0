#!/bin/bash
# round 2: full-size parity report (tools/parity_report.py)
mkdir -p gpurun_out
timeout 1500 python tools/parity_report.py 256 > gpurun_out/parity_report.json 2> gpurun_out/parity_report.err; tail -14 gpurun_out/parity_report.err | cut -c1-900

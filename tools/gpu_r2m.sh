#!/bin/bash
mkdir -p gpurun_out
for e in 0 1; do echo "== OBCA_B200_RESERVE_SM=$e"; OBCA_B200_RESERVE_SM=$e timeout 600 python tools/gpu_seeds.py 2>&1 | tail -10; done | tee gpurun_out/seeds_ab.log

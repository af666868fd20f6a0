#!/bin/bash
# A/B of developer builds of the library: gpu_ab.sh <lib.so>... ; prints the cfg 3 kernel time of each
for L in "$@"; do
  echo "== $L"
  OBCA_B200_LIB=$PWD/vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200/csrc/$L timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -3
done

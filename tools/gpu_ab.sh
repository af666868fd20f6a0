#!/bin/bash
# A/B of developer builds of the library: gpu_ab.sh <lib.so | lib.so@ENV=VAL>... ; prints the cfg 3 kernel time of each
for spec in "$@"; do
  L=${spec%%@*}; E=""; [ "$spec" != "$L" ] && E=${spec#*@}
  echo "== $spec"
  env $E OBCA_B200_LIB=$PWD/vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200/csrc/$L timeout 300 python tools/gpu_quick.py 3 8192 2>&1 | tail -3
done

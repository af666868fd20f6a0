#!/bin/bash
# stall breakdown of the cfg 3 first-pass kernel at 1, 2 and 3 resident blocks per SM (and of a developer build)
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,sm__icc_request_hit_rate.pct,sm__inst_executed.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,sass__inst_executed_shared_loads,sass__inst_executed_shared_stores,lts__t_sector_hit_rate.pct,dram__bytes_write.sum,dram__bytes_read.sum
mkdir -p gpurun_out
for spec in "$@"; do
  L=${spec%%@*}; E=""; [ "$spec" != "$L" ] && E=${spec#*@}
  tag=$(echo "$spec" | tr '@=/' '___')
  env $E OBCA_B200_LIB=$PWD/vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200/csrc/$L timeout 600 ncu --metrics $M --clock-control none -k regex:obca_solve -s 4 -c 1 --csv --log-file gpurun_out/occ_$tag.csv python tools/gpu_quick.py 3 8192 > /dev/null 2>&1
  echo "== $spec"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/occ_$tag.csv")) if len(r)>14 and r[0]=="0"]
for r in rows: print("  %-95s %s %s"%(r[12],r[14],r[13]))
PY
done

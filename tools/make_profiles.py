#!/usr/bin/env python
"""Developer tool: turn the artefacts of tools/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profiles.py <round-tag>     e.g. r1
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
LIB = os.path.join(ROOT, "vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200", "csrc", "libobca_b200.so")

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'sass__inst_executed_global_loads', 'sass__inst_executed_global_stores',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__icc_request_hit_rate.pct', 'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
        'gcc__average_cache_request_hit_rate.pct',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio']


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    rep = os.path.join(OUT, "prof.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, u, v = r[0], r[1], r[2]
    d = {k: (u[i], v[i]) for i, k in enumerate(h)}
    sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    rd = float(d['dram__bytes_read.sum'][1]) * sc[d['dram__bytes_read.sum'][0]]
    wr = float(d['dram__bytes_write.sum'][1]) * sc[d['dram__bytes_write.sum'][0]]
    kname = [x for x in h if False] or None
    rt_path = os.path.join(PROF, "roofline_traffic.json")
    rt = json.load(open(rt_path)) if os.path.exists(rt_path) else {}
    rt.update({"kernel": "obca_solve_kernel<4,128,3,20,4,16>", "workload": "cfg3 B=8192", "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": 7160 * 8192,
               "source": "ncu --set full --clock-control none, profiles/%s_ncu_metrics.md" % tag})
    # fp64 instruction counts of the first-pass kernel (tools/gpu_profile.sh: gpurun_out/fp64_counts_cfg3.csv) -> FLOP per iteration
    fc = os.path.join(OUT, "fp64_counts_cfg3.csv")
    if os.path.exists(fc):
        cnt = {}
        for row in csv.reader(open(fc)):
            if len(row) > 14 and "20, 4, 16" in row[4] and "_op_d" in row[12]:
                cnt[row[12].split("_op_")[1].split("_")[0]] = int(row[14])
        it = b_iters = None
        try:
            b_iters = json.loads(open(os.path.join(OUT, "bench.json")).read())["roofline_fp64"]["iterations_per_launch"]
        except Exception:
            pass
        if len(cnt) == 3 and b_iters:
            rt["fp64_thread_instructions_per_launch"] = dict(cnt, source="ncu smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on.sum of the first-pass kernel (profiles/%s_fp64_counts_cfg3.csv)" % tag)
            rt["iterations_per_launch"] = b_iters
            rt["fp64_flop_per_iteration_cfg3"] = (2 * cnt["dfma"] + cnt["dmul"] + cnt["dadd"]) / b_iters
    json.dump(rt, open(rt_path, "w"), indent=1)
    tab = ["| metric | value | unit |", "|---|---|---|"] + ["| `%s` | %s | %s |" % (k, d[k][1], d[k][0]) for k in KEYS if k in d]
    src = os.path.join("/tmp", "src_%s.csv" % tag)
    open(src, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)
    func = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_func.py"), src, LIB, "Li20ELi4ELi16"], capture_output=True, text=True).stdout
    subprocess.run(["cp", os.path.join(OUT, "launches.csv"), os.path.join(PROF, "%s_launches.csv" % tag)])
    subprocess.run(["cp", os.path.join(OUT, "phase_cfg3.log"), os.path.join(PROF, "%s_phase_cycles_cfg3.txt" % tag)])
    open(os.path.join(PROF, "%s_ncu_metrics.md" % tag), "w").write("\n".join(tab) + "\n")
    open(os.path.join(PROF, "%s_ncu_by_function.txt" % tag), "w").write(func)
    b = json.loads(open(os.path.join(OUT, "bench.json")).read())
    rf = json.loads(open(os.path.join(OUT, "bench_ref.json")).read())
    json.dump({"b200": b, "reference": rf}, open(os.path.join(PROF, "%s_bench_lines.json" % tag), "w"), indent=1)
    print("\n".join(tab))
    print(func[:3000])


if __name__ == "__main__":
    main()

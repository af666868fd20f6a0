#!/usr/bin/env python
"""cfg 4: closed-loop receding horizon on the demo9 map, one moving 2x2 box, lidar range 8, N = 5,
B Monte-Carlo scenarios advanced in lock-step (<= 30 steps).  Reports solves/s over all steps (orchestration,
copies and the FREE / FIXED_SET / FIXED_NOTERM launches of every step included) for the host-orchestrated driver
(ClosedLoopBatch) and the device-resident loop (ClosedLoopDevice, obca_b200_loop_*), one JSON line each.

    python tools/bench_closed_loop.py [B] [host|device|both]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import closed_loop as cl, demo_setting as ds  # noqa: E402


def line(kind, B, o, dt):
    m = o["mode"]
    return json.dumps({"driver": kind, "workload": "cfg4: demo9 closed loop, %d scenarios, N=5, lidar 8" % B,
                       "solves": int(o["solves"]), "launches": int(o["launches"]), "seconds": dt,
                       "solves_per_s": o["solves"] / dt, "steps_mean": float(o["steps"].mean()),
                       "failed": int(o["failed"].sum()), "reached": int(o["reached"].sum()),
                       "free_solves": int((m == 0).sum()), "fixed_set_solves": int((m == 1).sum()),
                       "fixed_noterm_solves": int((m == 2).sum())})


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    which = sys.argv[2] if len(sys.argv) > 2 else "both"
    for kind, cls in (("host", cl.ClosedLoopBatch), ("device", cl.ClosedLoopDevice)):
        if which not in (kind, "both"):
            continue
        s = ds.problemSetting("demo9"); s.senseDis = 8
        drv = cls(s, cl.demo9_monte_carlo(B), N=5, Q_free=0.5, sense=8.0)
        drv.run()                      # warm-up: contexts, kernels
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            o = drv.run()
            dt = time.perf_counter() - t0
            if best is None or dt < best[1]:
                best = (o, dt)
        print(line(kind, B, *best), flush=True)
        drv.close()


if __name__ == "__main__":
    main()

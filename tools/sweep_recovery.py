"""Developer tool: failure count and throughput of the device-resident closed loop (cfg 4) for each recovery setting."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import closed_loop as cl, demo_setting as ds, _abi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for soft in (0, 1, 2, 3, 5):
    for retry in (0, 1):
        s = ds.problemSetting("demo9"); s.senseDis = 8
        init = _abi.INIT_WARM | _abi.init_soft(soft) | (_abi.INIT_RETRY if retry else 0)
        d = cl.ClosedLoopDevice(s, cl.demo9_monte_carlo(B), N=5, Q_free=0.5, sense=8.0, init=init)
        d.run(); t = time.perf_counter(); o = d.run(); dt = time.perf_counter() - t
        print(json.dumps({"soft": soft, "retry": retry, "failed": int(o["failed"].sum()), "solves": int(o["solves"]),
                          "seconds": round(dt, 4), "solves_per_s": round(o["solves"] / dt)}), flush=True)
        d.close()

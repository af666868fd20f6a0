"""Host planner throughput: native batched A* (obca_b200_astar_batch) against the Python planner, same queries.
Prints one JSON line.  No GPU involved."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import a_star as A, demo_setting as ds  # noqa: E402


def main(n=65536, n_py=256):
    out = {}
    for demo in ("demo1", "demo9"):
        grid = ds.problemSetting(demo).org_gridMap
        rng = np.random.default_rng(1)
        free = np.argwhere(grid == 0)
        i = rng.integers(len(free), size=(n, 2))
        st = np.stack([free[i[:, 0], 1], free[i[:, 0], 0], np.zeros(n)], 1).astype(float)
        go = np.stack([free[i[:, 1], 1], free[i[:, 1], 0], np.zeros(n)], 1).astype(float)
        A.plan_batch(grid, st[:64], go[:64])
        t0 = time.perf_counter(); ref, ln = A.plan_batch(grid, st, go); t1 = time.perf_counter()
        t2 = time.perf_counter()
        for k in range(n_py):
            A.plan_reference(grid, st[k], go[k])
        t3 = time.perf_counter()
        out[demo] = {"grid": list(grid.shape), "queries": n, "native_queries_per_s": n / (t1 - t0),
                     "python_queries_per_s": n_py / (t3 - t2), "mean_path_len": float(ln.mean())}
    out["host_threads"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

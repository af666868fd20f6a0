#!/usr/bin/env python
"""Kernel-only throughput of the generic kernels (sizes from the parameter block) over problem shapes other than the
BASELINE ones: ragged edge counts, the 8-edge variant, the longest horizon, the most obstacles.  One JSON line per shape.

    python tools/bench_shapes.py [batch]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import obca_testlib as common  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import obca as om, scenario as sc  # noqa: E402

SHAPES = [("quads_N20 (cfg 3 scene through make_polygon_batch)", [4, 4, 4, 4], 20, 0),
          ("ragged_3_to_8_edges_N12", [3, 4, 5, 6, 7, 8], 12, 0), ("longest_horizon_N31", [4, 3], 31, 0),
          ("twelve_quads_48_rows_N10", [4] * 12, 10, 0), ("octagons_moving_N8", [8, 8, 5], 8, 1),
          ("48_rows_ragged_N6", [8, 8, 8, 8, 8, 4, 4], 6, 0)]


def main():
    import torch
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    for name, sides, N, moving in SHAPES:
        b = sc.make_polygon_batch(sides, B, N, seed=1, moving=moving)
        prm, a = sc.batch_arrays(b)
        s = om.BatchSolver(prm, a["edge_ptr"], B)
        t = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.float64, device="cuda").contiguous()
        dv = {k: t(a[k]) for k in ("x0", "u0", "xref", "A", "b0", "db", "T_max", "term")}
        out = s.alloc_outputs(B, "cuda")
        ms = []
        for _ in range(3):
            s.solve(dv["x0"], dv["u0"], dv["xref"], dv["A"], dv["b0"], dv["db"], T_max=dv["T_max"], term=dv["term"], out=out)
            torch.cuda.synchronize()
            ms.append(s.last_kernel_ms())
        st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
        print(json.dumps({"shape": name, "N": N, "n_obs": len(sides), "rows": int(prm.rows), "batch": B,
                          "kernel_ms": round(min(ms), 3), "solves_per_s": round(B / min(ms) * 1e3),
                          "feasible": round(float((st >= 0).mean()), 4), "iters_mean": round(float(it.mean()), 1),
                          "iters_max": int(it.max())}), flush=True)
        s.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Full-size parity report: the CUDA path against the C oracle on the BASELINE configurations at their full batch sizes
(cfg 2: 1,024; cfg 3: 8,192; cfg 5: 8,192 per GPU), plus the size-independent certificate of every feasible result
(duals >= 0, |A^T lam| <= 1, signed distance >= dmin, dynamics residual).  Prints one JSON object.

    python tools/parity_report.py            (needs a GPU; the oracle runs on all host cores)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import obca_testlib as common  # noqa: E402
from oracle import c_oracle  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, obca as om, scenario as sc  # noqa: E402


def certificate(prm, a, g, b):
    ok = g["status"] >= 0
    N = prm.N
    ego = b.ego; L = ego[0] + ego[2]; W = ego[1] + ego[3]
    gv = np.array([L / 2, W / 2, L / 2, W / 2]); off = L / 2 - ego[2]
    ep = a["edge_ptr"]; A = a["A"]; b0 = a["b0"]; db = a["db"]
    worst = dict(dual_neg=0.0, dual_norm=0.0, dist=0.0, dyn=0.0)
    worst["dual_neg"] = float(max(0.0, -min(g["lam"][ok].min(), g["mu"][ok].min())))
    for k in range(N + 1):
        bk = b0 + (k * db if (db is not None and prm.mode != _abi.MODE_FREE) else 0.0)
        th = g["x"][ok, k, 2]; ct, st = np.cos(th), np.sin(th)
        tx = g["x"][ok, k, 0] + off * ct; ty = g["x"][ok, k, 1] + off * st
        for i in range(prm.n_obs):
            lam = g["lam"][ok, k, ep[i]:ep[i + 1]]; mu = g["mu"][ok, k, 4 * i:4 * i + 4]
            a1 = lam @ A[ep[i]:ep[i + 1], 0]; a2 = lam @ A[ep[i]:ep[i + 1], 1]
            worst["dual_norm"] = max(worst["dual_norm"], float((a1 * a1 + a2 * a2 - 1).max()))
            dist = -(mu @ gv) + tx * a1 + ty * a2 - lam @ bk[ep[i]:ep[i + 1]]
            worst["dist"] = max(worst["dist"], float((b.dmin - dist).max()))
    h = (g["T"][ok] if _abi.is_free(prm.mode) else np.ones(ok.sum())) * b.Ts
    x = g["x"][ok]; u = g["u"][ok]
    nx = x[:, :-1, 0] + h[:, None] * u[:, :, 0] * np.cos(x[:, :-1, 2]); ny = x[:, :-1, 1] + h[:, None] * u[:, :, 0] * np.sin(x[:, :-1, 2])
    nt = x[:, :-1, 2] + h[:, None] * u[:, :, 1]
    worst["dyn"] = float(max(np.abs(nx - x[:, 1:, 0]).max(), np.abs(ny - x[:, 1:, 1]).max(), np.abs(nt - x[:, 1:, 2]).max()))
    return worst


def main():
    rep = {}
    for cfg, B in ((2, 1024), (3, 8192), (5, 8192)):
        b = sc.make_batch(cfg, B)
        prm, a = common.batch_arrays(b)
        s = om.BatchSolver(prm, a["edge_ptr"], B)
        g = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"], term=a["term"])
        s.close()
        c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"],
                           term=a["term"], nthreads=os.cpu_count() or 1)
        both = (g["status"] >= 0) & (c["status"] >= 0)
        rel = lambda k: (np.abs(g[k][both] - c[k][both]).reshape(both.sum(), -1).max(1)
                         / np.maximum(1.0, np.abs(c[k][both]).reshape(both.sum(), -1).max(1)))
        ex = np.maximum(rel("x"), rel("u")); eT = rel("T"); eo = rel("obj")
        rep["cfg%d" % cfg] = {
            "batch": B, "gpu_feasible": float((g["status"] >= 0).mean()), "oracle_feasible": float((c["status"] >= 0).mean()),
            "feasibility_agreement": float(((g["status"] >= 0) == (c["status"] >= 0)).mean()),
            "primal_within_1e-4": float((np.maximum(ex, eT) <= 1e-4).mean()), "primal_within_1e-8": float((np.maximum(ex, eT) <= 1e-8).mean()),
            "objective_within_1e-6": float((eo <= 1e-6).mean()), "objective_within_1e-10": float((eo <= 1e-10).mean()),
            "primal_rel_median": float(np.median(ex)), "primal_rel_p99": float(np.quantile(ex, 0.99)),
            "iters_equal": float((g["iters"][both] == c["iters"][both]).mean()),
            "iters_mean_gpu": float(g["iters"].mean()), "iters_mean_oracle": float(c["iters"].mean()),
            "certificate_worst_violation": certificate(prm, a, g, b)}
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()

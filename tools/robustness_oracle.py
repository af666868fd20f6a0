#!/usr/bin/env python
"""Developer tool (CPU): how robust is the interior-point method + restoration phase, measured on populations with
the C oracle (the kernel runs the same algorithm).  Prints solved fractions, status histograms and iteration counts for
  * the reference-generated fixtures from the reference's start (zeros) and the warm start,
  * cfg 2 / 3 from the reference's start, cfg 5 from the warm start,
  * the closed loop of cfg 4 (failed scenarios) driven by the oracle.
Knobs of the oracle's developer hooks can be set on the command line (defaults = shipped values):
    python tools/robustness_oracle.py [max_resto=2] [cap=1] [keep=1] [budget=300] [B=256] [loops=128]
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import obca_testlib as common  # noqa: E402
from oracle import c_oracle  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, closed_loop as cl, demo_setting as ds, scenario as sc  # noqa: E402

kw = dict(a.split("=") for a in sys.argv[1:])
L = c_oracle.lib()
L.obca_oracle_set_resto.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
L.obca_oracle_set_cap.argtypes = [C.c_int, C.c_double, C.c_int]
L.obca_oracle_set_budget.argtypes = [C.c_int]
if "max_resto" in kw or "mask" in kw:
    L.obca_oracle_set_resto(0.1, 1e-8, 1000.0, int(kw.get("mask", 41)), int(kw.get("max_resto", 2)), 4)
if "cap" in kw or "keep" in kw:
    L.obca_oracle_set_cap(int(kw.get("cap", 1)), 1.0, int(kw.get("keep", 1)))
if "budget" in kw:
    L.obca_oracle_set_budget(int(kw["budget"]))
B = int(kw.get("B", 256)); loops = int(kw.get("loops", 128))
NT = os.cpu_count() or 1


def run(tag, prm, a, Ts=None):
    t = time.time()
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a.get("T_max"),
                       term=a.get("term"), nthreads=NT, Ts=Ts)
    st = c["status"]
    print("  %-34s solved %.4f iters mean %6.1f max %4d  %s  %.1fs" % (
        tag, (st >= 0).mean(), c["iters"].mean(), c["iters"].max(),
        {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}, time.time() - t), flush=True)
    return c


for name in ["demo9_N10_sg_free", "demo9_N5_fixed", "demo1_N6_astar_free", "demo9_N6_astar_free"]:
    for init, nm in [(_abi.INIT_ZERO, "ZERO"), (_abi.INIT_ZERO | _abi.INIT_RETRY, "ZERO|RETRY"), (_abi.INIT_WARM, "WARM")]:
        for mi, bp in [(10.0, 0.1), (0.1, 0.01)]:
            prm, a, d = common.fixture_arrays(name, init=init, mu_init=mi, bound_push=bp)
            c = run("%s %s mu0=%g" % (name, nm, mi), prm, a)
prm, a, Ts = common.recovery_cases(_abi.INIT_WARM)
run("recovery cases WARM", prm, a, Ts)
for cfg, init, nm in [(3, _abi.INIT_ZERO, "ZERO"), (3, _abi.INIT_ZERO | _abi.INIT_RETRY, "ZERO|RETRY"), (2, _abi.INIT_ZERO, "ZERO"),
                      (3, _abi.INIT_WARM, "WARM"), (5, _abi.INIT_WARM, "WARM"), (5, _abi.INIT_WARM | _abi.INIT_RETRY, "WARM|RETRY")]:
    b = sc.make_batch(cfg, B)
    prm, a = common.batch_arrays(b, init=init)
    run("cfg%d %s" % (cfg, nm), prm, a)
for init, nm in [(_abi.INIT_WARM, "WARM"), (_abi.INIT_WARM | _abi.INIT_RETRY, "WARM|RETRY")]:
    s = ds.problemSetting("demo9"); s.senseDis = 8
    t = time.time()
    drv = cl.ClosedLoopBatch(s, cl.demo9_monte_carlo(loops), N=5, Q_free=0.5, sense=8.0, init=init,
                             solver_factory=lambda prm, ep, cap: common.OracleSolver(prm, ep, cap, nthreads=NT))
    o = drv.run()
    print("  closed loop cfg4 %-10s %d scenarios: failed %d (%.2f %%), reached %d, solves %d, %.1fs" % (
        nm, loops, int(o["failed"].sum()), 100.0 * o["failed"].mean(), int(o["reached"].sum()), int(o["solves"]), time.time() - t), flush=True)

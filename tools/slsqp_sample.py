#!/usr/bin/env python
"""Developer tool (CPU): an independent solver on a sample of the BASELINE batches.  SciPy SLSQP (an SQP code that shares
nothing with the interior-point method) runs on the NumPy restatement of the NLP (oracle/obca_nlp.py) from the
solver-independent warm start point, and its optimum is compared with the C oracle's - the strongest stand-in available
here for the IPOPT run that cannot be made (CasADi is not installable).

    python tools/slsqp_sample.py [n2=32] [n3=32] [procs=8]      -> JSON on stdout (profiles/r2_slsqp_sample.json)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import obca_testlib as common  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import scenario as sc  # noqa: E402

kw = dict(a.split("=") for a in sys.argv[1:])
PROCS = int(kw.get("procs", os.cpu_count() or 1))
_S = {}


def one(job):
    cfg, i = job
    prm, a = _S[cfg]
    t = time.time()
    try:
        f, T, cmax, dmin = common.slsqp_polish(prm, a, None, i=i, start="warm", maxiter=300)
    except Exception as e:                      # (SLSQP raises on a singular working set)
        return cfg, i, None, None, None, None, time.time() - t
    return cfg, i, f, T, cmax, dmin, time.time() - t


def main():
    from multiprocessing import Pool
    from oracle import c_oracle
    out = {}
    jobs = []
    ref = {}
    for cfg, n in ((2, int(kw.get("n2", 32))), (3, int(kw.get("n3", 32)))):
        b = sc.make_batch(cfg, max(n, 64))
        prm, a = sc.batch_arrays(b)
        _S[cfg] = (prm, a)
        ref[cfg] = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"],
                                  term=a["term"], nthreads=os.cpu_count() or 1)
        jobs += [(cfg, i) for i in range(n) if ref[cfg]["status"][i] >= 0]
    with Pool(PROCS) as pool:
        res = pool.map(one, jobs, chunksize=1)
    for cfg in (2, 3):
        rows = [r for r in res if r[0] == cfg]
        conv = [r for r in rows if r[2] is not None and r[4] <= 1e-5 and r[5] >= -1e-5]
        c = ref[cfg]
        rel = np.array([abs(r[2] - c["obj"][r[1]]) / abs(c["obj"][r[1]]) for r in conv])
        relT = np.array([abs(r[3] - c["T"][r[1]]) / abs(c["T"][r[1]]) for r in conv])
        out["cfg%d" % cfg] = dict(instances=len(rows), slsqp_converged_feasible=len(conv),
                                  objective_rel_diff_max=float(rel.max()) if len(rel) else None,
                                  objective_rel_diff_median=float(np.median(rel)) if len(rel) else None,
                                  within_1e6=int((rel <= 1e-6).sum()), T_rel_diff_max=float(relT.max()) if len(relT) else None,
                                  slsqp_lower_by_more_than_1e6=int(sum(1 for r in conv if r[2] < c["obj"][r[1]] * (1 - 1e-6))),
                                  seconds_per_instance_median=float(np.median([r[6] for r in rows])))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

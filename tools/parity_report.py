#!/usr/bin/env python
"""Full-size parity report: the CUDA path (through the C-ABI) against the C oracle on the BASELINE configurations at
their full batch sizes (cfg 2: 1,024; cfg 3: 8,192; cfg 5: 8,192 per GPU), from the A* warm start (batch defaults) and
from the reference's own start (every variable 0, T = 1, obca.py:856, with IPOPT's mu_init 0.1 / bound_push 1e-2), with
the restoration phase off (the interior-point pass alone: same algorithm, so results must agree) and on.  Per run: the
numbers of tests/obca_testlib.parity_summary, including first-order optimality certificates of 256 GPU results
evaluated with oracle/obca_nlp.py.  Prints one JSON object.

    python tools/parity_report.py [kkt_sample=256]     (needs a GPU; the oracle runs on all host cores)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import obca_testlib as common  # noqa: E402
from oracle import c_oracle  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, obca as om, scenario as sc  # noqa: E402

STARTS = {"warm": (_abi.INIT_WARM, {}), "reference_start": (_abi.INIT_ZERO, dict(mu_init=0.1, bound_push=1e-2))}


def main():
    ks = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rep = {}
    for cfg, B in ((2, 1024), (3, 8192), (5, 8192)):
        b = sc.make_batch(cfg, B)
        for sname, (init, opts) in STARTS.items():
            for fname, flags in (("pass_only", _abi.INIT_NORESTO), ("with_restoration", 0)):
                prm, a = common.batch_arrays(b, init=init | flags, **opts)
                s = om.BatchSolver(prm, a["edge_ptr"], B)
                g = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"], term=a["term"])
                s.close()
                c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"],
                                   term=a["term"], nthreads=os.cpu_count() or 1)
                rep["cfg%d/%s/%s" % (cfg, sname, fname)] = common.parity_summary(prm, a, g, c, kkt_sample=ks if fname == "with_restoration" else 0)
                print("cfg%d/%s/%s" % (cfg, sname, fname), json.dumps(rep["cfg%d/%s/%s" % (cfg, sname, fname)]), file=sys.stderr, flush=True)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()

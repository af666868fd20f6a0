"""Generates tests/golden/recovery_cases.npz: inputs of closed-loop solves (cfg 4: demo9 map, moving box, N = 5, mode
FIXED_NOTERM) whose line search fails from the warm start and which the recovery rules solve.  Produced with this
repo's own oracle-backed closed-loop driver (no reference code involved); inputs only - the tests recompute results.

    python tests/golden/make_recovery_cases.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import obca_testlib as common  # noqa: E402
from oracle import c_oracle  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, closed_loop as cl, demo_setting as ds  # noqa: E402


def main(B=192, keep=6):
    s = ds.problemSetting("demo9"); s.senseDis = 8
    cases = []

    class Rec(common.OracleSolver):
        def solve_host(self, x0, u0, xref, A, b0, db=None, T_max=None, term=None, uref=None, out=None, Ts=None):
            r = super().solve_host(x0, u0, xref, A, b0, db, T_max=T_max, term=term, uref=uref, out=out, Ts=Ts)
            if self.params.mode == _abi.MODE_FIXED_NOTERM:
                for j in np.where(r["status"] < 0)[0]:
                    cases.append(dict(prm=bytes(self.params), ep=self.edge_ptr.copy(), x0=x0[j], u0=u0[j], xref=xref[j],
                                      A=A[j], b0=b0[j], db=db[j], Ts=Ts[j]))
            return r
    drv = cl.ClosedLoopBatch(s, cl.demo9_monte_carlo(B), N=5, Q_free=0.5, sense=8.0, init=_abi.INIT_WARM,
                             solver_factory=lambda prm, ep, cap: Rec(prm, ep, cap, nthreads=8))
    drv.run()
    good = []
    for c in cases:
        p = _abi.ObcaParams.from_buffer_copy(c["prm"]); p.init = _abi.INIT_WARM | _abi.RECOVER
        a = dict(x0=c["x0"][None], u0=c["u0"][None], xref=c["xref"][None], edge_ptr=c["ep"], A=c["A"][None], b0=c["b0"][None],
                 db=c["db"][None], T_max=None, term=None)
        r = c_oracle.solve(p, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], Ts=np.array([c["Ts"]]))
        e = common.emu_solve(p, a, Ts=np.array([c["Ts"]]))
        if r["status"][0] >= 0 and e["status"][0] >= 0:
            good.append(c)
        if len(good) == keep:
            break
    print("failing NOTERM solves: %d, recoverable kept: %d" % (len(cases), len(good)))
    st = lambda k: np.stack([c[k] for c in good])
    np.savez(os.path.join(HERE, "recovery_cases.npz"), params=np.frombuffer(good[0]["prm"], dtype=np.uint8), edge_ptr=good[0]["ep"],
             x0=st("x0"), u0=st("u0"), xref=st("xref"), A=st("A"), b0=st("b0"), db=st("db"), Ts=np.array([c["Ts"] for c in good]))


if __name__ == "__main__":
    main()

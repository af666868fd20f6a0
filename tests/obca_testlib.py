"""Shared helpers for the tests: turn a golden fixture / a synthetic Batch into ABI-level arrays."""
import glob
import os

import numpy as np

from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, scenario as sc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# feasible reference fixtures (demo1 N=5 is infeasible; the start/goal-only windows violate Tmax)
FEASIBLE = ["demo1_N6_astar_free", "demo2_N6_astar_free", "demo6_N6_astar_free", "demo9_N5_astar_free",
            "demo9_N6_astar_free", "demo1_N6_fixed", "demo9_N5_fixed"]
INFEASIBLE = ["demo1_N5_astar_free"]


def load_fixture(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    mode = _abi.MODE_FREE if name.endswith("free") else _abi.MODE_FIXED_SET
    return mode, {k: d[k] for k in d.files}


def fixture_arrays(name, init=_abi.INIT_WARM, **opts):
    mode, d = load_fixture(name)
    N, nObs = int(d["N"]), int(d["nObs"])
    ep, A, b0, db = _abi.pack_obstacles(mode, N, nObs, d["vObs"], d["AObs"], d["bObs"])
    prm = _abi.make_params(mode, N, nObs, int(ep[-1]), float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], d["xL"],
                           d["xU"], d["uL"], d["uU"], float(d["dmin"]), d["ego"], init=init, **opts)
    x0 = np.asarray(d["x0"], float).reshape(1, 3)
    u0 = np.asarray(d["u0"], float).reshape(1, 2)
    xref = np.ascontiguousarray(np.asarray(d["xref"], float).T).reshape(1, N + 1, 3)
    Tm = np.array([_abi.tmax_of(d["xref"][:, N], d["x0"], N, d["uU"][0], float(d["Ts"]))]) if _abi.is_free(mode) else None
    term = _abi.term_of(d["terminal_set"]).reshape(1, 3) if "terminal_set" in d else None
    return prm, dict(x0=x0, u0=u0, xref=xref, edge_ptr=ep, A=A, b0=b0, db=db, T_max=Tm, term=term), d


def batch_arrays(b, init=_abi.INIT_WARM, **opts):
    ep, A, b0, db = _abi.pack_obstacles(b.mode, b.N, b.nObs, b.vObs, b.AObs, b.bObs)
    prm = _abi.make_params(b.mode, b.N, b.nObs, int(ep[-1]), b.Ts, b.P, b.Q, b.R, b.xL, b.xU, b.uL, b.uU, b.dmin,
                           b.ego, init=init, **opts)
    xref = np.ascontiguousarray(b.xref.transpose(0, 2, 1))
    Tm = None
    if _abi.is_free(b.mode):
        Tm = ((b.xref[:, 0, b.N] - b.x0[:, 0]) + (b.xref[:, 1, b.N] - b.x0[:, 1])) / (b.N * b.uU[0] * b.Ts) + 1.0
    term = None
    if b.terminal_set is not None:
        term = np.stack([b.terminal_set[:, 0, 0], b.terminal_set[:, 1, 0], b.terminal_set[:, 1, 1]], axis=1)
    return prm, dict(x0=b.x0, u0=b.u0, xref=xref, edge_ptr=ep, A=A, b0=b0, db=db, T_max=Tm, term=term)


def rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())

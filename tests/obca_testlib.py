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


# ---- certificates from the NumPy restatement of the NLP (oracle/obca_nlp.py), independent of any solver state ------
def nlp_problem(prm, a, i=0, Ts=None):
    """oracle/obca_nlp Problem + Layout of instance i of ABI-level arrays"""
    from oracle import obca_nlp as nlp
    shared = a["A"].ndim == 2
    pick = lambda v: None if v is None else v[i]
    p = nlp.problem_from_abi(prm, a["edge_ptr"], a["x0"][i], a["u0"][i], a["xref"][i], a["A"] if shared else a["A"][i],
                             a["b0"] if shared else a["b0"][i], None if a["db"] is None else (a["db"] if shared else a["db"][i]),
                             Ts=None if Ts is None else float(Ts[i]), T_max=pick(a.get("T_max")), term=pick(a.get("term")),
                             uref=pick(a.get("uref")))
    return p, nlp.Layout(p)


def kkt_of(prm, a, out, i=0, Ts=None):
    """First-order optimality certificate (oracle/obca_nlp.kkt_certificate) of result i of a solver output dict"""
    from oracle import obca_nlp as nlp
    p, lay = nlp_problem(prm, a, i, Ts)
    X = nlp.pack(p, lay, out["x"][i], out["u"][i], out["T"][i], out["lam"][i], out["mu"][i])
    return nlp.kkt_certificate(p, lay, X)


def slsqp_polish(prm, a, out, i=0, perturb=1e-3, maxiter=300, start=None):
    """SciPy SLSQP (an independent SQP code) on the NumPy restatement, started from result i (trajectory perturbed) or
    from one of obca_nlp.start_point's named starts -> (objective, T, max |c|, min d)"""
    from scipy.optimize import minimize
    from oracle import obca_nlp as nlp
    p, lay = nlp_problem(prm, a, i)
    if start is None:
        X0 = nlp.pack(p, lay, out["x"][i], out["u"][i], out["T"][i], out["lam"][i], out["mu"][i])
        X0[:lay.ntraj] += perturb * np.random.default_rng(0).standard_normal(lay.ntraj)
    else:
        X0 = nlp.start_point(p, lay, start)
    den = lambda M: M.toarray() if hasattr(M, "toarray") else np.asarray(M)
    ev = lambda X, w: nlp.evaluate(p, lay, X, want=w)
    cons = [dict(type="eq", fun=lambda X: ev(X, ("c", "J"))["c"], jac=lambda X: den(ev(X, ("c", "J"))["J"])),
            dict(type="ineq", fun=lambda X: ev(X, ("d", "Jd"))["d"], jac=lambda X: den(ev(X, ("d", "Jd"))["Jd"]))]
    s = minimize(lambda X: ev(X, ("f", "g"))["f"], X0, jac=lambda X: ev(X, ("f", "g"))["g"], constraints=cons,
                 method="SLSQP", options=dict(maxiter=maxiter, ftol=1e-12))
    e = ev(s.x, ("c", "d"))
    return float(s.fun), (float(s.x[lay.T]) if p.free else 1.0), float(np.abs(e["c"]).max()), float(e["d"].min())


# ---- oracle-backed stand-ins (CPU tests only): same interfaces as BatchSolver / obca, C oracle underneath ----------
class OracleSolver:
    """BatchSolver look-alike (``solve_host`` / ``params`` / ``close``) that runs the C oracle.  Lets the host-side
    orchestration (closed loops, sharding) be tested without a GPU; never used by the product."""

    def __init__(self, params, edge_ptr, max_batch, device=-1, nthreads=4):
        self.params = params
        self.edge_ptr = np.ascontiguousarray(edge_ptr, dtype=np.int32)
        self.max_batch = max_batch
        self.nthreads = nthreads
        self.launches = 0

    def solve_host(self, x0, u0, xref, A, b0, db=None, T_max=None, term=None, uref=None, out=None, Ts=None):
        from oracle import c_oracle
        self.launches += 1
        r = c_oracle.solve(self.params, x0, u0, xref, self.edge_ptr, A, b0, db, T_max=T_max, term=term, uref=uref,
                           nthreads=self.nthreads, Ts=Ts)
        if out is not None:
            for k in r:
                out[k][...] = r[k]
            return out
        return r

    def close(self):
        pass


def oracle_obca():
    """The product's ``obca`` class with its GPU context swapped for the oracle (monkeypatched BatchSolver)."""
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import obca as om

    class _Obca(om.obca):
        def _one(self, *a, **k):
            saved = om.BatchSolver
            om.BatchSolver = OracleSolver
            try:
                return super()._one(*a, **k)
            finally:
                om.BatchSolver = saved
    return _Obca()


# ---- host emulation of the CUDA kernel's phase code (tools/emu) -------------------------------------------------
_emu = None


def emu_lib():
    """g++ build of tools/emu/obca_emu.cpp: csrc/obca_cta.cuh compiled for the host, block threads run serially."""
    global _emu
    if _emu is None:
        import ctypes as C
        import subprocess
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        src = os.path.join(root, "tools", "emu", "obca_emu.cpp")
        hdr = os.path.join(root, "vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200", "csrc", "obca_cta.cuh")
        lib = os.path.join(root, "tools", "emu", "libobca_emu.so")
        if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", lib, src])
        _emu = C.CDLL(lib)
        _emu.obca_emu_solve.restype = C.c_int
        _emu.obca_emu_solve.argtypes = [C.POINTER(_abi.ObcaParams)] + _abi.SOLVE_ARGTYPES_HOST + [C.c_int]
    return _emu


def emu_solve(params, a, Ts=None, uref=None):
    import ctypes as C
    L = emu_lib()
    f64 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)
    x0, u0, xref, T_max, term, A, b0, db, Ts, uref = map(f64, (a["x0"], a["u0"], a["xref"], a["T_max"], a["term"], a["A"],
                                                                a["b0"], a["db"], Ts, uref))
    B = x0.shape[0]; N, R, no = params.N, params.rows, params.n_obs
    out = dict(x=np.zeros((B, N + 1, 3)), u=np.zeros((B, N, 2)), lam=np.zeros((B, N + 1, R)), mu=np.zeros((B, N + 1, 4 * no)),
               T=np.zeros(B), obj=np.zeros(B), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
    ep = np.ascontiguousarray(a["edge_ptr"], dtype=np.int32)
    p = _abi.ptr
    rc = L.obca_emu_solve(C.byref(params), B, p(x0), p(u0), p(xref), p(uref), p(T_max), p(term), p(Ts), p(ep, C.c_int32),
                          p(A), p(b0), p(db), int(A.ndim == 2), p(out["x"]), p(out["u"]), p(out["lam"]), p(out["mu"]),
                          p(out["T"]), p(out["obj"]), p(out["status"], C.c_int32), p(out["iters"], C.c_int32), 1)
    if rc != 0:
        raise RuntimeError("obca_emu_solve rc=%d" % rc)
    return out


def recovery_cases(init):
    """tests/golden/recovery_cases.npz (closed-loop FIXED_NOTERM solves that fail from the warm start and that the
    recovery rules solve; made by tests/golden/make_recovery_cases.py) -> (params with ``init``, ABI arrays, Ts)"""
    d = np.load(os.path.join(GOLDEN, "recovery_cases.npz"))
    prm = _abi.ObcaParams.from_buffer_copy(d["params"].tobytes())
    prm.init = init
    a = dict(x0=d["x0"], u0=d["u0"], xref=d["xref"], edge_ptr=d["edge_ptr"], A=d["A"], b0=d["b0"], db=d["db"], T_max=None,
             term=None)
    return prm, a, d["Ts"]

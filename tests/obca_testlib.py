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
    return sc.batch_arrays(b, init=init, **opts)


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


# ---- parity protocol (GPU tests, tools/parity_report.py) ---------------------------------------------------------
PRIMAL_RTOL, OBJ_RTOL = 1e-4, 1e-6    # north_star: 1e-4 relative on primal variables, 1e-6 relative on the objective


def component_errors(g, c, sel):
    """Per-instance relative errors of result dict g against c over the instances `sel`, every component against its
    own scale: positions by the instance's largest coordinate, heading, speed, turn rate and T by max(1, max |.|) of
    that component (so 1e-4 means 1e-4 rad, 1e-4 m/s ... - not 1e-4 of a 39 m coordinate), objective relative."""
    n = int(sel.sum())
    mx = lambda v: np.abs(v).reshape(n, -1).max(1)
    e = lambda gv, cv: mx(gv - cv) / np.maximum(1.0, mx(cv))
    gx, cx, gu, cu = g["x"][sel], c["x"][sel], g["u"][sel], c["u"][sel]
    return dict(pos=e(gx[..., :2], cx[..., :2]), hdg=e(gx[..., 2], cx[..., 2]), v=e(gu[..., 0], cu[..., 0]),
                w=e(gu[..., 1], cu[..., 1]), T=e(g["T"][sel], c["T"][sel]),
                obj=np.abs(g["obj"][sel] - c["obj"][sel]) / np.maximum(1e-300, np.abs(c["obj"][sel])))


def parity_summary(prm, a, g, c, kkt_sample=0, Ts=None):
    """The numbers the parity tests assert on.  `within` = instances solved by both whose x, u, T agree to PRIMAL_RTOL
    (component-wise, see component_errors) and whose objective agrees to OBJ_RTOL.  For the others (`outliers`) and for
    a random sample of `kkt_sample` feasible GPU results the first-order optimality certificate of the NumPy
    restatement is evaluated (kkt_of): an outlier with a valid certificate on both sides is a different local
    solution of a non-convex problem, not an error."""
    gs, cs = g["status"] >= 0, c["status"] >= 0
    both = gs & cs
    e = component_errors(g, c, both)
    prim = np.maximum.reduce([e["pos"], e["hdg"], e["v"], e["w"], e["T"]])
    bad = (prim > PRIMAL_RTOL) | (e["obj"] > OBJ_RTOL)
    idx = np.flatnonzero(both)
    out = dict(batch=int(len(gs)), gpu_feasible=int(gs.sum()), oracle_feasible=int(cs.sum()), both=int(both.sum()),
               feasibility_agreement=float((gs == cs).mean()), within=int((~bad).sum()), outliers=int(bad.sum()),
               primal_max=float(prim[~bad].max()) if (~bad).any() else 0.0, obj_max=float(e["obj"][~bad].max()) if (~bad).any() else 0.0,
               primal_within_1e8=float((prim <= 1e-8).mean()) if len(prim) else 0.0,
               iters_equal=float((g["iters"][both] == c["iters"][both]).mean()) if both.any() else 0.0,
               status_gpu={int(k): int(v) for k, v in zip(*np.unique(g["status"], return_counts=True))},
               status_oracle={int(k): int(v) for k, v in zip(*np.unique(c["status"], return_counts=True))},
               iters_mean_gpu=float(g["iters"].mean()), iters_mean_oracle=float(c["iters"].mean()))
    ok = lambda k: k["c_max"] <= 1e-6 and k["d_min"] >= -1e-6 and k["stat"] <= 1e-5 and k["z_min"] >= -1e-6 and k["compl"] <= 1e-4
    out["outliers_certified"] = int(sum(ok(kkt_of(prm, a, g, i, Ts)) and ok(kkt_of(prm, a, c, i, Ts)) for i in idx[bad][:64]))
    out["outliers_checked"] = int(min(64, bad.sum()))
    out["outliers_gpu_better"] = int((g["obj"][idx[bad]] < c["obj"][idx[bad]]).sum())
    if kkt_sample:
        pick = np.random.default_rng(0).permutation(np.flatnonzero(gs))[:kkt_sample]
        ks = [kkt_of(prm, a, g, i, Ts) for i in pick]
        out["kkt"] = dict(sample=int(len(ks)), c_max=float(max(k["c_max"] for k in ks)), d_min=float(min(k["d_min"] for k in ks)),
                          stat_max=float(max(k["stat"] for k in ks)), stat_p99=float(np.quantile([k["stat"] for k in ks], 0.99)),
                          z_min=float(min(k["z_min"] for k in ks)), valid=int(sum(ok(k) for k in ks)))
        # results that ended on the rounding-noise floor (status 2: the solver's own dual error <= 1e-3 only): how
        # stationary are they when the multipliers are fitted instead of taken from the noisy iterate?
        fl = np.flatnonzero(g["status"] == 2)[:kkt_sample]
        kf = [kkt_of(prm, a, g, i, Ts) for i in fl]
        if kf:
            out["kkt_floor"] = dict(sample=int(len(kf)), stat_max=float(max(k["stat"] for k in kf)),
                                    c_max=float(max(k["c_max"] for k in kf)), valid=int(sum(ok(k) for k in kf)))
    return out


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

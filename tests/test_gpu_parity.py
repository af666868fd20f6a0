"""Parity of the CUDA path (through the C-ABI) against the C oracle.  -m gpu only.

Tolerances are the north_star's: 1e-4 relative on primal variables (x, u, T), 1e-6 relative on the objective.
lambda / mu are not unique where a distance constraint is inactive, so they are certificate-checked
(non-negative, dual norm <= 1, signed distance >= dmin) rather than value-compared.
"""
import os

import numpy as np
import pytest

import obca_testlib as common
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, obca as obca_mod, scenario as sc

pytestmark = pytest.mark.gpu

PRIMAL_RTOL = 1e-4
OBJ_RTOL = 1e-6


def _gpu(prm, a):
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], a["x0"].shape[0])
    try:
        return s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"], term=a["term"])
    finally:
        s.close()


def _cpu(prm, a, nthreads=8):
    from oracle import c_oracle
    return c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"],
                          term=a["term"], nthreads=nthreads)


def _compare(g, c, min_ok=0.9):
    """small / odd-shaped batches: every instance solved by both within tolerance except a bounded few that took
    another path (checked strictly, at full size and with certificates, by test_full_size_parity)"""
    both = (g["status"] >= 0) & (c["status"] >= 0)
    assert both.mean() >= min_ok, "only %.3f of the instances solved by both" % both.mean()
    assert ((g["status"] >= 0) == (c["status"] >= 0)).mean() >= 0.97
    e = common.component_errors(g, c, both)
    prim = np.maximum.reduce([e["pos"], e["hdg"], e["v"], e["w"], e["T"]])
    assert (prim <= PRIMAL_RTOL).mean() >= 0.98, np.sort(prim)[-5:]
    assert (e["obj"] <= OBJ_RTOL).mean() >= 0.98, np.sort(e["obj"])[-5:]


def _certificate(prm, a, g, dmin, ego):
    """every feas=True output satisfies the OBCA constraints to 1e-6"""
    ok = g["status"] >= 0
    N = prm.N
    L = ego[0] + ego[2]; W = ego[1] + ego[3]
    gv = np.array([L / 2, W / 2, L / 2, W / 2]); off = L / 2 - ego[2]
    ep = a["edge_ptr"]
    A = a["A"]; b0 = a["b0"]; db = a["db"]
    assert (g["lam"][ok] >= -1e-9).all() and (g["mu"][ok] >= -1e-9).all()
    for k in range(N + 1):
        bk = b0 + (k * db if (db is not None and prm.mode != _abi.MODE_FREE) else 0.0)
        th = g["x"][ok, k, 2]; ct, st = np.cos(th), np.sin(th)
        tx = g["x"][ok, k, 0] + off * ct; ty = g["x"][ok, k, 1] + off * st
        for i in range(prm.n_obs):
            lam = g["lam"][ok, k, ep[i]:ep[i + 1]]; mu = g["mu"][ok, k, 4 * i:4 * i + 4]
            a1 = lam @ A[ep[i]:ep[i + 1], 0]; a2 = lam @ A[ep[i]:ep[i + 1], 1]
            assert (a1 * a1 + a2 * a2 <= 1 + 1e-6).all()
            assert np.abs(mu[:, 0] - mu[:, 2] + ct * a1 + st * a2).max() <= 1e-6
            assert np.abs(mu[:, 1] - mu[:, 3] - st * a1 + ct * a2).max() <= 1e-6
            dist = -(mu @ gv) + tx * a1 + ty * a2 - lam @ bk[ep[i]:ep[i + 1]]
            assert (dist >= dmin - 1e-6).all()


@pytest.mark.parametrize("name", common.FEASIBLE)
@pytest.mark.parametrize("init", [_abi.INIT_ZERO, _abi.INIT_WARM])
def test_reference_fixtures(name, init):
    """demo1 / demo2 / demo6 / demo9 inputs produced by the reference's own host code (tests/golden)"""
    prm, a, d = common.fixture_arrays(name, init=init)
    g = _gpu(prm, a); c = _cpu(prm, a, 1)
    if (name, init) == ("demo9_N5_fixed", _abi.INIT_ZERO):
        # from the reference's all-zero start this one ends at a local minimiser of the constraint violation (the
        # car drives straight at the wall it has to pass; IPOPT: "Converged to a point of local infeasibility"); both
        # must report it.  With the retry rule it is solved: test_reference_benchmark_problem_on_the_gpu
        assert c["status"][0] in (-6, -4) and g["status"][0] in (-6, -4)
        return
    assert c["status"][0] >= 0 and g["status"][0] >= 0
    assert common.rel(g["x"], c["x"]) <= PRIMAL_RTOL and common.rel(g["u"], c["u"]) <= PRIMAL_RTOL
    assert abs(g["T"][0] - c["T"][0]) <= PRIMAL_RTOL * max(1, abs(c["T"][0]))
    assert abs(g["obj"][0] - c["obj"][0]) <= OBJ_RTOL * max(1, abs(c["obj"][0]))
    _certificate(prm, a, g, float(d["dmin"]), d["ego"])


def test_infeasible_reported():
    """demo1 with N=5 is infeasible at step 0: feas must be False, not a hang"""
    prm, a, d = common.fixture_arrays("demo1_N5_astar_free")
    g = _gpu(prm, a)
    assert g["status"][0] < 0


# (cfg, batch, start, allowed fraction of outliers among the instances solved by both, feasibility agreement)
# Free-time configurations: EVERY instance solved by both agrees (0 outliers at the full BASELINE sizes, from the warm
# start and from the reference's own start with IPOPT's mu_init / bound_push).  Fixed-time cfg 5 has several local
# solutions per instance (which side of a moving box to pass): paths that go through the restoration phase are
# chaotic, a few end at a different local solution - each of those must carry a first-order optimality certificate
# on both sides.
FULL = [(2, 1024, "warm", 0.0, 0.997), (2, 1024, "reference", 0.0, 0.99), (3, 8192, "warm", 0.0, 0.999),
        (3, 8192, "reference", 0.0, 0.998), (5, 4096, "warm", 0.03, 0.97)]


@pytest.mark.parametrize("cfg,B,start,max_outliers,min_agree", FULL)
def test_full_size_parity(cfg, B, start, max_outliers, min_agree):
    """BASELINE configurations at their full batch sizes, default flags (restoration phase on), through the C-ABI,
    against the C oracle: component-wise 1e-4 on x, u, T and 1e-6 on the objective for EVERY instance solved by both
    (see FULL for cfg 5), feasibility agreement, and first-order optimality certificates (oracle/obca_nlp.py: primal
    feasibility, stationarity with fitted multipliers >= 0, complementarity) of 256 GPU results"""
    b = sc.make_batch(cfg, B)
    opts = {} if start == "warm" else dict(mu_init=0.1, bound_push=1e-2)
    prm, a = common.batch_arrays(b, init=_abi.INIT_WARM if start == "warm" else _abi.INIT_ZERO, **opts)
    g = _gpu(prm, a); c = _cpu(prm, a, nthreads=os.cpu_count() or 8)
    r = common.parity_summary(prm, a, g, c, kkt_sample=256)
    print(r)
    assert r["feasibility_agreement"] >= min_agree, r
    assert r["both"] >= (0.95 if cfg != 5 else 0.6) * B, r
    assert r["outliers"] <= max_outliers * r["both"], r
    assert r["outliers_certified"] == r["outliers_checked"], r          # another local solution, not an error
    assert r["kkt"]["valid"] >= 0.99 * r["kkt"]["sample"] and r["kkt"]["c_max"] <= 1e-6 and r["kkt"]["d_min"] >= -1e-6, r
    _certificate(prm, a, g, b.dmin, b.ego)


@pytest.mark.parametrize("cfg,B", [(3, 2048), (5, 2048)])
def test_interior_point_pass_alone(cfg, B):
    """restoration phase off: kernel and oracle run the same interior-point pass, so every instance solved by both
    agrees - also in the fixed-time configuration"""
    b = sc.make_batch(cfg, B)
    prm, a = common.batch_arrays(b, init=_abi.INIT_WARM | _abi.INIT_NORESTO)
    g = _gpu(prm, a); c = _cpu(prm, a, nthreads=os.cpu_count() or 8)
    r = common.parity_summary(prm, a, g, c)
    print(r)
    assert r["feasibility_agreement"] >= 0.99 and r["outliers"] <= 0.002 * r["both"], r
    assert r["outliers_certified"] == r["outliers_checked"], r


def test_reference_benchmark_problem_on_the_gpu():
    """demo9, N = 10, start/goal-only reference (simulation.calc_time, the reference's only timed datapoint) through the
    drop-in class: solved, same answer as the oracle, first-order optimality certificate; and demo9_N5_fixed from the
    reference's start with the retry rule"""
    from oracle import c_oracle
    mode, d = common.load_fixture("demo9_N10_sg_free")
    s = obca_mod.obca()
    x, u, feas, Ts_opt = s.obca_mpc4(float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], int(d["N"]), d["x0"], d["xL"],
                                     d["xU"], d["uL"], d["uU"], d["xref"], int(d["nObs"]), d["vObs"], d["AObs"],
                                     d["bObs"], float(d["dmin"]), d["ego"], d["u0"])
    assert feas is True and s.status >= 0
    assert np.abs(x[:, -1] - d["xref"][:, -1]).max() <= 1e-6
    init = _abi.INIT_ZERO | _abi.INIT_RETRY | _abi.INIT_PATIENT
    prm, a, _ = common.fixture_arrays("demo9_N10_sg_free", init=init, mu_init=0.1, bound_push=1e-2)
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"])
    g = _gpu(prm, a)
    assert g["status"][0] >= 0 and c["status"][0] >= 0
    k = common.kkt_of(prm, a, g)
    assert k["c_max"] <= 1e-6 and k["d_min"] >= -1e-6 and k["stat"] <= 1e-5, k
    assert abs(g["obj"][0] - c["obj"][0]) <= 1e-6 * c["obj"][0] and abs(s.obj - c["obj"][0]) <= 1e-6 * c["obj"][0]
    prm, a, _ = common.fixture_arrays("demo9_N5_fixed", init=_abi.INIT_ZERO | _abi.INIT_RETRY)
    g = _gpu(prm, a)
    assert g["status"][0] == 0 and abs(g["obj"][0] - 0.06455441) < 1e-7


def test_reference_call_surface():
    """obca().obca_mpc4 / obca_mpc6 / obca_mpc8 with the reference's positional arguments and 4-tuple return"""
    mode, d = common.load_fixture("demo1_N6_astar_free")
    s = obca_mod.obca()
    x, u, feas, Ts_opt = s.obca_mpc4(float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], int(d["N"]), d["x0"], d["xL"],
                                     d["xU"], d["uL"], d["uU"], d["xref"], int(d["nObs"]), d["vObs"], d["AObs"],
                                     d["bObs"], float(d["dmin"]), d["ego"], d["u0"])
    assert x.shape == (3, 7) and u.shape == (2, 6) and feas is True
    assert abs(Ts_opt - 2.0378865) < 1e-5            # T ~ 20.3789
    assert abs(s.obj - 4334.19729465) < 1e-4
    mode, d = common.load_fixture("demo1_N6_fixed")
    args = (float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], int(d["N"]), d["x0"], d["xL"], d["xU"], d["uL"], d["uU"],
            d["xref"], int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"], float(d["dmin"]), d["ego"], d["u0"], None)
    x, u, feas, Ts_opt = s.obca_mpc6(*args, d["terminal_set"])
    assert feas is True and Ts_opt == float(d["Ts"]) and abs(s.obj - 0.02347928) < 1e-6
    x8, u8, feas8, _ = s.obca_mpc8(*args)
    assert feas8 is True and x8.shape == (3, 7)


def test_device_path_matches_host_path():
    import torch
    b = sc.make_batch(2, 64)
    prm, a = common.batch_arrays(b)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], 64)
    h = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"])
    t = lambda v: torch.as_tensor(v, dtype=torch.float64, device="cuda").contiguous()
    o = s.solve(t(a["x0"]), t(a["u0"]), t(a["xref"]), t(a["A"]), t(a["b0"]), None, T_max=t(a["T_max"]))
    torch.cuda.synchronize()
    assert s.launches == 6 and s.last_kernel_ms() > 0      # per solve: first pass, recovery block beside it, recovery over the device
    for k in ("x", "u", "T", "obj", "lam", "mu"):
        assert np.array_equal(o[k].cpu().numpy(), h[k]), k      # same kernel, same inputs: bit-identical
    assert np.array_equal(o["status"].cpu().numpy(), h["status"])
    s.close()


def test_per_instance_Ts_and_obstacle_rows():
    """ABI v2: per-instance sampling time + per-instance obstacle rows (what the lock-step closed loop feeds)"""
    from oracle import c_oracle
    prm, a, d = common.fixture_arrays("demo9_N5_fixed")
    B = 3
    rep = lambda v: None if v is None else np.repeat(v, B, axis=0)
    Ts = np.array([2.0, 1.5, 2.5])
    A = np.repeat(a["A"][None], B, 0); b0 = np.repeat(a["b0"][None], B, 0)
    db = np.repeat(a["db"][None], B, 0) * (Ts / 2.0)[:, None]
    c = c_oracle.solve(prm, rep(a["x0"]), rep(a["u0"]), rep(a["xref"]), a["edge_ptr"], A, b0, db, term=rep(a["term"]), Ts=Ts)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], B)
    g = s.solve_host(rep(a["x0"]), rep(a["u0"]), rep(a["xref"]), A, b0, db, term=rep(a["term"]), Ts=Ts)
    s.close()
    ok = c["status"] >= 0
    assert np.array_equal(ok, g["status"] >= 0) and ok[0] and ok[2]
    assert common.rel(g["x"][ok], c["x"][ok]) <= PRIMAL_RTOL and common.rel(g["u"][ok], c["u"][ok]) <= PRIMAL_RTOL
    assert (np.abs(g["obj"][ok] - c["obj"][ok]) <= OBJ_RTOL * np.maximum(1, np.abs(c["obj"][ok]))).all()


def test_obca2_uref_and_obca_free():
    """obca2 (free, time-stacked rows, uref in the cost: obca.py:421-424) and obca(...,'big') through the drop-in"""
    from oracle import c_oracle
    mode, d = common.load_fixture("demo1_N6_astar_free")
    N = int(d["N"])
    s = obca_mod.obca()
    uref = np.tile(np.array([[0.5], [0.0]]), (1, N))
    x, u, feas, Ts_opt = s.obca2(float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], N, d["x0"], d["u0"], d["xL"], d["xU"],
                                 d["uL"], d["uU"], d["xref"], uref, int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"],
                                 float(d["dmin"]), d["ego"], 0, '', [])
    assert feas and x.shape == (3, N + 1)
    ep, A, b0, db = _abi.pack_obstacles(_abi.MODE_FREE_STACKED, N, int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"])
    prm = _abi.make_params(_abi.MODE_FREE_STACKED, N, int(d["nObs"]), int(ep[-1]), float(d["Ts"]), d["P"], d["Q"],
                           [d["R1"], d["R2"]], d["xL"], d["xU"], d["uL"], d["uU"], float(d["dmin"]), d["ego"])
    Tm = np.array([_abi.tmax_of(d["xref"][:, N], d["x0"], N, d["uU"][0], float(d["Ts"]))])
    c = c_oracle.solve(prm, d["x0"].reshape(1, 3), d["u0"].reshape(1, 2), np.ascontiguousarray(d["xref"].T)[None], ep, A, b0, db,
                       T_max=Tm, uref=np.ascontiguousarray(uref.T)[None])
    assert abs(s.obj - c["obj"][0]) <= OBJ_RTOL * abs(c["obj"][0]) and np.abs(x.T - c["x"][0]).max() <= 1e-4
    x2, u2, feas2, T2 = s.obca(float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], N, d["x0"], d["u0"], d["xL"], d["xU"], d["uL"],
                               d["uU"], d["xref"], [], int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"], float(d["dmin"]), d["ego"], 0, 'big')
    assert feas2 and abs(s.obj - 4334.19729465) < 1e-3


def _fixture_in_mode(name, mode, has_term, uref=None, **opts):
    """arrays of a fixed-time fixture re-posed in another mode of the same call family"""
    _, d = common.load_fixture(name)
    N, nObs = int(d["N"]), int(d["nObs"])
    ep, A, b0, db = _abi.pack_obstacles(mode, N, nObs, d["vObs"], d["AObs"], d["bObs"])
    prm = _abi.make_params(mode, N, nObs, int(ep[-1]), float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], d["xL"], d["xU"],
                           d["uL"], d["uU"], float(d["dmin"]), d["ego"], has_term=has_term, **opts)
    a = dict(x0=np.asarray(d["x0"], float).reshape(1, 3), u0=np.asarray(d["u0"], float).reshape(1, 2),
             xref=np.ascontiguousarray(np.asarray(d["xref"], float).T).reshape(1, N + 1, 3), edge_ptr=ep, A=A, b0=b0, db=db,
             T_max=None, term=_abi.term_of(d["terminal_set"]).reshape(1, 3) if has_term else None, uref=uref)
    return prm, a, d


def _value_parity(prm, a, g, c):
    assert g["status"][0] >= 0 and c["status"][0] >= 0, (g["status"], c["status"])
    e = common.component_errors(g, c, np.array([True]))
    assert max(e["pos"][0], e["hdg"][0], e["v"][0], e["w"][0], e["T"][0]) <= PRIMAL_RTOL, e
    assert e["obj"][0] <= OBJ_RTOL, e
    k = common.kkt_of(prm, a, g)
    assert k["c_max"] <= 1e-6 and k["d_min"] >= -1e-6 and k["stat"] <= 1e-5 and k["z_min"] >= -1e-6, k


@pytest.mark.parametrize("name", ["demo1_N6_fixed", "demo9_N5_fixed"])
@pytest.mark.parametrize("init", [_abi.INIT_WARM, _abi.INIT_ZERO | _abi.INIT_RETRY])
def test_fixed_time_legacy_modes_value_parity(name, init):
    """MODE_FIXED_NOTERM (obca_mpc8, obca.py:1564-1758) and MODE_FIXED_OBCA2 (obca2 with fixtime = 1, obca.py:518-521: with
    a terminal set, and without one but with uref in the cost) on the reference-generated fixed-time fixtures: x, u within
    1e-4, objective within 1e-6 of the oracle, first-order optimality certificate of the GPU result; then the same
    problems through the drop-in methods"""
    from oracle import c_oracle
    _, d = common.load_fixture(name)
    N = int(d["N"])
    uref = np.tile(np.array([[0.4, 0.0]]), (1, N, 1))
    cases = [(_abi.MODE_FIXED_NOTERM, False, None), (_abi.MODE_FIXED_OBCA2, True, None), (_abi.MODE_FIXED_OBCA2, False, uref)]
    objs = []
    for mode, has_term, ur in cases:
        prm, a, _ = _fixture_in_mode(name, mode, has_term, ur, init=init)
        s = obca_mod.BatchSolver(prm, a["edge_ptr"], 1)
        g = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], term=a["term"], uref=a["uref"])
        s.close()
        c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], term=a["term"], uref=a["uref"])
        _value_parity(prm, a, g, c)
        objs.append(float(c["obj"][0]))
    if init != _abi.INIT_WARM:
        return
    s = obca_mod.obca()
    args = (float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], N, d["x0"])
    rest = (d["xL"], d["xU"], d["uL"], d["uU"], d["xref"])
    obs = (int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"], float(d["dmin"]), d["ego"])
    x, u, feas, Ts_opt = s.obca_mpc8(*args, *rest, *obs, d["u0"], None)
    assert feas is True and Ts_opt == float(d["Ts"]) and abs(s.obj - objs[0]) <= OBJ_RTOL * max(1.0, abs(objs[0]))
    x, u, feas, Ts_opt = s.obca2(*args, d["u0"], *rest, [], *obs, 1, '', d["terminal_set"])
    assert feas is True and Ts_opt == float(d["Ts"]) and abs(s.obj - objs[1]) <= OBJ_RTOL * max(1.0, abs(objs[1]))
    x, u, feas, Ts_opt = s.obca2(*args, d["u0"], *rest, uref[0].T, *obs, 1, '', [])
    assert feas is True and abs(s.obj - objs[2]) <= OBJ_RTOL * max(1.0, abs(objs[2]))


@pytest.mark.parametrize("step,T_expected", [(0.045, 0.8), (0.05, None)])
def test_obca_small_time_scale_bounds(step, T_expected):
    """obca.obca(..., fixtime=0, timeScale_size='small'): the time scale confined to [0.8, 1.2] (obca.py:239-240), on a
    straight reference the car can follow at that scale (from 0.5 m/s); with 4.5 cm between reference points the lower
    bound is active.  Against the oracle with the same box, through the drop-in method"""
    from oracle import c_oracle
    _, d = common.load_fixture("demo1_N6_astar_free")
    N, nObs = int(d["N"]), int(d["nObs"])
    xref = np.stack([3 + step * np.arange(N + 1), np.full(N + 1, 4.0), np.zeros(N + 1)], 0)        # (3, N+1), the reference's layout
    x0 = np.array([3.0, 4.0, 0.0]); u0 = np.array([0.5, 0.0])
    ep, A, b0, db = _abi.pack_obstacles(_abi.MODE_FREE_STACKED, N, nObs, d["vObs"], d["AObs"], d["bObs"])
    prm = _abi.make_params(_abi.MODE_FREE_STACKED, N, nObs, int(ep[-1]), float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], d["xL"],
                           d["xU"], d["uL"], d["uU"], float(d["dmin"]), d["ego"], T_min=0.8)
    a = dict(x0=x0.reshape(1, 3), u0=u0.reshape(1, 2), xref=np.ascontiguousarray(xref.T)[None], edge_ptr=ep, A=A, b0=b0, db=db,
             T_max=np.array([1.2]), term=None)
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], ep, A, b0, db, T_max=a["T_max"])
    g = _gpu(prm, a)
    _value_parity(prm, a, g, c)
    assert 0.8 - 1e-7 <= g["T"][0] <= 1.2 + 1e-7
    if T_expected is not None:
        assert abs(g["T"][0] - T_expected) <= 1e-6
    s = obca_mod.obca()
    x, u, feas, Ts_opt = s.obca(float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], N, x0, u0, d["xL"], d["xU"], d["uL"], d["uU"],
                                xref, [], nObs, d["vObs"], d["AObs"], d["bObs"], float(d["dmin"]), d["ego"], 0, 'small')
    assert feas is True and abs(s.obj - c["obj"][0]) <= OBJ_RTOL * abs(c["obj"][0])
    assert abs(Ts_opt - c["T"][0] * float(d["Ts"])) <= 1e-6 and np.abs(x.T - c["x"][0]).max() <= 1e-4


def test_sharded_single_rank_path():
    import torch
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import sharding
    b = sc.make_batch(2, 96)
    prm, a = common.batch_arrays(b)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], 96)
    out, packed = sharding.solve_sharded(s, a, 96, 0, 1, torch.device("cuda", 0))
    torch.cuda.synchronize()
    h = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"])
    for k in ("x", "u", "lam", "mu", "T", "obj", "status", "iters"):
        assert np.array_equal(out[k].cpu().numpy(), h[k]), k
    assert packed.nbytes == packed.buf.numel() * 8
    s.close()


def test_closed_loop_lockstep_batch_matches_oracle_driver():
    """cfg 4: demo9 map, Monte-Carlo moving box, lock-step closed loops on the GPU vs the same driver on the oracle"""
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import closed_loop as cl, demo_setting as ds
    B, steps = 48, 8
    dyn = cl.demo9_monte_carlo(B)
    dyn[:, 1] = np.linspace(12, 30, B)            # close enough to meet the car within a few steps
    dyn[:, 6] = 0
    mk = lambda: (lambda s: (setattr(s, "senseDis", 8), s)[1])(ds.problemSetting("demo9"))
    g = cl.ClosedLoopBatch(mk(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps)
    og = g.run(); g.close()
    c = cl.ClosedLoopBatch(mk(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps,
                           solver_factory=lambda prm, ep, cap: common.OracleSolver(prm, ep, cap, nthreads=8))
    oc = c.run()
    assert og["launches"] >= steps and og["solves"] >= B * 2
    # first step is identical for everybody (same start, no obstacle in range yet): tight agreement
    assert np.nanmax(np.abs(og["traj"][:, 1] - oc["traj"][:, 1])) <= 1e-4
    same = (og["steps"] == oc["steps"]) & (og["failed"] == oc["failed"])
    assert same.mean() >= 0.9
    n = np.minimum(og["steps"], oc["steps"])
    close = [np.abs(og["traj"][i, :n[i] + 1] - oc["traj"][i, :n[i] + 1]).max() <= 1e-3 for i in range(B) if same[i]]
    assert np.mean(close) >= 0.9
    assert (og["mode"] == _abi.MODE_FIXED_SET).any() or (og["mode"] == _abi.MODE_FIXED_NOTERM).any()


def test_closed_loop_lockstep_batch_with_two_moving_boxes():
    """Two moving boxes per scenario (the reference's demos 6, 7, 8, 11): NLPs with 5 + 1 and 5 + 2 obstacles in one
    lock-step run, on the GPU against the same driver on the oracle"""
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import closed_loop as cl, demo_setting as ds
    B, steps = 32, 8
    one = cl.demo9_monte_carlo(B)
    one[:, 1] = np.linspace(14, 26, B); one[:, 6] = 0; one[:, 5] = np.linspace(0.3, 0.7, B)
    two = one.copy(); two[:, 1] += 3.5                          # the second box follows 3.5 m behind the first
    two[B // 2:, 0] = 30.0                                      # ... or stays out of the lidar's range for good
    dyn = np.stack([one, two], 1)                               # (B, 2, 7)
    mk = lambda: (lambda s: (setattr(s, "senseDis", 8), s)[1])(ds.problemSetting("demo9"))
    g = cl.ClosedLoopBatch(mk(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps)
    og = g.run()
    counts = {key[1] for key in g._solvers if key[0] != _abi.MODE_FREE}
    g.close()
    assert counts == {1, 2}, counts
    c = cl.ClosedLoopBatch(mk(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps,
                           solver_factory=lambda prm, ep, cap: common.OracleSolver(prm, ep, cap, nthreads=8))
    oc = c.run()
    assert np.nanmax(np.abs(og["traj"][:, 1] - oc["traj"][:, 1])) <= 1e-4
    same = (og["steps"] == oc["steps"]) & (og["failed"] == oc["failed"])
    assert same.mean() >= 0.9
    n = np.minimum(og["steps"], oc["steps"])
    close = [np.abs(og["traj"][i, :n[i] + 1] - oc["traj"][i, :n[i] + 1]).max() <= 1e-3 for i in range(B) if same[i]]
    assert np.mean(close) >= 0.9
    with pytest.raises(ValueError):
        cl.ClosedLoopDevice(mk(), dyn, N=5)                     # the device-resident loop carries one box per scenario


@pytest.mark.parametrize("name,sides,N,moving", [("ragged_3_to_8_edges", [3, 4, 5, 6, 7, 8], 12, 0),
                                                 ("longest_horizon", [4, 3], 31, 0),
                                                 ("twelve_obstacles_48_rows", [4] * 12, 10, 0),
                                                 ("octagons_moving", [8, 8, 5], 8, 1),
                                                 ("48_rows_ragged", [8, 8, 8, 8, 8, 4, 4], 6, 0)])
def test_size_limits_and_ragged_polygons(name, sides, N, moving):
    """the generic kernels (sizes from the parameter block, 4- and 8-edge variants) at the compiled limits:
    N + 1 = 32 stages, 12 obstacles, 48 rows, 3..8 edges per obstacle"""
    b = sc.make_polygon_batch(sides, 160, N, seed=1, moving=moving)
    prm, a = common.batch_arrays(b)
    g = _gpu(prm, a); c = _cpu(prm, a)
    _compare(g, c, min_ok=0.6)
    _certificate(prm, a, g, b.dmin, b.ego)


def test_batch_of_one_and_empty_batch():
    b = sc.make_batch(2, 3)
    prm, a = common.batch_arrays(b)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], 8)
    full = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"])
    one = s.solve_host(a["x0"][1:2], a["u0"][1:2], a["xref"][1:2], a["A"], a["b0"], a["db"], T_max=a["T_max"][1:2])
    for k in ("x", "u", "T", "obj", "status", "iters"):
        assert np.array_equal(one[k][0], full[k][1]), k
    n0 = s.launches
    empty = s.solve_host(a["x0"][:0], a["u0"][:0], a["xref"][:0], a["A"], a["b0"], a["db"], T_max=a["T_max"][:0])
    assert empty["x"].shape == (0, prm.N + 1, 3) and s.launches == n0
    with pytest.raises(RuntimeError):
        s.solve_host(np.tile(a["x0"], (3, 1)), np.tile(a["u0"], (3, 1)), np.tile(a["xref"], (3, 1, 1)), a["A"], a["b0"], a["db"],
                     T_max=np.tile(a["T_max"], 3))                # 9 > max_batch
    s.close()


def test_large_host_batch_goes_through_in_chunks():
    """solve_host splits a large batch over several streams (copy-back overlaps the next chunk's solve): results are
    bit-identical to one launch over device tensors, for shared and per-instance obstacle rows, pinned or pageable"""
    import torch
    B = 2500
    b = sc.make_batch(3, B)
    prm, a = common.batch_arrays(b)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], B)
    t = lambda v: torch.as_tensor(v, dtype=torch.float64, device="cuda").contiguous()
    o = s.solve(t(a["x0"]), t(a["u0"]), t(a["xref"]), t(a["A"]), t(a["b0"]), None, T_max=t(a["T_max"]))
    torch.cuda.synchronize()
    n0 = s.launches
    h = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"],
                     out=s.alloc_host_outputs(B, pinned=True))
    assert s.launches - n0 == 12
    A_i = np.tile(a["A"][None], (B, 1, 1)); b_i = np.tile(a["b0"][None], (B, 1))
    h2 = s.solve_host(a["x0"], a["u0"], a["xref"], A_i, b_i, None, T_max=a["T_max"])
    for k in ("x", "u", "T", "obj", "lam", "mu", "status", "iters"):
        assert np.array_equal(o[k].cpu().numpy(), h[k]), k
        assert np.array_equal(h2[k], h[k]), k
    o2 = s.solve(t(a["x0"]), t(a["u0"]), t(a["xref"]), t(a["A"]), t(a["b0"]), None, T_max=t(a["T_max"]))   # device path after
    torch.cuda.synchronize()
    assert np.array_equal(o2["x"].cpu().numpy(), h["x"])
    s.close()
    # fixed-time mode: terminal sets, moving-obstacle increments and per-instance sampling times ride in the chunks too
    B = 2100
    b = sc.make_batch(5, B)
    prm, a = common.batch_arrays(b)
    Ts = np.linspace(1.5, 2.5, B)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], B)
    o = s.solve(t(a["x0"]), t(a["u0"]), t(a["xref"]), t(a["A"]), t(a["b0"]), t(a["db"]), term=t(a["term"]), Ts=t(Ts))
    torch.cuda.synchronize()
    n0 = s.launches
    h = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], term=a["term"], Ts=Ts)
    assert s.launches - n0 == 12
    for k in ("x", "u", "obj", "lam", "mu", "status", "iters"):
        assert np.array_equal(o[k].cpu().numpy(), h[k]), k
    s.close()


def test_restoration_phase_raises_the_success_rate():
    """The feasibility-restoration phase on the GPU: closed-loop solves whose line search fails from the warm start
    (status -4 with the phase switched off) are solved as in the oracle, every result carries a valid certificate;
    results of instances whose first pass succeeds do not change under the recovery flags"""
    prm0, a, Ts = common.recovery_cases(_abi.INIT_WARM | _abi.INIT_NORESTO)
    prm, _, _ = common.recovery_cases(_abi.INIT_WARM)

    def gpu(p):
        s = obca_mod.BatchSolver(p, a["edge_ptr"], a["x0"].shape[0])
        try:
            return s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], Ts=Ts)
        finally:
            s.close()
    from oracle import c_oracle
    g0 = gpu(prm0); g = gpu(prm)
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], Ts=Ts)
    assert (g0["status"] == -4).all()
    assert (g["status"] >= 0).all() and (c["status"] >= 0).all()
    assert (g["iters"] > g0["iters"]).all()
    same = np.abs(g["obj"] - c["obj"]) <= 1e-6 * np.maximum(1.0, np.abs(c["obj"]))
    assert same.mean() >= 0.8, (g["obj"], c["obj"])                         # non-convex: a restart may end elsewhere
    # certificate of every recovered result (per-instance obstacle rows)
    ego = np.array(list(prm.ego)); L = ego[0] + ego[2]; W = ego[1] + ego[3]
    gv = np.array([L / 2, W / 2, L / 2, W / 2]); off = L / 2 - ego[2]
    ep = a["edge_ptr"]
    for b_ in range(a["x0"].shape[0]):
        for k in range(prm.N + 1):
            bk = a["b0"][b_] + k * a["db"][b_]
            th = g["x"][b_, k, 2]; ct, st = np.cos(th), np.sin(th)
            tx = g["x"][b_, k, 0] + off * ct; ty = g["x"][b_, k, 1] + off * st
            for i in range(prm.n_obs):
                lam = g["lam"][b_, k, ep[i]:ep[i + 1]]; mu = g["mu"][b_, k, 4 * i:4 * i + 4]
                Ai = a["A"][b_, ep[i]:ep[i + 1]]
                a1, a2 = lam @ Ai[:, 0], lam @ Ai[:, 1]
                assert (lam >= -1e-9).all() and (mu >= -1e-9).all() and a1 * a1 + a2 * a2 <= 1 + 1e-6
                assert -(mu @ gv) + tx * a1 + ty * a2 - lam @ bk[ep[i]:ep[i + 1]] >= prm.dmin - 1e-6
    # a batch that needs no recovery is untouched by the flags
    b = sc.make_batch(3, 512)
    p0, a3 = common.batch_arrays(b); p1, _ = common.batch_arrays(b, soft_restarts=3, retry=True)
    r0 = _gpu(p0, a3); r1 = _gpu(p1, a3)
    ok = r0["status"] >= 0
    assert ok.mean() >= 0.99
    for k in ("x", "u", "T", "obj", "iters", "status"):
        assert np.array_equal(r0[k][ok], r1[k][ok]), k


def test_bulk_copy_staging_completes():
    """cp.async.bulk prefetch of the next instance's inputs / bulk store of the duals: no prefetch may time out (a timeout
    falls back to plain loads, so results alone would not show it), for shared and per-instance obstacle rows, odd row
    sizes (rows that are not 16-byte aligned take the coalesced-copy path) and a batch smaller than the grid"""
    for cfg, B in ((3, 2048), (2, 700), (3, 5)):
        b = sc.make_batch(cfg, B)
        prm, a = common.batch_arrays(b)
        s = obca_mod.BatchSolver(prm, a["edge_ptr"], B)
        g = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"])
        # per-instance obstacle rows: same numbers, every instance its own copy
        rep = lambda v: None if v is None else np.repeat(v[None], B, 0)
        g2 = s.solve_host(a["x0"], a["u0"], a["xref"], rep(a["A"]), rep(a["b0"]), rep(a["db"]), T_max=a["T_max"])
        assert s.bulk_timeouts == 0
        s.close()
        for k in ("x", "u", "lam", "mu", "T", "obj", "status", "iters"):
            assert np.array_equal(g[k], g2[k]), k
        c = _cpu(prm, a)
        _compare(g, c)
    b = sc.make_polygon_batch([3, 5, 7], 96, 9, seed=2)          # R = 15 (odd), N + 1 = 10: lam rows of odd instances unaligned
    prm, a = common.batch_arrays(b)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], 96)
    g = s.solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"], T_max=a["T_max"], term=a["term"])
    assert s.bulk_timeouts == 0
    s.close()
    _compare(g, _cpu(prm, a), min_ok=0.6)
    _certificate(prm, a, g, b.dmin, b.ego)

"""CPU tests of the oracle itself: the structured C restatement against the dense NumPy specification, KKT
certificates, an independent SciPy cross-check, and the regression values recorded in tests/golden.

PARITY UNPINNED: the reference ships no golden outputs and CasADi/IPOPT cannot be installed
here, so these tests pin the oracle against (i) an independent dense implementation of the same algorithm,
(ii) first-order optimality certificates evaluated with the NumPy restatement of the NLP, and (iii) SciPy's SLSQP
on the same NLP.  Inputs are the reference's own (tests/golden/make_reference_fixtures.py ran its host code)."""
import numpy as np
import pytest

import obca_testlib as common
from oracle import c_oracle, ipm_dense, obca_nlp as nlp
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi


def _problem(name):
    mode, d = common.load_fixture(name)
    ts = d["terminal_set"] if "terminal_set" in d else None
    return nlp.build_problem(mode, float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], int(d["N"]), d["x0"], d["xL"],
                             d["xU"], d["uL"], d["uU"], d["xref"], int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"],
                             float(d["dmin"]), d["ego"], d["u0"], terminal_set=ts)


def _c(name, init):
    prm, a, d = common.fixture_arrays(name, init=init)
    return prm, a, c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"],
                                  T_max=a["T_max"], term=a["term"])


@pytest.mark.parametrize("name,init", [("demo1_N6_astar_free", "zero"), ("demo1_N6_astar_free", "warm"),
                                       ("demo1_N6_fixed", "warm"), ("demo9_N5_astar_free", "warm")])
def test_c_oracle_matches_dense_spec(name, init):
    p = _problem(name)
    r = ipm_dense.solve(p, dict(init=init, max_iter=3000 if p.free else 1000, acceptable_tol=1e-6 if p.free else 1e-8))
    _, _, c = _c(name, _abi.INIT_ZERO if init == "zero" else _abi.INIT_WARM)
    assert r["status"] == 0 and c["status"][0] == 0
    assert np.abs(r["x"].T - c["x"][0]).max() <= 1e-9
    assert np.abs(r["u"].T - c["u"][0]).max() <= 1e-9
    assert abs(r["T"] - c["T"][0]) <= 1e-9
    assert abs(r["obj"] - c["obj"][0]) <= 1e-9 * max(1, abs(r["obj"]))


@pytest.mark.parametrize("name", common.FEASIBLE)
def test_kkt_certificate(name):
    """primal feasibility and stationarity (bounded least-squares multipliers, z >= 0, on the active set) at the
    oracle's solution, evaluated with the independent NumPy restatement of the NLP"""
    prm, a, c = _c(name, _abi.INIT_WARM)
    assert c["status"][0] >= 0
    k = common.kkt_of(prm, a, c)
    assert k["c_max"] <= 1e-7 and k["d_min"] >= -1e-7
    assert abs(k["f"] - c["obj"][0]) <= 1e-8 * max(1, abs(k["f"]))
    assert k["stat"] <= 1e-4 and k["z_min"] >= 0


@pytest.mark.parametrize("name", common.FEASIBLE)
def test_scipy_cross_check(name):
    """SLSQP (independent SQP code, SciPy) on the NumPy restatement, started near the oracle's solution, ends at the same
    objective to 1e-6 relative"""
    prm, a, c = _c(name, _abi.INIT_WARM)
    f, T, cmax, dmin = common.slsqp_polish(prm, a, c)
    assert cmax <= 1e-5 and dmin >= -1e-6
    assert abs(f - c["obj"][0]) <= 1e-6 * max(1e-2, abs(c["obj"][0])), (f, c["obj"][0])
    assert abs(T - c["T"][0]) <= 1e-4 * c["T"][0]


def test_scipy_from_its_own_start_demo9():
    """demo9 / N = 5: SLSQP from the interpolated start (poses on the reference window, everything else zero) - no
    information from the oracle at all - reaches the oracle's objective 7396.134.  SURVEY Appendix C quotes 7392.016 for
    an SLSQP run that ended with exit status 8 (not converged): 5.6e-4 lower at the same T, which at this problem's
    multipliers (2.7e3 on the terminal equality, test_kkt_certificate) is a constraint violation of 1.5e-3 - that
    figure is not an optimum of this NLP."""
    prm, a, c = _c("demo9_N5_astar_free", _abi.INIT_WARM)
    f, T, cmax, dmin = common.slsqp_polish(prm, a, None, start="xref", maxiter=200)
    assert cmax <= 1e-5 and dmin >= -1e-6
    assert abs(f - c["obj"][0]) <= 1e-6 * c["obj"][0] and abs(f - 7396.1345) < 0.01
    assert abs(T - 30.4518) < 1e-3


def test_reference_benchmark_problem_is_solved():
    """tests/golden/demo9_N10_sg_free.npz: the reference's own timed problem (simulation.calc_time, N_free = 10,
    start/goal-only reference, closed_loop.py:113-120).  From the reference's start (zeros) the first attempt ends in
    a failed restoration; the next start point of the sequence solves it.  The result carries a first-order optimality
    certificate and SLSQP started from it stays there (objective within 1e-6)."""
    init = _abi.INIT_ZERO | _abi.INIT_RETRY | _abi.INIT_PATIENT           # what the `obca` class uses for this reference
    prm, a, d = common.fixture_arrays("demo9_N10_sg_free", init=init, mu_init=0.1, bound_push=1e-2)
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"])
    assert c["status"][0] >= 0
    assert np.abs(c["x"][0, -1] - a["xref"][0, -1]).max() <= 1e-6 and 1e-4 <= c["T"][0] <= a["T_max"][0] + 1e-9
    k = common.kkt_of(prm, a, c)
    assert k["c_max"] <= 1e-7 and k["d_min"] >= -1e-7 and k["stat"] <= 1e-6 and k["z_min"] >= 0
    f, T, cmax, dmin = common.slsqp_polish(prm, a, c)
    assert cmax <= 1e-4 and dmin >= -1e-6
    assert abs(f - c["obj"][0]) <= 1e-6 * c["obj"][0], (f, c["obj"][0])


def test_fixed_time_from_the_reference_start():
    """demo9_N5_fixed from zeros: the first attempt converges to a local minimiser of the violation (the car drives
    straight at the wall it has to pass), status -6 like IPOPT's 'Converged to a point of local infeasibility'; with the
    retry rule the warm start follows and reaches the optimum 0.0645544"""
    prm, a, d = common.fixture_arrays("demo9_N5_fixed", init=_abi.INIT_ZERO)
    solve = lambda prm: c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], term=a["term"])
    assert solve(prm)["status"][0] in (-6, -4)
    prm, _, _ = common.fixture_arrays("demo9_N5_fixed", init=_abi.INIT_ZERO | _abi.INIT_RETRY)
    c = solve(prm)
    assert c["status"][0] == 0 and abs(c["obj"][0] - 0.06455441) < 1e-7
    k = common.kkt_of(prm, a, c)
    assert k["c_max"] <= 1e-7 and k["d_min"] >= -1e-7 and k["stat"] <= 1e-6


def test_regression_values():
    """objective / time scale recorded when the oracle was frozen (SURVEY Appendix C quotes the SLSQP values)"""
    want = {"demo1_N6_astar_free": (4334.19729465, 20.378865), "demo2_N6_astar_free": (4007.99042717, 19.444444),
            "demo9_N5_astar_free": (7396.13450351, 30.451764), "demo9_N6_astar_free": (7214.11928021, 27.474406),
            "demo1_N6_fixed": (0.02347928, 1.0), "demo9_N5_fixed": (0.06455441, 1.0)}
    for name, (obj, T) in want.items():
        for init in (_abi.INIT_ZERO, _abi.INIT_WARM):
            if name == "demo9_N5_fixed" and init == _abi.INIT_ZERO:
                continue
            _, _, c = _c(name, init)
            assert c["status"][0] >= 0, (name, init)
            assert abs(c["obj"][0] - obj) <= 1e-6 * max(1, abs(obj)), (name, init, c["obj"][0])
            assert abs(c["T"][0] - T) <= 1e-5 * max(1, T)


def test_infeasible_is_reported():
    """demo1 with N = 5 puts the terminal pose in collision: feas must be False"""
    _, _, c = _c("demo1_N5_astar_free", _abi.INIT_WARM)
    assert c["status"][0] < 0


def test_per_instance_Ts_equals_param_Ts():
    prm, a, d = common.fixture_arrays("demo1_N6_astar_free")
    c0 = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"])
    c1 = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], T_max=a["T_max"],
                        Ts=np.array([float(d["Ts"])]))
    assert np.array_equal(c0["x"], c1["x"]) and np.array_equal(c0["obj"], c1["obj"])


def test_objective_restatement():
    """objective_of (obca.py:859-895 written out) equals the f the solver minimised"""
    p = _problem("demo9_N6_astar_free")
    _, _, c = _c("demo9_N6_astar_free", _abi.INIT_WARM)
    f = nlp.objective_of(p, c["x"][0].T, c["u"][0].T, c["T"][0])
    assert abs(f - c["obj"][0]) <= 1e-9 * abs(f)


def test_watchdog_path_matches_dense_spec():
    """two cfg-2 instances whose solve goes through the watchdog (10 shortened steps -> full step on trust):
    the structured C oracle follows the dense specification step for step; without the watchdog they crawl"""
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import scenario as sc
    b = sc.make_batch(2, 400)
    prm, a = common.batch_arrays(b)
    L = c_oracle.lib()
    for i in (14, 383):
        sl = lambda v: None if v is None else v[i:i + 1]
        args = (prm, sl(a["x0"]), sl(a["u0"]), sl(a["xref"]), a["edge_ptr"], a["A"], a["b0"], a["db"])
        try:
            L.obca_oracle_set_watchdog(10, 0)
            c0 = c_oracle.solve(*args, T_max=sl(a["T_max"]))
        finally:
            L.obca_oracle_set_watchdog(10, 3)
        c1 = c_oracle.solve(*args, T_max=sl(a["T_max"]))
        assert c1["status"][0] == 0 and c1["iters"][0] < c0["iters"][0] - 15
        p = nlp.build_problem(b.mode, b.Ts, b.P, b.Q, b.R, b.N, b.x0[i], b.xL, b.xU, b.uL, b.uU, b.xref[i], b.nObs, b.vObs,
                              b.AObs, b.bObs, b.dmin, b.ego, b.u0[i])
        r = ipm_dense.solve(p, dict(init="warm"))
        assert r["status"] == 0 and r["iters"] == c1["iters"][0]
        assert np.abs(r["x"].T - c1["x"][0]).max() <= 1e-9 and abs(r["obj"] - c1["obj"][0]) <= 1e-9 * abs(r["obj"])
        assert abs(c0["obj"][0] - c1["obj"][0]) <= 1e-7 * abs(c1["obj"][0])      # same minimiser either way


def test_spec_recovery_sequence_matches_c_oracle():
    """the dense spec solver and the C oracle run the same recovery sequence (soft restarts, then other start points)"""
    prm, a, Ts = common.recovery_cases(_abi.INIT_WARM | _abi.RECOVER)
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], Ts=Ts)
    agree = 0
    for i in range(2):
        p = nlp.problem_from_abi(prm, a["edge_ptr"], a["x0"][i], a["u0"][i], a["xref"][i], a["A"][i], a["b0"][i], a["db"][i],
                                 Ts=float(Ts[i]))
        r0 = ipm_dense.solve(p, dict(init="warm", max_iter=1000, acceptable_tol=1e-8))
        r = ipm_dense.solve(p, dict(init="warm", max_iter=1000, acceptable_tol=1e-8, soft_restarts=3, retry=True))
        assert r["status"] >= 0 and c["status"][i] >= 0
        assert r["iters"] >= r0["iters"]
        agree += abs(r["obj"] - c["obj"][i]) <= 1e-6 * max(1.0, abs(c["obj"][i]))
    assert agree >= 1


def test_cfg5_outcomes_against_independent_geometry():
    """cfg 5 (fixed time, moving boxes): an audit that uses none of the solver's NLP code (tools/audit_cfg5.py: rectangle
    corners + separating-axis clearance at the 21 sample times).  No instance whose terminal set lies outside the map, or
    whose fixed start pose is already closer than dmin to an obstacle, may be reported feasible; every trajectory reported
    feasible keeps the clearance at every sample time and ends in its terminal set.  (The full audit, with the lattice
    search for the failures that are in neither class, is profiles/r2_cfg5_audit.json.)"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("audit_cfg5", os.path.join(os.path.dirname(common.GOLDEN), "..", "tools", "audit_cfg5.py"))
    import sys
    argv = sys.argv; sys.argv = ["audit_cfg5.py"]
    try:
        au = importlib.util.module_from_spec(spec); spec.loader.exec_module(au)
    finally:
        sys.argv = argv
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import scenario as sc
    B = 96
    b = au.prepare(sc.make_batch(5, B))
    prm, a = sc.batch_arrays(b, init=_abi.INIT_WARM)
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], term=a["term"], nthreads=os.cpu_count() or 1)
    ok = c["status"] >= 0
    clsA = b.x0[:, 0] + 5 > b.xU[0]
    clsB = ~au.clear(b.x0, au.polys_at(b, 0), b.ego, b.dmin)
    assert (clsA | clsB).sum() >= 5 and ok.sum() >= B // 2
    assert not (ok & (clsA | clsB)).any()
    for i in np.flatnonzero(ok):
        for k in range(b.N + 1):
            assert au.clear(c["x"][i, k][None], au.polys_at(b, k), b.ego, b.dmin - 1e-6)[0], (i, k)
        xN = c["x"][i, b.N]
        assert xN[0] >= b.x0[i, 0] + 5 - 1e-6 and 1 - 1e-6 <= xN[1] <= 9 + 1e-6


@pytest.mark.parametrize("name", common.FEASIBLE)
def test_scipy_from_the_warm_start_reaches_the_oracle_optimum(name):
    """SLSQP - an SQP code that shares nothing with the interior-point method - started from the warm start point of
    oracle/obca_nlp.start_point (poses on the reference window, T from the arc length, duals from the most separating
    face: no solver output in it) ends at the oracle's objective to 1e-6 on every reference-generated fixture, free-time and
    fixed-time.  With test_kkt_certificate this is what stands in for the IPOPT run that cannot be made here."""
    prm, a, c = _c(name, _abi.INIT_WARM)
    f, T, cmax, dmin = common.slsqp_polish(prm, a, None, start="warm", maxiter=300)
    assert cmax <= 1e-6 and dmin >= -1e-6
    assert abs(f - c["obj"][0]) <= 1e-6 * abs(c["obj"][0]), (f, c["obj"][0])
    if _abi.is_free(prm.mode):
        assert abs(T - c["T"][0]) <= 1e-5 * c["T"][0]

"""CPU tests of the receding-horizon layer with the oracle-backed solver stand-ins: the reference-shaped
``closedLoop`` object, and the lock-step ``ClosedLoopBatch`` driver against it (same scenario -> same closed loop)."""
import copy

import numpy as np

import obca_testlib as common
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, closed_loop as cl, demo_setting as ds


def _demo9_setting(dyn_row=None):
    s = ds.problemSetting("demo9")
    if dyn_row is not None:
        s.add_dynamic_obstacle([list(dyn_row)])
    s.senseDis = 8
    return s


def _single(dyn_row, steps):
    s = _demo9_setting(dyn_row)
    solver = common.oracle_obca()
    # the solver settings of the lock-step drivers (the `obca` class defaults to IPOPT's mu_init / bound_push)
    solver.mu_init, solver.bound_push, solver.recover = 10.0, 0.1, _abi.RECOVER
    c = cl.closedLoop(s, solver=solver)
    c.N_free = c.N_fix = 5
    c.Q_free = 0.5 * np.eye(3); c.P_free = c.Q_free            # simulation.py:68-70
    c.max_steps = steps
    logs = c.closed_loop_mpc4()
    return c, logs


def test_open_loop_two_stage_demo9_reference_benchmark():
    """simulation.calc_time (the reference's only timed datapoint, 3.69 s for N = 10; simulation.py:225-231): demo9,
    N_free = 10, free-time solve from the start/goal-only reference (closed_loop.py:113-120, 535-544), then
    simulation.run's second stage - the fixed-time solve on its result.  Real start, real goal."""
    s = ds.problemSetting("demo9")
    c = cl.closedLoop(s, solver=common.oracle_obca())
    c.N_free = 10
    c.mpc_openLoop_freeTime()
    assert c.xref.shape == (3, 11) and np.array_equal(c.xref[:, 1], c.xref[:, 10])      # start, then the goal ten times
    assert c.feas and c.xOpt.shape == (3, 11) and c.uOpt.shape == (2, 10)
    assert np.allclose(c.xOpt[:, -1], s.goalPose, atol=1e-6)      # terminal equality (obca.py:951)
    T = c.Ts_opt / c.Ts
    Tmax = ((s.goalPose[0] - s.startPose[0]) + (s.goalPose[1] - s.startPose[1])) / (10 * 0.6 * c.Ts) + 1     # obca.py:961-962
    assert 1e-4 <= T <= Tmax + 1e-6 and abs(T - 133.959) < 0.01
    # every pose of the plan keeps the car outside every obstacle by d_min (certificate through the duals)
    lam, mu = c.obca_solver.lam, c.obca_solver.mu
    assert lam.min() >= -1e-9 and mu.min() >= -1e-9
    c.N_fix = 10
    c.mpc_openLoop_fixTime()
    assert c.feas and c.Ts == c.Ts_opt


def test_open_loop_demo1():
    """demo1 with its real goal is infeasible as an open-loop problem: T_max (obca.py:961-962) allows 0.6 m more than the
    straight line and the box [10,15]x[1,5] lies on it - the solver must say so (IPOPT: restoration converges to a
    point of local infeasibility; the reference prints 'MPC1 -- Failed', simulation.py:48).  With a goal before the box
    the two stages run and reach the optimum of the A*-window fixture."""
    c = cl.closedLoop(ds.problemSetting("demo1"), solver=common.oracle_obca())
    c.N_free = 10
    c.mpc_openLoop_freeTime()
    assert not c.feas and c.obca_solver.status in (-4, -6)
    c = cl.closedLoop(ds.problemSetting("demo1"), solver=common.oracle_obca())
    c.x0 = [3, 4, 0]
    c.xF = [10, 6, 0]
    c.mpc_openLoop_freeTime()
    assert c.feas and c.xOpt.shape == (3, 7) and c.uOpt.shape == (2, 6)
    assert np.allclose(c.xOpt[:, -1], [10, 6, 0], atol=1e-6)
    T1 = c.Ts_opt
    assert abs(T1 - 2.0378865) < 1e-5            # same optimum as from the A* window (tests/golden demo1_N6)
    c.terminal_set = None
    c.setting.terminal_set = np.array([[8, 39], [1, 9]])
    c.mpc_openLoop_fixTime()
    assert c.feas and c.Ts == c.Ts_opt
    assert c.xOpt[0, -1] >= 8 - 1e-6


def test_closed_loop_mpc4_demo9_first_steps():
    row = [8, 50, -np.pi / 2, 2, 2, 0.5, 8, 10, -np.pi / 2, 0, 100]
    c, (x_open, x_opt, u_opt, T_opt) = _single(row, 4)
    assert len(x_opt) == 5 and len(u_opt) == 4 and len(T_opt) == 4
    xs = np.asarray(x_opt)
    assert (np.diff(xs[:, 0]) > 0).all()              # drives along the A* path away from the start
    for xo in x_open:
        assert xo.shape == (6, 3)
    # the plant is the prediction (Q11): state k+1 == second column of the k-th open-loop plan
    for k in range(4):
        assert np.array_equal(np.asarray(x_opt[k + 1]), x_open[k][1])


def test_lockstep_batch_equals_single_loops():
    """ClosedLoopBatch on 3 scenarios == three independent closedLoop.closed_loop_mpc4 runs"""
    dyn = np.array([[8, 50, -np.pi / 2, 2, 2, 0.5, 0], [8, 22, -np.pi / 2, 2, 2, 0.8, 0], [8, 16, -np.pi / 2, 2, 2, 0.3, 2]], float)
    steps = 6
    drv = cl.ClosedLoopBatch(_demo9_setting(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps,
                             solver_factory=lambda prm, ep, cap: common.OracleSolver(prm, ep, cap))
    out = drv.run()
    assert out["solves"] >= 3 * steps - 3 and out["launches"] >= steps
    modes = set(np.unique(out["mode"])) - {-1}
    assert _abi.MODE_FREE in modes
    for i in range(3):
        row = [dyn[i, 0], dyn[i, 1], dyn[i, 2], dyn[i, 3], dyn[i, 4], dyn[i, 5], 8, 10, -np.pi / 2, int(dyn[i, 6]), 100]
        c, (x_open, x_opt, u_opt, T_opt) = _single(row, steps)
        n = len(x_opt) - 1
        assert n == out["steps"][i], (i, n, out["steps"][i])
        got = out["traj"][i, :n + 1]
        assert np.allclose(got, np.asarray(x_opt, float), rtol=0, atol=1e-7), (i, np.abs(got - np.asarray(x_opt, float)).max())
    # at least one scenario must have met the obstacle (fixed-time phase exercised)
    assert (_abi.MODE_FIXED_SET in modes) or (_abi.MODE_FIXED_NOTERM in modes) or out["failed"].any()


def test_lockstep_batch_with_two_moving_boxes_equals_single_loops():
    """Two moving boxes per scenario (the reference's demos 6, 7, 8, 11 carry two): the fixed-time solve is built from the
    boxes the lidar sees at that step - both, or only the second one, in which case polygon and velocity of THAT box must
    be the ones in the NLP.  Lock-step batch == independent closedLoop.closed_loop_mpc4 runs with the same two rows."""
    hp = -np.pi / 2
    dyn = np.array([[[8, 20, hp, 2, 2, 0.6, 0], [8, 23.5, hp, 2, 2, 0.6, 0]],       # both come into range
                    [[30, 50, hp, 2, 2, 0.4, 0], [8, 18, hp, 2, 2, 0.3, 2]],        # the first never does
                    [[8, 50, hp, 2, 2, 0.5, 0], [8, 16, hp, 2, 2, 0.3, 0]]], float)
    steps = 7
    drv = cl.ClosedLoopBatch(_demo9_setting(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps,
                             solver_factory=lambda prm, ep, cap: common.OracleSolver(prm, ep, cap))
    # the lock-step driver runs the terminal-set solve without the recovery rules (it has its own fallback); the
    # single-scenario object has one setting for all its solves - compare like with like
    drv._init_of = lambda mode: drv.init
    out = drv.run()
    counts = {key[1] for key in drv._solvers if key[0] != _abi.MODE_FREE}
    assert 1 in counts and 2 in counts, counts                  # NLPs with one and with two moving boxes were built
    for i in range(3):
        s = ds.problemSetting("demo9")
        s.add_dynamic_obstacle([[r[0], r[1], r[2], r[3], r[4], r[5], 8, 10, hp, int(r[6]), 100] for r in dyn[i]])
        s.senseDis = 8
        solver = common.oracle_obca()
        solver.mu_init, solver.bound_push, solver.recover = 10.0, 0.1, _abi.RECOVER
        c = cl.closedLoop(s, solver=solver)
        c.N_free = c.N_fix = 5
        c.Q_free = 0.5 * np.eye(3); c.P_free = c.Q_free
        c.max_steps = steps
        x_open, x_opt, u_opt, T_opt = c.closed_loop_mpc4()
        n = len(x_opt) - 1
        assert n == out["steps"][i], (i, n, out["steps"][i])
        got = out["traj"][i, :n + 1]
        assert np.allclose(got, np.asarray(x_opt, float), rtol=0, atol=1e-7), (i, np.abs(got - np.asarray(x_opt, float)).max())


def test_sensor_keeps_polygon_and_velocity_paired():
    """two live moving obstacles (demo11), only the second within lidar range: the fixed-time NLP must be built from the
    second obstacle's polygon AND the second obstacle's velocity row"""
    s = ds.problemSetting("demo11")
    s.senseDis = 8
    c = cl.closedLoop(s, solver=common.oracle_obca())
    rows = [list(r) for r in s.dyn_obs_info]
    assert len(rows) == 2
    c.x0 = np.array([rows[1][0] - 5.0, rows[1][1], 0.0])          # next to the second obstacle, far from the first
    rows[0][0] += 60.0                                            # (make sure the first one is out of range)
    s.dyn_obs_info = rows
    c.update_obstacle(max(int(r[9]) for r in rows), c.Ts)
    c.sensor()
    assert c.fixtime == 1 and s.dyn_nObs == 1 and len(s.dyn_lObs) == 1
    want = ds.mo.get_obstacle(*s.dyn_obs_info[0][:5])
    assert np.allclose(np.asarray(s.dyn_lObs[0], float)[:, :2], np.asarray(want, float)[:, :2])
    assert abs(s.dyn_obs_info[0][0] - (c.x0[0] + 5.0)) < 3.0       # it is the obstacle next to the car
    nObs, vObs, lObs, info = s.combine_obstacle(1)
    assert nObs == s.static_nObs + 1 and info[-1] == list(s.dyn_obs_info[0])

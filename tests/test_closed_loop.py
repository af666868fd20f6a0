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
    c = cl.closedLoop(s, solver=common.oracle_obca())
    c.N_free = c.N_fix = 5
    c.Q_free = 0.5 * np.eye(3); c.P_free = c.Q_free            # simulation.py:68-70
    c.max_steps = steps
    logs = c.closed_loop_mpc4()
    return c, logs


def test_open_loop_two_stage_demo1():
    """simulation.run: free-time solve from start/goal reference, then the fixed-time solve on its result"""
    c = cl.closedLoop(ds.problemSetting("demo1"), solver=common.oracle_obca())
    c.x0 = [3, 4, 0]
    c.xF = [10, 6, 0]
    c.mpc_openLoop_freeTime()
    assert c.feas and c.xOpt.shape == (3, 7) and c.uOpt.shape == (2, 6)
    assert np.allclose(c.xOpt[:, -1], [10, 6, 0], atol=1e-6)      # terminal equality (obca.py:951)
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

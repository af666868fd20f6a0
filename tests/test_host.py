"""CPU tests of the host-side input layer against golden vectors produced by RUNNING the reference's own
host code (tests/golden/make_reference_fixtures.py)."""
import hashlib

import numpy as np
import pytest

import obca_testlib as common
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import (_abi, a_star as astar_mod, closed_loop as cl,
                                                                                demo_setting as ds, model_obstacle as mo,
                                                                                scenario as sc)

GOLD = common.GOLDEN


def test_astar_known_answer():
    """a_star.demo_data route (a_star.py:202-232): 27 cells, goal -> first cell after start"""
    d = np.load(GOLD + "/astar_demo_data.npz")
    grid, start, goal = astar_mod.a_star().demo_data()
    assert np.array_equal(grid, d["grid"]) and start == tuple(d["start"]) and goal == tuple(d["goal"])
    route = astar_mod.a_star(grid, start, goal).solve(grid, start, goal)
    assert np.array_equal(np.array(route), d["route"])
    assert len(route) == 27 and route[0] == (0, 19) and route[-1] == (0, 1)


@pytest.mark.parametrize("demo,N,sha", [("demo1", 6, "7ed9838f0d160643"), ("demo2", 6, "cd5bd15addb12837"),
                                        ("demo6", 6, "c5c2afb8839caf5b"), ("demo9", 5, "137463c114a72b9f")])
def test_demo_inputs_match_reference(demo, N, sha):
    """problemSetting + closedLoop input builders reproduce the reference's A* path, stacked H-rep and window"""
    d = np.load(GOLD + "/%s_N%d_astar_free.npz" % (demo, N))
    c = cl.closedLoop(ds.problemSetting(demo), solver=object())
    path = c.update_path(0, c.x0, c.xF, 0, "A_star")
    assert np.array_equal(path, d["path"])
    assert hashlib.sha256(np.ascontiguousarray(path).tobytes()).hexdigest().startswith(sha)
    c.update_obstacle_constraint(N, c.Ts, 0)
    assert np.array_equal(c.AObs, d["AObs"]) and np.array_equal(c.bObs, d["bObs"])
    assert list(c.vObs) == list(d["vObs"]) and c.nObs == int(d["nObs"])
    assert np.array_equal(c.update_reference_trajectory(N, path, c.x0), d["xref"])


@pytest.mark.parametrize("demo,N", [("demo1", 6), ("demo9", 5)])
def test_time_stacked_dynamic_obstacle(demo, N):
    d = np.load(GOLD + "/%s_N%d_fixed.npz" % (demo, N))
    c = cl.closedLoop(ds.problemSetting(demo), solver=object())
    c.update_obstacle_constraint(N, 2.0, 1)
    assert np.array_equal(c.AObs, d["AObs"]) and np.array_equal(c.bObs, d["bObs"])
    ep, A, b0, db = _abi.pack_obstacles(_abi.MODE_FIXED_SET, N, c.nObs, c.vObs, c.AObs, c.bObs)
    R = int(ep[-1])
    bk = np.asarray(c.bObs).reshape(N + 1, R)
    assert np.allclose(b0[None] + np.arange(N + 1)[:, None] * db[None], bk, atol=1e-12)
    # mpc4 reads the first block only (obca.py:969)
    ep4, A4, b4, db4 = _abi.pack_obstacles(_abi.MODE_FREE, N, c.nObs, c.vObs, c.AObs, c.bObs)
    assert db4 is None and np.array_equal(b4, bk[0])


def test_hrep_rotated_rectangle():
    d = np.load(GOLD + "/hrep_rotated_rect.npz")
    v = mo.get_obstacle(*d["args"])
    assert np.allclose(np.asarray(v), d["verts"], atol=0, rtol=0)
    A, b = mo.obstacleModel().obstacle_H_Represent(1, [5], [v])
    assert np.array_equal(A, d["A"]) and np.array_equal(b, d["b"])
    # unnormalised rows, norms 1.1547, 2, 1.1547, 2
    assert np.allclose(np.sqrt((A ** 2).sum(1)), [1.1547005, 2, 1.1547005, 2], atol=1e-6)


def test_demo1_hrep_rows():
    """demo1, one time block static + dynamic, R = 10"""
    c = cl.closedLoop(ds.problemSetting("demo1"), solver=object())
    c.update_obstacle_constraint(6, 0.1, 1)
    got = np.hstack([c.AObs[:10], c.bObs[:10]])
    want = np.array([[0, -1, -9], [-1, 0, -10], [0, 1, 5], [1, 0, 15], [0, -1, -1], [0, 1, 1], [-1, 0, -21],
                     [0, 1, 1.5], [1, 0, 24], [0, -1, 1.5]], float)
    assert np.array_equal(got, want)
    assert c.AObs.shape == (70, 2)
    c.update_obstacle_constraint(6, 0.1, 0)
    assert c.AObs.shape == (42, 2)


def test_reference_window_and_tmax():
    d = np.load(GOLD + "/demo1_N6_astar_free.npz")
    want = np.array([[4, 5, 6, 7, 8, 9, 10], [4, 4, 4, 4, 4, 5, 6], [0, 0, 0, 0, np.pi / 4, np.pi / 4, 0]])
    assert np.allclose(d["xref"], want)
    assert abs(_abi.tmax_of(d["xref"][:, 6], d["x0"], 6, d["uU"][0], float(d["Ts"])) - 26.0) < 1e-12
    # window clamps to the last path point
    ref = d["path"]
    w = sc.update_reference_trajectory(6, ref, np.array([37.0, 4.0, 0.0]))
    assert np.array_equal(w[:, -1], ref[:, -1]) and np.array_equal(w[:, -2], ref[:, -1])


def test_update_path_variants():
    c = cl.closedLoop(ds.problemSetting("demo1"), solver=object())
    r = c.update_path(6, [3, 4, 0], [38, 4, 0], 0, "startGoal_only")
    assert r.shape == (3, 7) and np.array_equal(r[:, 0], [3, 4, 0]) and np.all(r[0, 1:] == 38)
    r = c.update_path(5, [0, 0, 0], [5, 5, 0], 0, "startGoal_smooth")
    assert np.allclose(r[2], np.pi / 4) and np.allclose(r[0], np.arange(6))
    # allAviable=1 with N_fix == N_free keeps the points, recomputes the yaw and inherits the step (Q6)
    c.N_free = c.N_fix = 6
    c.xref = np.array([[0, 1, 2, 3, 3, 3, 3.0], [0, 0, 0, 1, 2, 3, 4.0], np.zeros(7)])
    c.Ts_opt = 2.5
    r = c.update_path(0, 0, 0, 1, "")
    assert r.shape == (3, 7) and np.array_equal(r[:2], c.xref[:2])
    assert np.allclose(r[2], [0, 0, np.pi / 4, np.pi / 2, np.pi / 2, np.pi / 2, np.pi / 2])
    assert c.Ts == 2.5 and c.N_fix == 6


def test_sensor_and_obstacle_update():
    c = cl.closedLoop(ds.problemSetting("demo1"), solver=object())
    c.update_obstacle(0, 0.1)
    c.sensor()
    assert c.fixtime == 0                      # box at (22.5, 0) is ~17.8 m from the car front at (4.7, 4)
    c.x0 = [14.0, 4.0, 0.0]
    c.update_obstacle(1, 2.0)                  # moved 2.0 * 0.2 up
    assert abs(c.dyn_orignal_info[0][1] - 0.4) < 1e-12
    c.sensor()
    assert c.fixtime == 1 and c.setting.dyn_nObs == 1


def test_box_hrep_batch_matches_scalar_builder():
    rng = np.random.default_rng(0)
    n = 64
    cx = rng.uniform(5, 30, n); cy = rng.uniform(5, 50, n)
    th = np.where(rng.random(n) < 0.5, -np.pi / 2, rng.uniform(-3, 3, n))
    A, b, V = cl.box_hrep_batch(cx, cy, th, np.full(n, 2.0), np.full(n, 3.0))
    for i in range(n):
        v = mo.get_obstacle(cx[i], cy[i], th[i], 2.0, 3.0)
        A1, b1 = mo.obstacleModel().obstacle_H_Represent(1, [5], [v])
        assert np.array_equal(A1, A[i]) and np.array_equal(b1.ravel(), b[i])


def test_synthetic_batches_are_deterministic_and_feasible_shaped():
    b1 = sc.make_batch(2, 32); b2 = sc.make_batch(2, 32)
    assert np.array_equal(b1.x0, b2.x0) and np.array_equal(b1.AObs, b2.AObs)
    assert b1.AObs.shape == (11 * 8, 2) and b1.xref.shape == (32, 3, 11) and b1.nObs == 2
    b3 = sc.make_batch(3, 16)
    assert b3.AObs.shape == (21 * 16, 2) and b3.N == 20
    b5 = sc.make_batch(5, 16)
    assert b5.nObs == 6 and b5.AObs.shape == (21 * 24, 2) and b5.terminal_set.shape == (16, 2, 2)
    ep, A, b0, db = _abi.pack_obstacles(b5.mode, b5.N, b5.nObs, b5.vObs, b5.AObs, b5.bObs)
    assert db is not None and np.count_nonzero(db) > 0 and np.count_nonzero(db[:16]) == 0
    # other pose stream, same scene
    b4 = sc.make_batch(3, 16, pose_seed=5)
    assert np.array_equal(b4.AObs, b3.AObs) and not np.array_equal(b4.x0, b3.x0)


def test_pack_obstacles_rejects_non_translation():
    d = np.load(GOLD + "/demo1_N6_fixed.npz")
    bad = d["bObs"].copy(); bad[25] += 0.3
    with pytest.raises(ValueError):
        _abi.pack_obstacles(_abi.MODE_FIXED_SET, 6, int(d["nObs"]), d["vObs"], d["AObs"], bad)
    with pytest.raises(ValueError):
        _abi.make_params(_abi.MODE_FREE, 40, 1, 4, 0.1, np.eye(3), np.eye(3), [np.eye(2), np.eye(2)], [0, 0], [1, 1],
                         [-1, -1], [1, 1], 0.05, [1, 1, 1, 1])

"""Native batched planner (obca_b200_astar_batch / obca_b200_reference_windows) against the
reference-generated golden paths and the Python planner, bit for bit.  Host code only: runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import (_lib, a_star as astar_mod,
                                                                               demo_setting as ds, scenario as sc)

GOLD = __file__.rsplit("/", 1)[0] + "/golden"


def test_known_answer_grid():
    """a_star.demo_data (a_star.py:202-232): the reference's own 27-cell route"""
    d = np.load(GOLD + "/astar_demo_data.npz")
    s, g = d["start"], d["goal"]
    ref, n = astar_mod.plan_batch(d["grid"], [[s[1], s[0], 0]], [[g[1], g[0], 0]])
    assert n[0] == 27
    route = d["route"][::-1]                                    # golden is goal -> start order, (row, col)
    assert np.array_equal(ref[0, :27, 0], route[:, 1]) and np.array_equal(ref[0, :27, 1], route[:, 0])


@pytest.mark.parametrize("demo,N", [("demo1", 6), ("demo2", 6), ("demo6", 6), ("demo9", 5)])
def test_demo_paths_match_reference(demo, N):
    d = np.load(GOLD + "/%s_N%d_astar_free.npz" % (demo, N))
    S = ds.problemSetting(demo)
    ref, n = astar_mod.plan_batch(S.org_gridMap, [S.startPose], [S.goalPose])
    assert n[0] == d["path"].shape[1]
    assert np.array_equal(ref[0, :n[0]].T, d["path"])           # cells and yaw, bit for bit
    w = astar_mod.reference_windows(ref, n, [S.startPose], N)
    assert np.array_equal(w[0].T, d["xref"])


@pytest.mark.parametrize("demo", ["demo1", "demo9"])
def test_random_queries_match_python_planner(demo):
    rng = np.random.default_rng(7)
    grid = ds.problemSetting(demo).org_gridMap
    free = np.argwhere(grid == 0)
    q = 200
    i = rng.integers(len(free), size=(q, 2))
    st = np.stack([free[i[:, 0], 1], free[i[:, 0], 0], np.zeros(q)], 1).astype(float)
    go = np.stack([free[i[:, 1], 1], free[i[:, 1], 0], np.zeros(q)], 1).astype(float)
    st[:5] = go[:5]                                             # start == goal: no path
    for threads in (1, 0):
        ref, n = astar_mod.plan_batch(grid, st, go, threads=threads)
        for k in range(q):
            py = astar_mod.plan_reference(grid, st[k], go[k])
            if py is None:
                assert n[k] == 0
            else:
                assert n[k] == py.shape[1] and np.array_equal(ref[k, :n[k]].T, py)
    ok = np.where(n > 0)[0]
    x0 = st[ok] + rng.uniform(-0.7, 0.7, (len(ok), 3))
    w = astar_mod.reference_windows(ref, n, x0, 9, ok)
    for j, k in enumerate(ok):
        assert np.array_equal(w[j].T, sc.update_reference_trajectory(9, ref[k, :n[k]].T, x0[j]))


def test_several_grids_and_blocked_goal():
    g = np.zeros((3, 6, 8))
    g[1, :, 4] = 1                                              # wall: no route
    g[2, 1:, 4] = 1                                             # wall with a gap at row 0
    st = [[0, 3, 0]] * 3; go = [[7, 3, 0]] * 3
    ref, n = astar_mod.plan_batch(g, st, go, grid_index=[0, 1, 2])
    assert n[0] == 7 and n[1] == 0 and n[2] == 7 and ref[2, 3, 1] == 0     # through the gap in row 0
    for k in (0, 2):
        assert np.array_equal(ref[k, :n[k]].T, astar_mod.plan_reference(g[k], st[k], go[k]))
    assert astar_mod.plan_reference(g[1], st[1], go[1]) is None


def test_argument_errors():
    L = _lib.lib()
    ip = C.POINTER(C.c_int32)
    occ = np.zeros((1, 4, 4), np.uint8)
    s = np.array([[0, 0]], np.int32); g = np.array([[3, 3]], np.int32)
    ref = np.zeros((1, 2, 3)); n = np.zeros(1, np.int32)
    call = lambda s_, g_, ml, gi=None: L.obca_b200_astar_batch(
        1, occ.ctypes.data, 1, 4, 4, gi, s_.ctypes.data_as(ip), g_.ctypes.data_as(ip), ml, ref.ctypes.data,
        n.ctypes.data_as(ip), 1)
    assert call(s, g, 2) == -5 and n[0] == -3                   # OBCA_E_SIZE: the route has 3 cells
    assert call(np.array([[0, 9]], np.int32), g, 2) == -1       # start outside the grid
    assert call(s, g, 1) == -1
    bad = np.array([3], np.int32)
    assert call(s, g, 2, bad.ctypes.data_as(ip)) == -1          # grid index out of range
    with pytest.raises(RuntimeError):
        astar_mod.plan_batch(occ[0], [[0, 0, 0]], [[3, 3, 0]], max_len=2)

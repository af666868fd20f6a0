"""The callers' side of the solve on the device: work-list solves, device-built obstacle rows and
the device-resident receding-horizon loop, against their host counterparts.  -m gpu only."""
import numpy as np
import pytest

import obca_testlib as common
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import (_abi, closed_loop as cl, demo_setting as ds,
                                                                               model_obstacle as mo, obca as obca_mod,
                                                                               scenario as sc)

pytestmark = pytest.mark.gpu


def test_indexed_solve_matches_plain_solve():
    """a work list on the device solves exactly the listed instances, bit-identically, and touches nothing else"""
    import torch
    B = 96
    b = sc.make_batch(2, B)
    prm, a = common.batch_arrays(b)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], B)
    t = lambda v: torch.as_tensor(v, dtype=torch.float64, device="cuda").contiguous()
    args = (t(a["x0"]), t(a["u0"]), t(a["xref"]), t(a["A"]), t(a["b0"]), None)
    full = s.solve(*args, T_max=t(a["T_max"]))
    torch.cuda.synchronize()
    rng = np.random.default_rng(3)
    pick = rng.permutation(B)[:37].astype(np.int32)
    index = torch.zeros(B, dtype=torch.int32, device="cuda"); index[:37] = torch.as_tensor(pick, device="cuda")
    count = torch.tensor([37], dtype=torch.int32, device="cuda")
    out = s.alloc_outputs(B, "cuda")
    for k in out:
        out[k].fill_(-7)
    s.solve(*args, T_max=t(a["T_max"]), out=out, index=index, count=count)
    torch.cuda.synchronize()
    rest = np.setdiff1d(np.arange(B), pick)
    for k in out:
        o = out[k].cpu().numpy(); f = full[k].cpu().numpy()
        assert np.array_equal(o[pick], f[pick]), k
        assert (o[rest] == -7).all(), k
    count.zero_()                                               # an empty list is a no-op
    for k in out:
        out[k].fill_(-7)
    s.solve(*args, T_max=t(a["T_max"]), out=out, index=index, count=count)
    torch.cuda.synchronize()
    assert all((out[k] == -7).all().item() for k in out)
    s.close()


def test_longest_first_order_changes_nothing_but_the_time():
    """A plain solve of 4,096 instances is handed out longest-first (difficulty estimate computed on the device from the
    inputs, csrc/obca_b200.cu); the same solve over an explicit work list keeps the caller's order.  Every instance is
    solved independently, so the two must agree bit for bit - and so must a reversed list."""
    import torch
    B = 4096
    b = sc.make_batch(3, B)
    prm, a = sc.batch_arrays(b)
    s = obca_mod.BatchSolver(prm, a["edge_ptr"], B)
    t = lambda v: torch.as_tensor(v, dtype=torch.float64, device="cuda").contiguous()
    args = (t(a["x0"]), t(a["u0"]), t(a["xref"]), t(a["A"]), t(a["b0"]), None)
    n0 = s.launches
    o = {k: v.clone() for k, v in s.solve(*args, T_max=t(a["T_max"])).items()}
    torch.cuda.synchronize()
    assert s.launches - n0 == 6            # estimate, rank, scatter, first pass, recovery block beside it, recovery over the device
    cnt = torch.tensor([B], dtype=torch.int32, device="cuda")
    for order in (np.arange(B), np.arange(B)[::-1]):
        idx = torch.as_tensor(np.ascontiguousarray(order, dtype=np.int32), device="cuda")
        n0 = s.launches
        o2 = s.solve(*args, T_max=t(a["T_max"]), index=idx, count=cnt)
        torch.cuda.synchronize()
        assert s.launches - n0 == 3
        for k in ("x", "u", "T", "obj", "lam", "mu", "status", "iters"):
            assert torch.equal(o[k], o2[k]), k
    # a key that does not compare (NaN pose in one reference window) must not cost any instance its place in the list
    xr = a["xref"].copy(); xr[7, 3, 2] = np.nan
    o3 = s.solve(args[0], args[1], t(xr), *args[3:], T_max=t(a["T_max"]))
    torch.cuda.synchronize()
    keep = np.arange(B) != 7
    assert int(o3["status"][7]) < 0
    for k in ("x", "u", "T", "obj", "status", "iters"):
        assert torch.equal(o[k][keep], o3[k][keep]), k
    s.close()


def test_device_rows_match_host_hrep():
    """obca_b200_build_rows == obstacle_H_Represent on the first time block (bit for bit, axis-aligned and slanted
    edges, two-vertex walls) and b0 + k*db == the reference's time-stacked rows"""
    import torch
    rng = np.random.default_rng(11)
    B, N = 257, 7
    vObs = [2, 5, 5, 4, 2]
    polys_b, info_b = [], []
    for i in range(B):
        th = rng.uniform(-np.pi, np.pi) if i % 3 else [0.0, np.pi / 2, -np.pi / 2][(i // 3) % 3]
        r1 = mo.get_obstacle(rng.uniform(5, 30), rng.uniform(2, 8), th, rng.uniform(1, 4), rng.uniform(1, 4))
        r2 = mo.get_obstacle(rng.uniform(5, 30), rng.uniform(2, 8), 0.0, 3.0, 3.0)
        tri = [[1.0, 1.0], [2.0 + rng.uniform(), 4.0], [5.0, 1.0 + (i % 2) * rng.uniform()], [1.0, 1.0]]
        polys_b.append([[[39, 9], [0, 9]], r1, r2, tri, [[0, 1], [39, 1]]])
        info_b.append([[0] * 11, [0, 0, rng.uniform(-3, 3), 0, 0, rng.uniform(0, 1)] + [0] * 5,
                       [0, 0, np.pi / 2, 0, 0, 0.3] + [0] * 5, [0] * 11, [0] * 11])
    verts = np.array([[v for p in polys for v in p] for polys in polys_b], float)            # (B, sum(vObs), 2)
    Ts = rng.uniform(0.05, 2.0, B)
    vel = np.array([[[r[5], np.cos(r[2]), np.sin(r[2])] for r in info] for info in info_b])
    t = lambda v: torch.as_tensor(v, dtype=torch.float64, device="cuda").contiguous()
    ep, A, b0, db = cl.build_rows_device(t(verts), vObs, t(vel), t(Ts))
    torch.cuda.synchronize()
    A, b0, db = A.cpu().numpy(), b0.cpu().numpy(), db.cpu().numpy()
    R = int(ep[-1])
    assert R == sum(vObs) - len(vObs)
    for i in range(B):
        Ah, bh = mo.obstacleModel().obstacle_H_Represent(len(vObs), vObs, polys_b[i])
        assert np.array_equal(A[i], Ah) and np.array_equal(b0[i], bh.reshape(-1)), i
        As, bs = mo.stacked_H_rep(polys_b[i], vObs, info_b[i], N, Ts[i])
        bs = bs.reshape(N + 1, R)
        stacked_A = As.reshape(N + 1, R, 2)
        if np.abs(stacked_A - Ah[None]).max() < 1e-9:           # translation kept every edge in its branch
            assert np.abs(b0[i][None] + np.arange(N + 1)[:, None] * db[i][None] - bs).max() <= 1e-9, i
    # static scene, scalar Ts, no db
    ep2, A2, b2, db2 = cl.build_rows_device(t(verts), vObs, None, 0.1, with_db=False)
    assert db2 is None and np.array_equal(A2.cpu().numpy(), A) and np.array_equal(b2.cpu().numpy(), b0)


def _setting():
    s = ds.problemSetting("demo9")
    s.senseDis = 8
    return s


@pytest.mark.parametrize("rule", ["shipped", "demo9"])
def test_device_loop_matches_host_loop(rule):
    """obca_b200_loop_* against ClosedLoopBatch (host-built inputs, same kernel): same modes and step counts, same
    trajectories (inputs differ only in the last bits of cos/sin/atan2/hypot)"""
    B, steps = 64, 10
    dyn = cl.demo9_monte_carlo(B)
    dyn[:, 1] = np.linspace(12, 30, B)
    dyn[:, 6] = np.arange(B) % 3
    h = cl.ClosedLoopBatch(_setting(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps)
    oh = h.run(rule); h.close()
    d = cl.ClosedLoopDevice(_setting(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps)
    od = d.run(rule)
    od2 = d.run(rule)                                           # reset + rerun on the same object
    d.close()
    assert od["solves"] == oh["solves"] or abs(od["solves"] - oh["solves"]) <= 0.02 * oh["solves"]
    assert sum(od["solves_by_mode"]) == od["solves"] and od["solves_by_mode"][0] > 0
    assert (od["mode"] != -1).sum() + (od["mode"] == _abi.MODE_FIXED_NOTERM).sum() == od["solves"]
    same = (od["steps"] == oh["steps"]) & (od["failed"] == oh["failed"]) & (od["mode"] == oh["mode"]).all(1)
    assert same.mean() >= 0.95, same.mean()
    err = [np.nanmax(np.abs(od["traj"][i] - oh["traj"][i])) for i in range(B) if same[i]]
    assert np.mean(np.array(err) <= 1e-6) >= 0.95, np.sort(err)[-5:]
    assert np.array_equal(np.isnan(od["traj"][same]), np.isnan(oh["traj"][same]))
    assert ((od["mode"] == _abi.MODE_FIXED_SET) | (od["mode"] == _abi.MODE_FIXED_NOTERM)).any()
    for k in ("traj", "steps", "failed", "mode"):
        assert np.array_equal(od[k], od2[k], equal_nan=True), k  # deterministic across runs
    # speculative fallback: the solve without the terminal set runs beside the one with it; same results
    sp = cl.ClosedLoopDevice(_setting(), dyn, N=5, Q_free=0.5, sense=8.0, max_steps=steps, speculative=True)
    os_ = sp.run(rule); sp.close()
    for k in ("traj", "steps", "failed", "mode", "Ts_opt"):
        assert np.array_equal(od[k], os_[k], equal_nan=True), k
    assert os_["solves"] == od["solves"]

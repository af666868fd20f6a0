#!/usr/bin/env python
"""Developer tool (CPU): independent audit of the cfg-5 instances the solver reports as failed ("35 % of cfg 5
is declared infeasible with no independent check").

cfg 5 = fixed-time OBCA (obca_mpc6) with Ts = 2 s, N = 20, 4 static + 2 moving boxes, terminal set x_N >= x0.x + 5,
1 <= y_N <= 9.  At Ts = 2 s the acceleration rows cannot bind (|dv| <= 1.2 m/s, |dw| <= pi/3 per step cover the whole input
box), so the feasible set of an instance is: a unicycle with v in [-0.6, 0.6], w in [-pi/6, pi/6], 20 Euler steps of 2 s,
pose point inside the map, ego rectangle at least dmin away from every obstacle AT THE 21 SAMPLE TIMES, terminal set.
Nothing of the solver, its NLP restatement or its derivatives is used here - only that geometry:

  class A   the terminal set lies outside the map (x0.x + 5 > xU.x)                          -> infeasible, trivially
  class B   the ego rectangle at the fixed start pose is closer than dmin to an obstacle     -> infeasible at k = 0
  search    breadth-first search over a pose lattice (dedup on a (dx, dx, dth) grid), discrete controls, exact
            separating-axis clearance test with margin dmin at every sample time.  A trajectory found is a PROOF of
            feasibility (the lattice controls are admissible inputs); none found at the finest lattice is evidence - not a
            proof - of infeasibility.
  confirm   a trajectory found is handed to the oracle as its start guess (OBCA_INIT_GUESS): if the oracle then solves
            the instance, the solver's failure from its own start was a miss.

    python tools/audit_cfg5.py [B=1024] [procs=8]      -> JSON on stdout (profiles/r2_cfg5_audit.json)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, scenario as sc  # noqa: E402

kw = dict(a.split("=") for a in sys.argv[1:])
B = int(kw.get("B", 1024)); PROCS = int(kw.get("procs", os.cpu_count() or 1))
LEVELS = [dict(dx=0.5, nth=24, nv=3, nw=3), dict(dx=0.25, nth=48, nv=5, nw=5), dict(dx=0.125, nth=96, nv=7, nw=7)]


def corners(P, ego):
    """ego rectangle corners of poses P (n,3) -> (n,4,2)"""
    c, s = np.cos(P[:, 2]), np.sin(P[:, 2])
    pts = np.array([(ego[0], ego[1]), (ego[0], -ego[3]), (-ego[2], -ego[3]), (-ego[2], ego[1])])
    x = P[:, None, 0] + c[:, None] * pts[None, :, 0] - s[:, None] * pts[None, :, 1]
    y = P[:, None, 1] + s[:, None] * pts[None, :, 0] + c[:, None] * pts[None, :, 1]
    return np.stack([x, y], -1)


def clear(P, polys, ego, margin):
    """separating-axis clearance: True where the ego rectangle at pose P is separated from every polygon by >= margin
    along one of the edge normals (sufficient for Euclidean distance >= margin)"""
    C = corners(P, ego)
    ok = np.ones(len(P), bool)
    th = P[:, 2]
    ego_axes = np.stack([np.stack([np.cos(th), np.sin(th)], -1), np.stack([-np.sin(th), np.cos(th)], -1)], 1)   # (n,2,2)
    for Q in polys:
        sep = np.full(len(P), -np.inf)
        e = np.roll(Q, -1, 0) - Q
        nrm = np.stack([e[:, 1], -e[:, 0]], -1); nrm = nrm / np.linalg.norm(nrm, axis=1)[:, None]
        for ax in nrm:                                   # polygon edge normals (fixed)
            p = C @ ax; q = Q @ ax
            sep = np.maximum(sep, np.maximum(q.min() - p.max(1), p.min(1) - q.max()))
        for j in range(2):                               # ego axes (per pose)
            ax = ego_axes[:, j]
            p = np.einsum("nij,nj->ni", C, ax); q = Q @ ax.T           # (n,4), (4,n)
            sep = np.maximum(sep, np.maximum(q.min(0) - p.max(1), p.min(1) - q.max(0)))
        ok &= sep >= margin
    return ok


def polys_at(b, k):
    """obstacle polygons at sample time k (static ones as they are, moving ones translated: demo_setting.py:461-468)"""
    out = []
    for i, poly in enumerate(b.polygons):
        Q = np.asarray(poly[:4], float)
        d = b.dyn[i] if i in b.dyn else None
        if d is not None:
            Q = Q + k * b.Ts * d[0] * np.array([np.cos(d[1]), np.sin(d[1])])
        out.append(Q)
    return out


def search(b, x0, lvl, margin):
    """BFS over the pose lattice; returns the pose sequence (N+1,3) of a trajectory into the terminal set, or None"""
    N, Ts = b.N, b.Ts
    dx, nth = lvl["dx"], lvl["nth"]
    vs = np.linspace(-0.6, 0.6, lvl["nv"]); ws = np.linspace(-np.pi / 6, np.pi / 6, lvl["nw"])
    V, W = [a.ravel() for a in np.meshgrid(vs, ws, indexing="ij")]
    xmin, xmax = max(b.xL[0], x0[0] - 6.0), min(b.xU[0], x0[0] + 16.0)
    layers = [np.asarray(x0, float)[None]]; parents = [np.array([-1])]
    for k in range(N):
        P = layers[-1]
        n = len(P)
        nx = P[:, None, 0] + Ts * V[None] * np.cos(P[:, None, 2]); ny = P[:, None, 1] + Ts * V[None] * np.sin(P[:, None, 2])
        nt = P[:, None, 2] + Ts * W[None] + 0 * nx
        Q = np.stack([nx.ravel(), ny.ravel(), nt.ravel()], -1); par = np.repeat(np.arange(n), len(V))
        keep = (Q[:, 0] >= xmin) & (Q[:, 0] <= xmax) & (Q[:, 1] >= b.xL[1]) & (Q[:, 1] <= b.xU[1])
        Q, par = Q[keep], par[keep]
        if len(Q) == 0:
            return None
        key = (np.floor(Q[:, 0] / dx).astype(np.int64) * 4096 + np.floor(Q[:, 1] / dx).astype(np.int64)) * 4096 + \
            np.floor(np.mod(Q[:, 2], 2 * np.pi) / (2 * np.pi / nth)).astype(np.int64)
        _, first = np.unique(key, return_index=True)
        Q, par = Q[first], par[first]
        ok = clear(Q, polys_at(b, k + 1), b.ego, margin)
        Q, par = Q[ok], par[ok]
        if len(Q) == 0:
            return None
        layers.append(Q); parents.append(par)
    P = layers[-1]
    hit = np.flatnonzero((P[:, 0] >= x0[0] + 5) & (P[:, 1] >= 1) & (P[:, 1] <= 9))
    if len(hit) == 0:
        return None
    i = int(hit[np.argmax(P[hit, 0])])
    traj = []
    for k in range(N, -1, -1):
        traj.append(layers[k][i]); i = int(parents[k][i])
    return np.array(traj[::-1])


_B = None


def prepare(b):
    """attach the motion (speed, heading) of the two moving boxes to a cfg-5 Batch (make_batch appends them after the
    static obstacles): recovered from the time-stacked rows, b_k = b_0 + k Ts v A (cos, sin)"""
    nq = len(b.polygons) - 2
    b.dyn = {}
    AObs = np.asarray(b.AObs).reshape(b.N + 1, -1, 2); bObs = np.asarray(b.bObs).reshape(b.N + 1, -1)
    R0 = 4 * nq
    for j in range(2):
        rows = slice(R0 + 4 * j, R0 + 4 * j + 4)
        db = (bObs[b.N, rows] - bObs[0, rows]) / b.N
        vel = np.linalg.lstsq(AObs[0, rows], db, rcond=None)[0] / b.Ts
        b.dyn[nq + j] = (float(np.hypot(*vel)), float(np.arctan2(vel[1], vel[0])))
    return b


def audit_one(i):
    b = _B
    x0 = b.x0[i]
    for li, lvl in enumerate(LEVELS):
        t = time.time()
        tr = search(b, x0, lvl, b.dmin + 1e-3)
        if tr is not None:
            return i, li, tr, time.time() - t
    return i, -1, None, 0.0


def main():
    global _B
    from oracle import c_oracle
    b = prepare(sc.make_batch(5, B))
    _B = b
    prm, a = sc.batch_arrays(b, init=_abi.INIT_WARM)
    c = c_oracle.solve(prm, a["x0"], a["u0"], a["xref"], a["edge_ptr"], a["A"], a["b0"], a["db"], term=a["term"],
                       nthreads=os.cpu_count() or 1)
    st = c["status"]
    failed = np.flatnonzero(st < 0)
    polys0 = polys_at(b, 0)
    clsA = b.x0[:, 0] + 5 > b.xU[0]
    clsB = ~clear(b.x0, polys0, b.ego, b.dmin)
    # sanity of the geometry against the solver's own successes: every solved trajectory must pass the same clearance test
    chk = np.flatnonzero(st >= 0)[:64]
    viol = 0
    for i in chk:
        for k in range(b.N + 1):
            viol += int(not clear(c["x"][i, k][None], polys_at(b, k), b.ego, b.dmin - 1e-6)[0])
    todo = [int(i) for i in failed if not clsA[i] and not clsB[i]]
    from multiprocessing import Pool
    t0 = time.time()
    with Pool(PROCS) as pool:
        res = pool.map(audit_one, todo, chunksize=1)
    found = {i: (li, tr) for i, li, tr, _ in res if tr is not None}
    # confirmation: the oracle from the audit's trajectory (OBCA_INIT_GUESS reads the poses from the output array x)
    confirmed = []
    if found:
        idx = np.array(sorted(found))
        prm_g, _ = sc.batch_arrays(b, init=_abi.INIT_GUESS)
        guess = np.stack([found[i][1] for i in idx])
        g = c_oracle.solve(prm_g, a["x0"][idx], a["u0"][idx], a["xref"][idx], a["edge_ptr"], a["A"], a["b0"], a["db"],
                           term=a["term"][idx], nthreads=os.cpu_count() or 1, guess=guess)
        confirmed = [int(i) for i, s_ in zip(idx, g["status"]) if s_ >= 0]
    hist = lambda m: {int(k): int(v) for k, v in zip(*np.unique(st[m], return_counts=True))}
    out = dict(workload="cfg 5, %d instances, oracle from the warm start (restoration on)" % B, status=hist(np.ones(B, bool)),
               failed=int(len(failed)),
               class_A_terminal_set_outside_map=int((clsA & (st < 0)).sum()),
               class_B_start_pose_in_collision=int((clsB & ~clsA & (st < 0)).sum()),
               solved_instances_in_class_A_or_B=int(((clsA | clsB) & (st >= 0)).sum()),
               searched=len(todo), search_levels=LEVELS,
               trajectory_found=len(found), found_by_level=[int(sum(1 for v in found.values() if v[0] == l)) for l in range(len(LEVELS))],
               found_status=hist(np.isin(np.arange(B), list(found))) if found else {},
               confirmed_by_oracle_from_the_audit_guess=len(confirmed),
               no_trajectory_on_the_finest_lattice=len(todo) - len(found),
               none_found_status=hist(np.isin(np.arange(B), [i for i in todo if i not in found])) if len(todo) > len(found) else {},
               clearance_violations_of_64_solved_trajectories=int(viol), search_seconds=round(time.time() - t0, 1))
    out["failed_explained_infeasible"] = out["class_A_terminal_set_outside_map"] + out["class_B_start_pose_in_collision"] + out["no_trajectory_on_the_finest_lattice"]
    out["solver_misses"] = out["trajectory_found"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

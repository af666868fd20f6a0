"""Synthetic OBCA-MPC workloads and the reference's MPC constants.

Everything here is host-side NumPy and deterministic in ``numpy.random.default_rng(20221209 + cfg)``.
Scene = a 40 x 11 corridor (xL=[0,0], xU=[39,10]) with rotated-rectangle obstacles; each instance is one
ego start pose with its A* reference window - the arguments ``closedLoop`` would hand to
``obca.obca_mpc4`` / ``obca_mpc6`` (/root/reference/src/closed_loop.py:118,131,382,389).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import model_obstacle as mo
from .a_star import plan_batch, plan_reference

# closed_loop.py:32-104
TS = 0.1
U_L = np.array([-0.6, -np.pi / 6]); U_U = np.array([0.6, np.pi / 6])
EGO = np.array([1.7, 0.75, 1.7, 0.75])
DMIN = 0.05
Q_FREE = 0.1 * np.eye(3); R_FREE = [0.01 * np.eye(2), 0.1 * np.eye(2)]
Q_FIX = 0.001 * np.eye(3); R_FIX = [0.01 * np.eye(2), 1.0 * np.eye(2)]


@dataclasses.dataclass
class Batch:
    """One synthetic batch: shared scene + per-instance start poses / reference windows (all float64)."""
    mode: int
    N: int
    Ts: float
    Q: np.ndarray
    P: np.ndarray
    R: list
    xL: np.ndarray
    xU: np.ndarray
    uL: np.ndarray
    uU: np.ndarray
    ego: np.ndarray
    dmin: float
    nObs: int
    vObs: list
    AObs: np.ndarray          # ((N+1)*R, 2) time-stacked exactly as closed_loop.py:500 builds it
    bObs: np.ndarray          # ((N+1)*R, 1)
    x0: np.ndarray            # (B, 3)
    u0: np.ndarray            # (B, 2)
    xref: np.ndarray          # (B, 3, N+1)
    terminal_set: np.ndarray | None = None   # (B, 2, 2)
    polygons: list | None = None

    @property
    def B(self):
        return self.x0.shape[0]


def ego_corners(pose, ego=EGO):
    x, y, th = pose
    c, s = np.cos(th), np.sin(th)
    pts = [(ego[0], ego[1]), (ego[0], -ego[3]), (-ego[2], -ego[3]), (-ego[2], ego[1])]
    return np.array([[x + c * a - s * b, y + s * a + c * b] for a, b in pts])


def _sat_separated(P, Q, margin):
    """Separating-axis test between two convex polygons (vertex arrays), with a margin."""
    for poly in (P, Q):
        n = len(poly)
        for i in range(n):
            e = poly[(i + 1) % n] - poly[i]
            nrm = np.array([e[1], -e[0]])
            ln = np.linalg.norm(nrm)
            if ln == 0:
                continue
            nrm = nrm / ln
            p = P @ nrm; q = Q @ nrm
            if p.max() + margin < q.min() or q.max() + margin < p.min():
                return True
    return False


def pose_clear(pose, polygons, margin):
    c = ego_corners(pose)
    return all(_sat_separated(c, np.asarray(poly[:4], float), margin) for poly in polygons)


def make_scene(rng, n_quads, cx_range, cy_range, side=(2.0, 4.0), gap=4.0, max_tries=10000):
    """Rotated rectangles, resampled until pairwise AABB gap >= ``gap`` (SURVEY 8(d) cfg 2/3)."""
    polys = []; boxes = []
    tries = 0
    while len(polys) < n_quads:
        tries += 1
        if tries > max_tries:
            polys, boxes, tries = [], [], 0
        cx = rng.uniform(*cx_range); cy = rng.uniform(*cy_range)
        l = rng.uniform(*side); w = rng.uniform(*side); th = rng.uniform(-np.pi / 4, np.pi / 4)
        v = mo.get_obstacle(cx, cy, th, l, w)
        a = np.asarray(v[:4])
        bb = (a[:, 0].min(), a[:, 0].max(), a[:, 1].min(), a[:, 1].max())
        ok = True
        for o in boxes:
            dx = max(o[0] - bb[1], bb[0] - o[1]); dy = max(o[2] - bb[3], bb[2] - o[3])
            if max(dx, dy) < gap:
                ok = False
                break
        if ok:
            polys.append(v); boxes.append(bb)
    return polys


def update_reference_trajectory(N, ref, x0):
    """Window of N+1 consecutive path points starting at the one closest to x0, clamped to the last
    (closed_loop.py:502-528)."""
    M = ref.shape[1]
    d = (x0[0] - ref[0]) ** 2 + (x0[1] - ref[1]) ** 2
    i0 = int(np.argmin(d))            # first minimum, as the reference's strict '<' scan (514)
    idx = np.minimum(np.arange(N + 1) + i0, M - 1)
    return ref[:, idx].copy()


def sample_instances(rng, grid, static_polys, goal, N, B, free, x_cells=(1, 36), y_cells=(1, 10), native_planner=True):
    """B start poses on free grid cells (AABB clearance 2) with their A* reference windows; poses whose start or
    terminal pose collides, or whose T_max (obca.py:961-962) leaves no room for the distance, are redrawn.
    ``native_planner=False`` plans with the Python A* (a_star.py; identical paths, tests/test_planner.py) so that the
    CUDA library is never loaded - what bench.py's CPU arm does."""
    cells = [(cx, cy) for cx in range(*x_cells) for cy in range(*y_cells)
             if grid[max(cy - 2, 0):cy + 3, max(cx - 2, 0):cx + 3].sum() == 0]
    if native_planner:
        # one native batched A* call for every candidate cell (identical to plan_reference per cell)
        pref, plen_ = plan_batch(grid, [[cx, cy, 0] for cx, cy in cells], [goal])
        paths = {c: (pref[j, :plen_[j]].T.copy() if plen_[j] > 0 else None) for j, c in enumerate(cells)}
    else:
        paths = {c: plan_reference(grid, (c[0], c[1], 0), goal) for c in cells}
    x0 = np.zeros((B, 3)); xref = np.zeros((B, 3, N + 1))
    n = 0
    while n < B:
        cx, cy = cells[rng.integers(len(cells))]
        th0 = rng.uniform(-np.pi / 4, np.pi / 4)
        ref = paths[(cx, cy)]
        if ref is None or ref.shape[1] < 3:
            continue
        pose = np.array([cx, cy, th0], float)
        win = update_reference_trajectory(N, ref, pose)
        if not pose_clear(pose, static_polys, DMIN + 0.05):
            continue
        if free:
            if not pose_clear(win[:, N], static_polys, DMIN + 0.05):
                continue
            # Tmax of obca.py:961-962 must leave room for the distance to cover
            Tmax = ((win[0, N] - cx) + (win[1, N] - cy)) / (N * U_U[0] * TS) + 1
            seg = np.diff(np.concatenate([pose[:2, None], win[:2]], axis=1), axis=1)
            plen = np.sqrt((seg ** 2).sum(0)).sum()
            if Tmax * N * U_U[0] * TS < 1.02 * plen + 0.3:
                continue
        x0[n] = pose; xref[n] = win
        n += 1
    return x0, xref


def make_batch(cfg, B, N=None, n_quads=None, seed=None, goal=(38, 4, 0), pose_seed=None, native_planner=True):
    """cfg 2: N=10, 2 quads, FREE.  cfg 3: N=20, 4 quads, FREE (headline).  cfg 5: cfg 3 + 2 dynamic boxes,
    FIXED_SET with Ts = 2.0 and terminal set [x0.x+5, 99] x [1, 9] (closed_loop.py:371).
    ``pose_seed`` draws the start poses from their own stream (same scene, different instances: one shard per
    rank in the multi-GPU runs)."""
    from . import obca as _o
    rng = np.random.default_rng((20221209 + cfg) if seed is None else seed)
    if cfg == 2:
        N = 10 if N is None else N; nq = 2 if n_quads is None else n_quads
        rx, ry = (8, 30), (3, 7)
    else:
        N = 20 if N is None else N; nq = 4 if n_quads is None else n_quads
        rx, ry = (6, 34), (2, 8)
    xL = np.array([0.0, 0.0]); xU = np.array([39.0, 10.0])
    while True:
        polys = make_scene(rng, nq, rx, ry)
        grid = mo.shape2grid([40, 11], [p[:4] for p in polys])
        if grid[int(goal[1]), int(goal[0])] == 0 and plan_reference(grid, (1, 5, 0), goal) is not None:
            break
    vObs = [5] * nq
    info = [[0] * 11 for _ in range(nq)]
    mode = _o.MODE_FREE
    Ts = TS
    if cfg == 5:
        mode = _o.MODE_FIXED_SET
        Ts = 2.0
        for sgn in (+1, -1):
            cx = rng.uniform(12, 30); sp = rng.uniform(0.05, 0.2)
            cy = 0.0 if sgn > 0 else 10.0
            polys.append(mo.get_obstacle(cx, cy, sgn * np.pi / 2, 3, 3))
            info.append([cx, cy, sgn * np.pi / 2, 3, 3, sp, 0, 0, 0, 0, 0])
            vObs.append(5)
    AObs, bObs = mo.stacked_H_rep(polys, vObs, info, N, Ts)

    if pose_seed is not None:
        rng = np.random.default_rng(pose_seed)
    x0, xref = sample_instances(rng, grid, polys[:nq], goal, N, B, mode == _o.MODE_FREE, native_planner=native_planner)
    free = mode == _o.MODE_FREE
    ts = None
    if cfg == 5:
        ts = np.zeros((B, 2, 2))
        ts[:, 0, 0] = x0[:, 0] + 5; ts[:, 0, 1] = 99; ts[:, 1, 0] = 1; ts[:, 1, 1] = 9
    return Batch(mode=mode, N=N, Ts=Ts, Q=(Q_FREE if free else Q_FIX).copy(), P=(Q_FREE if free else Q_FIX).copy(),
                 R=[r.copy() for r in (R_FREE if free else R_FIX)], xL=xL, xU=xU, uL=U_L.copy(), uU=U_U.copy(),
                 ego=EGO.copy(), dmin=DMIN, nObs=len(vObs), vObs=vObs, AObs=AObs, bObs=bObs, x0=x0,
                 u0=np.zeros((B, 2)), xref=xref, terminal_set=ts, polygons=polys)


def batch_arrays(b, init=None, **opts):
    """A Batch as the arrays of the C-ABI (include/obca_b200.h): -> (obca_params, dict(x0 [B,3], u0 [B,2],
    xref [B,N+1,3], edge_ptr, A [R,2], b0 [R], db [R] | None, T_max [B] | None, term [B,3] | None)).  ``opts`` go to
    _abi.make_params (mu_init, bound_push, retry, ...)."""
    from . import _abi
    ep, A, b0, db = _abi.pack_obstacles(b.mode, b.N, b.nObs, b.vObs, b.AObs, b.bObs)
    prm = _abi.make_params(b.mode, b.N, b.nObs, int(ep[-1]), b.Ts, b.P, b.Q, b.R, b.xL, b.xU, b.uL, b.uU, b.dmin,
                           b.ego, init=_abi.INIT_WARM if init is None else init, **opts)
    xref = np.ascontiguousarray(b.xref.transpose(0, 2, 1))
    Tm = None
    if _abi.is_free(b.mode):
        Tm = ((b.xref[:, 0, b.N] - b.x0[:, 0]) + (b.xref[:, 1, b.N] - b.x0[:, 1])) / (b.N * b.uU[0] * b.Ts) + 1.0
    term = None
    if b.terminal_set is not None:
        term = np.stack([b.terminal_set[:, 0, 0], b.terminal_set[:, 1, 0], b.terminal_set[:, 1, 1]], axis=1)
    return prm, dict(x0=b.x0, u0=b.u0, xref=xref, edge_ptr=ep, A=A, b0=b0, db=db, T_max=Tm, term=term)


def regular_polygon(cx, cy, radius, sides, phase=0.0):
    """Clockwise regular polygon, first vertex repeated (the reference's vertex-list convention)."""
    ang = phase - 2 * np.pi * np.arange(sides) / sides
    v = [[cx + radius * np.cos(a), cy + radius * np.sin(a)] for a in ang]
    return v + [v[0]]


def make_polygon_batch(sides, B, N, seed=0, map_size=(60, 24), moving=0, mode=None):
    """A scene of convex polygons with ``sides[i]`` edges each (3..8) on a ``map_size`` field, laid out on a jittered
    lattice so that they do not touch; the last ``moving`` of them drift (FIXED modes).  Exercises ragged edge counts,
    the 8-edge kernel variant and the size limits (N + 1 <= 32 stages, 12 obstacles, 48 rows)."""
    from . import obca as _o
    rng = np.random.default_rng(seed)
    W, H = map_size
    n = len(sides)
    cols = int(np.ceil(n / 2))
    while True:
        polys, info = [], []
        for i, k in enumerate(sides):
            cx = 8 + (W - 18) * ((i // 2) + 0.5) / cols + rng.uniform(-0.8, 0.8)
            cy = (H * 0.27 if i % 2 == 0 else H * 0.73) + rng.uniform(-1.0, 1.0)
            polys.append(regular_polygon(cx, cy, rng.uniform(1.2, 2.0), int(k), rng.uniform(0, 2 * np.pi)))
            info.append([0] * 11)
        goal = (W - 3, H // 2, 0)
        grid = mo.shape2grid([W, H], [p[:-1] for p in polys])
        if grid[int(goal[1]), int(goal[0])] == 0 and plan_reference(grid, (1, H // 2, 0), goal) is not None:
            break
    if mode is None:
        mode = _o.MODE_FIXED_SET if moving else _o.MODE_FREE
    free = mode in (_o.MODE_FREE, _o.MODE_FREE_STACKED)
    Ts = TS if free else 1.0
    for i in range(n - moving, n):
        info[i] = [0, 0, rng.uniform(-np.pi, np.pi), 0, 0, rng.uniform(0.02, 0.08), 0, 0, 0, 0, 0]
    vObs = [int(k) + 1 for k in sides]
    AObs, bObs = mo.stacked_H_rep(polys, vObs, info, N, Ts)
    x0, xref = sample_instances(rng, grid, polys, goal, N, B, free, x_cells=(1, W - 8), y_cells=(1, H - 1))
    ts = None
    if mode == _o.MODE_FIXED_SET:
        ts = np.zeros((B, 2, 2))
        ts[:, 0, 0] = x0[:, 0] + 3; ts[:, 0, 1] = 99; ts[:, 1, 0] = 0; ts[:, 1, 1] = H
    return Batch(mode=mode, N=N, Ts=Ts, Q=(Q_FREE if free else Q_FIX).copy(), P=(Q_FREE if free else Q_FIX).copy(),
                 R=[r.copy() for r in (R_FREE if free else R_FIX)], xL=np.array([0.0, 0.0]), xU=np.array([W - 1.0, H - 1.0]),
                 uL=U_L.copy(), uU=U_U.copy(), ego=EGO.copy(), dmin=DMIN, nObs=n, vObs=vObs, AObs=AObs, bObs=bObs, x0=x0,
                 u0=np.zeros((B, 2)), xref=xref, terminal_set=ts, polygons=polys)

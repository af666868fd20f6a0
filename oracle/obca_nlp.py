"""ORACLE (test infrastructure, not product code): NumPy restatement of the reference's OBCA NLPs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this package.  The product path (``vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200``)
never does.

**Parity unpinned**: the reference hands these NLPs to CasADi ``Opti`` + IPOPT (third-party, not vendored,
version not pinned, not installable here - SURVEY.md 8(c)); it ships no tests or golden outputs.  This file
restates the *problem* (cost, constraints, bounds, start point) line by line from ``src/obca.py``; the
*solver* in ``ipm_dense.py`` is our own primal-dual interior-point method (IPOPT-style defaults).

Reference lines restated here (all in /root/reference/src/obca.py):

* MODE_FREE          obca_mpc4  828-1071  (vars 842-856, cost 859-897, dynamics 902-911, bounds 916-923,
                                           accel 928-939, init/terminal 944/951, duals>=0 + T bounds 956-963,
                                           obstacle rows from the FIRST time block 968-1042)
* MODE_FIXED_SET     obca_mpc6  1361-1562 (cost 1385-1414, dynamics 1419-1422, terminal set 1465-1466,
                                           obstacle rows advance through the time-stacked AObs 1482-1536)
* MODE_FIXED_NOTERM  obca_mpc8  1564-1758 (as mpc6, no terminal constraint 1662-1665)
* MODE_FREE_STACKED  obca2(fixtime=0) 338-629 (as mpc4 but time-stacked rows 538, optional uref 421-424)

Compact variable set: the reference declares ``l`` as rows(AObs) x (N+1)
(obca.py:848) but only R*(N+1) of those entries appear in a constraint other than ``l >= 0``; the unused
ones have no influence on (x, u, T, objective) at a KKT point and are dropped.  ``Topt[0..N]`` are chained
equal (obca.py:911) so they are one scalar ``T``; the cost term sum_k (10*Topt[k] + Topt[k]**2) becomes
(N+1)*(10*T + T**2).  ``x[:,0] == x0`` (obca.py:944) is eliminated by substitution.

Variable vector X (all float64):
    z_k = (x,y,theta)_k, k=1..N | u_k = (v,w)_k, k=0..N-1 | T (free modes) |
    for k=0..N, for obstacle i: lambda_{k,i} (E_i), mu_{k,i} (4)
Equalities c(X) = 0:
    dyn_k (3) k=0..N-1 | terminal (3, free modes) | for k, i: e1, e2
Inequalities g(X) >= 0 (every one - simple bounds included - is a general constraint with its own slack,
as ``Opti`` hands them to IPOPT):
    for k=1..N: x-xL, y-yL, xU-x, yU-y (theta free, obca.py:916 ``range(nx-1)``) |
    for k=0..N-1: u-uL (2), uU-u (2), accel+amax (2), amax-accel (2) | T-Tmin, Tmax-T |
    terminal set (FIXED_SET): x_N-ts00, y_N-ts10, ts11-y_N |
    for k, i: lambda (E_i), mu (4), 1-|A^T lambda|^2, dist-dmin
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

MODE_FREE = 0          # obca_mpc4
MODE_FIXED_SET = 1     # obca_mpc6
MODE_FIXED_NOTERM = 2  # obca_mpc8
MODE_FREE_STACKED = 3  # obca2, fixtime == 0
MODE_FIXED_OBCA2 = 4   # obca2, fixtime == 1 (terminal_set optional: [] -> none)

INF = 1e20


@dataclasses.dataclass
class Problem:
    mode: int
    N: int
    Ts: float
    Q: np.ndarray          # (3,3) symmetrised
    P: np.ndarray
    R1: np.ndarray         # (2,2)
    R2: np.ndarray
    x0: np.ndarray         # (3,)
    u0: np.ndarray         # (2,)
    xref: np.ndarray       # (3,N+1)
    uref: np.ndarray       # (2,N)
    xL: np.ndarray         # (2,)
    xU: np.ndarray
    uL: np.ndarray
    uU: np.ndarray
    amax: np.ndarray       # (2,)  {0.6, pi/6}
    edges: np.ndarray      # (nObs,) int, E_i = vObs[i]-1
    A: np.ndarray          # (N+1, R, 2)  per-time-step rows actually used
    b: np.ndarray          # (N+1, R)
    dmin: float
    g: np.ndarray          # (4,)  [L/2, W/2, L/2, W/2]
    off: float
    tcost: tuple = (10.0, 1.0)
    Tmin: float = 1e-4
    Tmax: float = 1.0
    term: Optional[np.ndarray] = None   # (3,) xmin, ymin, ymax (FIXED_SET)

    @property
    def free(self):
        return self.mode in (MODE_FREE, MODE_FREE_STACKED)

    @property
    def nobs(self):
        return len(self.edges)

    @property
    def R(self):
        return int(np.sum(self.edges))


def build_problem(mode, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0,
                  uref=None, terminal_set=None) -> Problem:
    """Pack the reference's positional arguments (closed_loop.py:118/131/137/170) into a Problem."""
    Q = np.asarray(Q, float); P = np.asarray(P, float)
    R1 = np.asarray(R[0], float); R2 = np.asarray(R[1], float)
    sym = lambda M: 0.5 * (M + M.T)
    x0 = np.asarray(x0, float).reshape(3)
    u0 = np.asarray(u0, float).reshape(2)
    xref = np.asarray(xref, float).reshape(3, N + 1)
    edges = np.asarray([int(v) - 1 for v in list(vObs)[:nObs]], int)
    Rr = int(edges.sum())
    AObs = np.asarray(AObs, float).reshape(-1, 2)
    bObs = np.asarray(bObs, float).reshape(-1)
    A = np.zeros((N + 1, Rr, 2)); b = np.zeros((N + 1, Rr))
    for k in range(N + 1):
        # mpc4 resets the row counter every k (obca.py:969) -> first block; the others advance (1482, 538)
        blk = 0 if mode == MODE_FREE else k
        A[k] = AObs[blk * Rr:(blk + 1) * Rr]
        b[k] = bObs[blk * Rr:(blk + 1) * Rr]
    L = ego[0] + ego[2]; W = ego[1] + ego[3]
    g = np.array([L / 2, W / 2, L / 2, W / 2])
    off = (ego[0] + ego[2]) / 2 - ego[2]                       # obca.py:1026
    free = mode in (MODE_FREE, MODE_FREE_STACKED)
    Tmax = 1.0
    if free:
        dis = (xref[0, N] - x0[0]) + (xref[1, N] - x0[1])      # obca.py:961 (signed sum)
        Tmax = dis / (N * uU[0] * Ts) + 1                      # obca.py:962
    term = None
    if mode == MODE_FIXED_SET or (mode == MODE_FIXED_OBCA2 and terminal_set is not None and len(terminal_set)):
        ts = np.asarray(terminal_set, float)
        term = np.array([ts[0, 0], ts[1, 0], ts[1, 1]])        # obca.py:1465-1466
    if uref is None or len(uref) == 0:
        uref = np.zeros((2, N))
    return Problem(mode=mode, N=N, Ts=float(Ts), Q=sym(Q), P=sym(P), R1=sym(R1), R2=sym(R2), x0=x0, u0=u0,
                   xref=xref, uref=np.asarray(uref, float).reshape(2, N), xL=np.asarray(xL, float)[:2],
                   xU=np.asarray(xU, float)[:2], uL=np.asarray(uL, float), uU=np.asarray(uU, float),
                   amax=np.array([0.6, np.pi / 6]), edges=edges, A=A, b=b, dmin=float(dmin), g=g, off=float(off),
                   Tmax=float(Tmax), term=term)


def problem_from_abi(prm, edge_ptr, x0, u0, xref, A, b0, db=None, Ts=None, T_max=None, term=None, uref=None) -> Problem:
    """A Problem from ABI-level arrays of ONE instance (include/obca_b200.h layouts: xref (N+1,3), A (R,2), b0 (R,),
    db (R,) with b_k = b0 + k*db) and the obca_params struct - what the C oracle and the kernel are handed."""
    N = int(prm.N)
    ep = np.asarray(edge_ptr, int)
    A = np.asarray(A, float); b0 = np.asarray(b0, float)
    stacked = int(prm.mode) != MODE_FREE and db is not None
    b = np.stack([b0 + (k * np.asarray(db, float) if stacked else 0.0) for k in range(N + 1)])
    sym = lambda M: 0.5 * (M + M.T)
    m = lambda v, shape: np.array(list(v), float).reshape(shape)
    ego = m(prm.ego, 4)
    L = ego[0] + ego[2]; W = ego[1] + ego[3]
    free = int(prm.mode) in (MODE_FREE, MODE_FREE_STACKED)
    return Problem(mode=int(prm.mode), N=N, Ts=float(prm.Ts if Ts is None else Ts), Q=sym(m(prm.Q, (3, 3))), P=sym(m(prm.P, (3, 3))),
                   R1=sym(m(prm.R1, (2, 2))), R2=sym(m(prm.R2, (2, 2))), x0=np.asarray(x0, float).reshape(3),
                   u0=np.asarray(u0, float).reshape(2), xref=np.ascontiguousarray(np.asarray(xref, float).reshape(N + 1, 3).T),
                   uref=np.zeros((2, N)) if uref is None else np.ascontiguousarray(np.asarray(uref, float).reshape(N, 2).T),
                   xL=m(prm.xL, 2), xU=m(prm.xU, 2), uL=m(prm.uL, 2), uU=m(prm.uU, 2), amax=m(prm.acc_max, 2),
                   edges=np.diff(ep), A=np.tile(A[None], (N + 1, 1, 1)), b=b, dmin=float(prm.dmin),
                   g=np.array([L / 2, W / 2, L / 2, W / 2]), off=float(L / 2 - ego[2]), tcost=tuple(prm.time_cost),
                   Tmin=float(prm.T_min), Tmax=float(T_max) if (free and T_max is not None) else 1.0,
                   term=None if term is None else np.asarray(term, float).reshape(3))


class Layout:
    """Index maps of the compact NLP (variables X, equalities c, inequalities g >= 0)."""

    def __init__(self, p: Problem):
        N, nobs = p.N, p.nobs
        self.N = N
        n = 0
        self.z = n; n += 3 * N            # z_k at z + 3(k-1), k=1..N
        self.u = n; n += 2 * N
        self.T = n if p.free else -1
        n += 1 if p.free else 0
        self.ntraj = n
        self.eoff = np.concatenate([[0], np.cumsum(p.edges)]).astype(int)
        self.lam = np.zeros((N + 1, nobs), int); self.mu = np.zeros((N + 1, nobs), int)
        for k in range(N + 1):
            for i in range(nobs):
                self.lam[k, i] = n; n += int(p.edges[i])
                self.mu[k, i] = n; n += 4
        self.n = n
        m = 0
        self.c_dyn = m; m += 3 * N
        self.c_term = m if p.free else -1
        m += 3 if p.free else 0
        self.c_e = np.zeros((N + 1, nobs), int)
        for k in range(N + 1):
            for i in range(nobs):
                self.c_e[k, i] = m; m += 2
        self.m = m
        q = 0
        self.g_xy = q; q += 4 * N          # stage k (1..N): [x-xL, y-yL, xU-x, yU-y] at g_xy + 4(k-1)
        self.g_u = q; q += 8 * N           # stage k: [u-uL (2), uU-u (2), acc+amax (2), amax-acc (2)]
        self.g_T = q if p.free else -1
        q += 2 if p.free else 0
        self.g_term = q if p.term is not None else -1
        q += 3 if p.term is not None else 0
        self.g_w = np.zeros((N + 1, nobs), int)   # block: lambda (E), mu (4), norm, dist
        for k in range(N + 1):
            for i in range(nobs):
                self.g_w[k, i] = q; q += int(p.edges[i]) + 6
        self.q = q

    def iz(self, k):
        return self.z + 3 * (k - 1)

    def iu(self, k):
        return self.u + 2 * k


def start_point(p: Problem, lay: Layout, init="zero"):
    """init="zero": the reference's start - every Opti variable 0 except Topt = 1 (obca.py:856).
    init="xref": poses from the A* reference window, rest as "zero".
    init="warm": A* warm start - poses from the reference window, T from the window's arc length, inputs
    by finite differences (clipped), OBCA duals from the most-separating obstacle face."""
    X = np.zeros(lay.n)
    N = p.N
    if p.free:
        X[lay.T] = 1.0
    if init in ("xref", "warm"):
        for k in range(1, N + 1):
            X[lay.iz(k):lay.iz(k) + 3] = p.xref[:, k]
    if init == "warm":
        P = np.concatenate([p.x0[:, None], p.xref[:, 1:]], axis=1)
        seg = np.sqrt(np.sum(np.diff(P[:2], axis=1) ** 2, axis=0))
        if p.free:
            T0 = seg.sum() / (N * p.uU[0] * p.Ts)
            T0 = min(max(T0, 1.0), max(p.Tmax, p.Tmin))
            X[lay.T] = T0
            h = T0 * p.Ts
        else:
            h = p.Ts
        for k in range(N):
            dth = P[2, k + 1] - P[2, k]
            dth = (dth + np.pi) % (2 * np.pi) - np.pi
            fwd = np.cos(P[2, k]) * (P[0, k + 1] - P[0, k]) + np.sin(P[2, k]) * (P[1, k + 1] - P[1, k])
            X[lay.iu(k)] = np.clip(fwd / h, p.uL[0], p.uU[0])
            X[lay.iu(k) + 1] = np.clip(dth / h, p.uL[1], p.uU[1])
        for k in range(N + 1):
            zk = P[:, k]
            ct, st = np.cos(zk[2]), np.sin(zk[2])
            t = np.array([zk[0] + p.off * ct, zk[1] + p.off * st])
            for i in range(p.nobs):
                E = int(p.edges[i]); o = lay.eoff[i]
                Ai = p.A[k, o:o + E]; bi = p.b[k, o:o + E]
                nrm = np.sqrt((Ai ** 2).sum(1))
                sep = (Ai @ t - bi) / nrm
                j = int(np.argmax(sep))
                lam = np.zeros(E); lam[j] = 0.9 / nrm[j]
                a1, a2 = Ai[:, 0] @ lam, Ai[:, 1] @ lam
                r1 = -(ct * a1 + st * a2); r2 = -(-st * a1 + ct * a2)
                mu = np.array([max(r1, 0), max(r2, 0), max(-r1, 0), max(-r2, 0)])
                X[lay.lam[k, i]:lay.lam[k, i] + E] = lam
                X[lay.mu[k, i]:lay.mu[k, i] + 4] = mu
    return X


def pose(p, lay, X, k):
    return p.x0 if k == 0 else X[lay.iz(k):lay.iz(k) + 3]


def evaluate(p: Problem, lay: Layout, X, y=None, zi=None, want=("f", "g", "c", "J", "d", "Jd", "W")):
    """Objective f, gradient g, equality residual c with dense Jacobian J (m,n), inequality values d (>= 0
    feasible) with Jacobian Jd (q,n), dense Hessian of the Lagrangian
    W = d2f + sum_j y_j d2c_j - sum_i zi_i d2d_i.  Straight loops, written for clarity not speed."""
    N = p.N
    free = p.free
    T = X[lay.T] if free else 1.0
    h = T * p.Ts
    out = {}
    need_g = "g" in want; need_J = "J" in want or "Jd" in want; need_W = "W" in want
    f = 0.0
    g = np.zeros(lay.n) if need_g else None
    c = np.zeros(lay.m)
    d = np.zeros(lay.q)
    J = np.zeros((lay.m, lay.n)) if need_J else None
    Jd = np.zeros((lay.q, lay.n)) if need_J else None
    W = np.zeros((lay.n, lay.n)) if need_W else None
    iT = lay.T
    if y is None: y = np.zeros(lay.m)
    if zi is None: zi = np.zeros(lay.q)

    def addW(i, j, v):
        W[i, j] += v
        if i != j:
            W[j, i] += v

    # ---------------- objective (obca.py:859-895 / 1385-1412)
    for k in range(N):
        e = pose(p, lay, X, k) - p.xref[:, k]
        f += e @ p.Q @ e
        if k >= 1:
            i0 = lay.iz(k)
            if need_g: g[i0:i0 + 3] += 2 * p.Q @ e
            if need_W: W[i0:i0 + 3, i0:i0 + 3] += 2 * p.Q
        uk = X[lay.iu(k):lay.iu(k) + 2] - p.uref[:, k]
        f += uk @ p.R1 @ uk
        i0 = lay.iu(k)
        if need_g: g[i0:i0 + 2] += 2 * p.R1 @ uk
        if need_W: W[i0:i0 + 2, i0:i0 + 2] += 2 * p.R1
    e = pose(p, lay, X, N) - p.xref[:, N]
    f += e @ p.P @ e
    i0 = lay.iz(N)
    if need_g: g[i0:i0 + 3] += 2 * p.P @ e
    if need_W: W[i0:i0 + 3, i0:i0 + 3] += 2 * p.P
    Aacc = 0.0
    for k in range(N - 1):
        ia, ib = lay.iu(k), lay.iu(k + 1)
        du = X[ib:ib + 2] - X[ia:ia + 2]
        qv = p.R2 @ du
        Aacc += du @ qv / h ** 2
        if need_g:
            g[ib:ib + 2] += 2 * qv / h ** 2
            g[ia:ia + 2] -= 2 * qv / h ** 2
        if need_W:
            M = 2 * p.R2 / h ** 2
            W[ib:ib + 2, ib:ib + 2] += M; W[ia:ia + 2, ia:ia + 2] += M
            W[ia:ia + 2, ib:ib + 2] -= M; W[ib:ib + 2, ia:ia + 2] -= M
            if free:   # d/dT of 2 q / h^2 = -2/T * (...)
                for j in range(2):
                    addW(iT, ib + j, -4 * qv[j] / (h ** 2 * T))
                    addW(iT, ia + j, +4 * qv[j] / (h ** 2 * T))
    f += Aacc
    if free:
        c1, c2 = p.tcost
        f += (N + 1) * (c1 * T + c2 * T * T)
        if need_g: g[iT] += (N + 1) * (c1 + 2 * c2 * T) - 2 * Aacc / T
        if need_W: W[iT, iT] += (N + 1) * 2 * c2 + 6 * Aacc / T ** 2
    out["f"] = f

    # ---------------- dynamics (obca.py:902-905): c = z_k + h F(z_k,u_k) - z_{k+1}
    for k in range(N):
        zk = pose(p, lay, X, k)
        iu = lay.iu(k)
        v, w = X[iu], X[iu + 1]
        ct, st = np.cos(zk[2]), np.sin(zk[2])
        zn = X[lay.iz(k + 1):lay.iz(k + 1) + 3]
        r = lay.c_dyn + 3 * k
        F = np.array([v * ct, v * st, w])
        c[r:r + 3] = zk + h * F - zn
        if need_J:
            if k >= 1:
                i0 = lay.iz(k)
                J[r, i0] = 1; J[r + 1, i0 + 1] = 1; J[r + 2, i0 + 2] = 1
                J[r, i0 + 2] = -h * v * st
                J[r + 1, i0 + 2] = h * v * ct
            J[r, iu] = h * ct; J[r + 1, iu] = h * st; J[r + 2, iu + 1] = h
            j0 = lay.iz(k + 1)
            J[r, j0] = -1; J[r + 1, j0 + 1] = -1; J[r + 2, j0 + 2] = -1
            if free:
                J[r:r + 3, iT] = p.Ts * F
        if need_W:
            px, py, pt = y[r:r + 3]
            if k >= 1:
                ith = lay.iz(k) + 2
                addW(ith, ith, h * v * (-px * ct - py * st))
                addW(ith, iu, h * (-px * st + py * ct))
                if free:
                    addW(iT, ith, p.Ts * v * (-px * st + py * ct))
            if free:
                addW(iT, iu, p.Ts * (px * ct + py * st))
                addW(iT, iu + 1, p.Ts * pt)

    # ---------------- terminal equality (obca.py:951)
    if free:
        r = lay.c_term
        i0 = lay.iz(N)
        c[r:r + 3] = X[i0:i0 + 3] - p.xref[:, N]
        if need_J:
            for j in range(3):
                J[r + j, i0 + j] = 1

    # ---------------- state bounds (obca.py:916-917), theta unbounded
    for k in range(1, N + 1):
        r = lay.g_xy + 4 * (k - 1); i0 = lay.iz(k)
        for j in range(2):
            d[r + j] = X[i0 + j] - p.xL[j]
            d[r + 2 + j] = p.xU[j] - X[i0 + j]
            if need_J:
                Jd[r + j, i0 + j] = 1; Jd[r + 2 + j, i0 + j] = -1

    # ---------------- input bounds (922-923) and accel rows (928-939)
    for k in range(N):
        iu = lay.iu(k)
        r = lay.g_u + 8 * k
        up = p.u0 if k == 0 else X[lay.iu(k - 1):lay.iu(k - 1) + 2]
        ga = (up - X[iu:iu + 2]) / h
        for j in range(2):
            d[r + j] = X[iu + j] - p.uL[j]
            d[r + 2 + j] = p.uU[j] - X[iu + j]
            d[r + 4 + j] = ga[j] + p.amax[j]
            d[r + 6 + j] = p.amax[j] - ga[j]
            if need_J:
                Jd[r + j, iu + j] = 1; Jd[r + 2 + j, iu + j] = -1
                for sgn, rr in ((1.0, r + 4 + j), (-1.0, r + 6 + j)):
                    Jd[rr, iu + j] = -sgn / h
                    if k >= 1:
                        Jd[rr, lay.iu(k - 1) + j] = sgn / h
                    if free:
                        Jd[rr, iT] = -sgn * ga[j] / T
            if need_W and free:
                yj = -(zi[r + 4 + j] - zi[r + 6 + j])      # W -= z * d2(+-ga)
                addW(iT, iu + j, yj / (h * T))
                if k >= 1:
                    addW(iT, lay.iu(k - 1) + j, -yj / (h * T))
                addW(iT, iT, yj * 2 * ga[j] / T ** 2)

    # ---------------- T bounds (obca.py:959-963)
    if free:
        d[lay.g_T] = T - p.Tmin; d[lay.g_T + 1] = p.Tmax - T
        if need_J:
            Jd[lay.g_T, iT] = 1; Jd[lay.g_T + 1, iT] = -1

    # ---------------- terminal set (obca.py:1465-1466)
    if p.term is not None:
        r = lay.g_term; i0 = lay.iz(N)
        d[r] = X[i0] - p.term[0]; d[r + 1] = X[i0 + 1] - p.term[1]; d[r + 2] = p.term[2] - X[i0 + 1]
        if need_J:
            Jd[r, i0] = 1; Jd[r + 1, i0 + 1] = 1; Jd[r + 2, i0 + 1] = -1

    # ---------------- obstacle rows (obca.py:956-958, 968-1042)
    for k in range(N + 1):
        zk = pose(p, lay, X, k)
        ct, st = np.cos(zk[2]), np.sin(zk[2])
        ip = lay.iz(k) if k >= 1 else -1
        for i in range(p.nobs):
            E = int(p.edges[i]); o = lay.eoff[i]
            Ai = p.A[k, o:o + E]; bi = p.b[k, o:o + E]
            il = lay.lam[k, i]; im = lay.mu[k, i]
            lam = X[il:il + E]; mu = X[im:im + 4]
            a1 = Ai[:, 0] @ lam; a2 = Ai[:, 1] @ lam
            re = lay.c_e[k, i]
            c[re] = mu[0] - mu[2] + ct * a1 + st * a2
            c[re + 1] = mu[1] - mu[3] - st * a1 + ct * a2
            rg = lay.g_w[k, i]
            rn = rg + E + 4; rd = rn + 1
            d[rg:rg + E] = lam
            d[rg + E:rg + E + 4] = mu
            d[rn] = 1.0 - a1 * a1 - a2 * a2
            tx = zk[0] + p.off * ct; ty = zk[1] + p.off * st
            d[rd] = -p.g @ mu + tx * a1 + ty * a2 - bi @ lam - p.dmin
            if need_J:
                J[re, il:il + E] = ct * Ai[:, 0] + st * Ai[:, 1]
                J[re + 1, il:il + E] = -st * Ai[:, 0] + ct * Ai[:, 1]
                J[re, im] = 1; J[re, im + 2] = -1
                J[re + 1, im + 1] = 1; J[re + 1, im + 3] = -1
                for j in range(E):
                    Jd[rg + j, il + j] = 1
                for j in range(4):
                    Jd[rg + E + j, im + j] = 1
                Jd[rn, il:il + E] = -2 * (a1 * Ai[:, 0] + a2 * Ai[:, 1])
                Jd[rd, il:il + E] = tx * Ai[:, 0] + ty * Ai[:, 1] - bi
                Jd[rd, im:im + 4] = -p.g
                if ip >= 0:
                    J[re, ip + 2] = -st * a1 + ct * a2
                    J[re + 1, ip + 2] = -ct * a1 - st * a2
                    Jd[rd, ip] = a1; Jd[rd, ip + 1] = a2
                    Jd[rd, ip + 2] = p.off * (-st * a1 + ct * a2)
            if need_W:
                y1, y2, yn, yd = y[re], y[re + 1], -zi[rn], -zi[rd]
                W[il:il + E, il:il + E] += yn * (-2.0) * (np.outer(Ai[:, 0], Ai[:, 0]) + np.outer(Ai[:, 1], Ai[:, 1]))
                if ip >= 0:
                    ith = ip + 2
                    dA = -st * Ai[:, 0] + ct * Ai[:, 1]     # d/dtheta of (ct A1 + st A2)
                    dB = -ct * Ai[:, 0] - st * Ai[:, 1]     # d/dtheta of (-st A1 + ct A2)
                    col = y1 * dA + y2 * dB + yd * p.off * dA
                    for j in range(E):
                        addW(ith, il + j, col[j])
                        addW(ip, il + j, yd * Ai[j, 0])
                        addW(ip + 1, il + j, yd * Ai[j, 1])
                    addW(ith, ith, y1 * (-ct * a1 - st * a2) + y2 * (st * a1 - ct * a2)
                         + yd * p.off * (-ct * a1 - st * a2))
    out["c"] = c; out["d"] = d
    if need_g: out["g"] = g
    if need_J: out["J"] = J; out["Jd"] = Jd
    if need_W: out["W"] = W
    return out


def sign_rows(p: Problem, lay: Layout):
    """(inequality rows, variable columns) of the sign constraints lambda >= 0, mu >= 0."""
    rows, cols = [], []
    for k in range(p.N + 1):
        for i in range(p.nobs):
            E = int(p.edges[i])
            rows += list(range(lay.g_w[k, i], lay.g_w[k, i] + E + 4))
            cols += list(range(lay.lam[k, i], lay.lam[k, i] + E)) + list(range(lay.mu[k, i], lay.mu[k, i] + 4))
    return np.asarray(rows, int), np.asarray(cols, int)


def unpack(p: Problem, lay: Layout, X):
    """-> x (3,N+1), u (2,N), T, lam (N+1,R), mu (N+1,4*nobs)"""
    N = p.N
    x = np.zeros((3, N + 1)); x[:, 0] = p.x0
    for k in range(1, N + 1):
        x[:, k] = X[lay.iz(k):lay.iz(k) + 3]
    u = X[lay.u:lay.u + 2 * N].reshape(N, 2).T.copy()
    T = X[lay.T] if p.free else 1.0
    lam = np.zeros((N + 1, p.R)); mu = np.zeros((N + 1, 4 * p.nobs))
    for k in range(N + 1):
        for i in range(p.nobs):
            E = int(p.edges[i]); o = lay.eoff[i]
            lam[k, o:o + E] = X[lay.lam[k, i]:lay.lam[k, i] + E]
            mu[k, 4 * i:4 * i + 4] = X[lay.mu[k, i]:lay.mu[k, i] + 4]
    return x, u, T, lam, mu


def pack(p: Problem, lay: Layout, x, u, T, lam, mu):
    """inverse of unpack with the ABI layouts of ONE instance: x (N+1,3), u (N,2), T, lam (N+1,R), mu (N+1,4*nobs) -> X"""
    N = p.N
    X = np.zeros(lay.n)
    x = np.asarray(x, float).reshape(N + 1, 3); u = np.asarray(u, float).reshape(N, 2)
    for k in range(1, N + 1):
        X[lay.iz(k):lay.iz(k) + 3] = x[k]
    X[lay.u:lay.u + 2 * N] = u.reshape(-1)
    if p.free:
        X[lay.T] = float(T)
    for k in range(N + 1):
        for i in range(p.nobs):
            E = int(p.edges[i]); o = lay.eoff[i]
            X[lay.lam[k, i]:lay.lam[k, i] + E] = lam[k, o:o + E]
            X[lay.mu[k, i]:lay.mu[k, i] + 4] = mu[k, 4 * i:4 * i + 4]
    return X


def kkt_certificate(p: Problem, lay: Layout, X, act_tol=1e-3):
    """First-order optimality certificate of a point X, from this restatement alone (no solver state): primal
    infeasibility (max |c|, max violation of d >= 0), and the stationarity residual of the best multipliers the
    active set admits - bounded least squares  min |grad f + J^T y - Jd_A^T z|  s.t. z >= 0  (y free).
    Rows with d <= act_tol enter the fit (generously: a row at distance 1e-4 may still carry a multiplier of 1e-5 at the
    final barrier parameter); `compl` = max z_i d_i over them reports the complementarity of the fitted multipliers.
    -> dict(f, c_max, d_min, stat = max |residual| / max(1, max |grad f|), z_min, compl, n_active)"""
    from scipy.optimize import lsq_linear
    ev = evaluate(p, lay, X, want=("f", "g", "c", "J", "d", "Jd"))
    act = ev["d"] <= act_tol
    J = ev["J"]; JdA = ev["Jd"][act]
    J = J.toarray() if hasattr(J, "toarray") else np.asarray(J)
    JdA = JdA.toarray() if hasattr(JdA, "toarray") else np.asarray(JdA)
    M = np.hstack([J.T, -JdA.T])
    lb = np.concatenate([np.full(J.shape[0], -np.inf), np.zeros(JdA.shape[0])])
    r = lsq_linear(M, -ev["g"], bounds=(lb, np.full(M.shape[1], np.inf)), method="bvls" if M.shape[1] < 400 else "trf",
                   tol=1e-13, max_iter=400)
    res = M @ r.x + ev["g"]
    z = r.x[J.shape[0]:]
    return dict(f=float(ev["f"]), c_max=float(np.abs(ev["c"]).max()) if len(ev["c"]) else 0.0, d_min=float(ev["d"].min()),
                stat=float(np.abs(res).max() / max(1.0, np.abs(ev["g"]).max())), z_min=float(z.min()) if len(z) else 0.0,
                compl=float((z * np.maximum(ev["d"][act], 0.0)).max()) if len(z) else 0.0, n_active=int(act.sum()))


def objective_of(p: Problem, x, u, T):
    """Objective exactly as the reference's obj_rule writes it (obca.py:859-895), from (x,u,T) arrays."""
    N = p.N
    h = (T if p.free else 1.0) * p.Ts
    f = 0.0
    for t in range(N):
        e = x[:, t] - p.xref[:, t]
        f += e @ p.Q @ e
        uu = u[:, t] - p.uref[:, t]
        f += uu @ p.R1 @ uu
        if t < N - 1:
            du = (u[:, t + 1] - u[:, t]) / h
            f += du @ p.R2 @ du
    e = x[:, N] - p.xref[:, N]
    f += e @ p.P @ e
    if p.free:
        f += (N + 1) * (p.tcost[0] * T + p.tcost[1] * T ** 2)
    return f

"""ctypes mirror of ``include/obca_b200.h`` and the host-side packing of the reference's arguments.

``pack_problem`` turns the positional arguments the reference's ``closedLoop`` hands to
``obca.obca_mpc4 / obca_mpc6 / obca_mpc8 / obca2`` (/root/reference/src/closed_loop.py:118,131,137,170)
into the batch-major float64 arrays of the C-ABI.  No solver code lives here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MODE_FREE = 0          # obca_mpc4            obca.py:828-1071
MODE_FIXED_SET = 1     # obca_mpc6            obca.py:1361-1562
MODE_FIXED_NOTERM = 2  # obca_mpc8            obca.py:1564-1758
MODE_FREE_STACKED = 3  # obca2, fixtime == 0  obca.py:338-629
MODE_FIXED_OBCA2 = 4   # obca2, fixtime == 1  (terminal_set optional)

INIT_ZERO, INIT_XREF, INIT_WARM = 0, 1, 2
INIT_GUESS = 4                        # OBCA_INIT_GUESS: poses of the start point are read from the output array x
INIT_RETRY = 16                       # OBCA_INIT_RETRY: failed attempts restart from the other start points
INIT_NORESTO = 32                     # OBCA_INIT_NORESTO: no feasibility-restoration phase
INIT_PATIENT = 64                     # OBCA_INIT_PATIENT: the iteration budget counts per start point
RECOVER = INIT_RETRY                  # what the receding-horizon drivers use: restoration phase, then the other start points


def init_soft(n):
    """OBCA_INIT_SOFT(n): up to n soft restarts (multipliers, slacks, barrier parameter, filter) per attempt"""
    return (int(n) & 15) << 8

ST_OK, ST_ACCEPTABLE, ST_MAXITER, ST_REGFAIL, ST_EMPTYBOX, ST_LSFAIL, ST_STALL = 0, 1, -1, -2, -3, -4, -5
ST_INFEASIBLE, ST_RESTOFAIL = -6, -7
ST_FLOOR = 2                          # ended on the rounding-noise floor (see include/obca_b200.h)

MAX_STAGES, MAX_OBS, MAX_ROWS = 32, 12, 48


class ObcaParams(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("N", C.c_int32), ("n_obs", C.c_int32), ("rows", C.c_int32),
        ("init", C.c_int32), ("max_iter", C.c_int32), ("has_term", C.c_int32), ("acceptable_iter", C.c_int32),
        ("Ts", C.c_double), ("dmin", C.c_double), ("ego", C.c_double * 4),
        ("Q", C.c_double * 9), ("P", C.c_double * 9), ("R1", C.c_double * 4), ("R2", C.c_double * 4),
        ("xL", C.c_double * 2), ("xU", C.c_double * 2), ("uL", C.c_double * 2), ("uU", C.c_double * 2),
        ("acc_max", C.c_double * 2), ("time_cost", C.c_double * 2), ("T_min", C.c_double),
        ("tol", C.c_double), ("acceptable_tol", C.c_double), ("mu_init", C.c_double), ("bound_push", C.c_double),
    ]


class LoopParams(C.Structure):
    """obca_loop_params of include/obca_b200.h"""
    _fields_ = [
        ("N", C.c_int32), ("max_steps", C.c_int32), ("n_static", C.c_int32), ("rows_static", C.c_int32),
        ("path_len", C.c_int32), ("terminal_rule", C.c_int32),
        ("sense", C.c_double), ("goal", C.c_double * 2), ("goal_tol", C.c_double), ("start", C.c_double * 3),
        ("Ts0", C.c_double), ("speculative", C.c_int32), ("reserved", C.c_int32),
    ]


def is_free(mode):
    return mode in (MODE_FREE, MODE_FREE_STACKED)


def make_params(mode, N, n_obs, rows, Ts, P, Q, R, xL, xU, uL, uU, dmin, ego, *, init=INIT_WARM, has_term=None,
                max_iter=None, tol=1e-8, acceptable_tol=None, acceptable_iter=15, mu_init=10.0, bound_push=0.1, T_min=1e-4,
                soft_restarts=0, retry=False):
    """Solver options default to what the reference passes to IPOPT: mpc4 -> IPOPT defaults (max_iter 3000,
    acceptable_tol 1e-6; obca.py:1044); mpc6/mpc8/obca2-fixed -> max_iter 1000, acceptable_tol 1e-8
    (obca.py:1538-1539, 1734-1735, 596-598).  ``soft_restarts`` / ``retry`` switch on the recovery rules that stand
    in for IPOPT's restoration phase (OBCA_INIT_SOFT / OBCA_INIT_RETRY in include/obca_b200.h); they can also be
    OR-ed into ``init`` directly (``init=INIT_WARM | RECOVER``)."""
    if N + 1 > MAX_STAGES or N < 1:
        raise ValueError("horizon N=%d outside 1..%d" % (N, MAX_STAGES - 1))
    if n_obs > MAX_OBS or rows > MAX_ROWS:
        raise ValueError("too many obstacles/rows for one stage (%d obstacles, %d rows)" % (n_obs, rows))
    free = is_free(mode)
    p = ObcaParams()
    p.mode, p.N, p.n_obs, p.rows = mode, N, n_obs, rows
    p.init = int(init) | init_soft(soft_restarts) | (INIT_RETRY if retry else 0)
    p.max_iter = (3000 if free else 1000) if max_iter is None else max_iter
    p.has_term = int(mode == MODE_FIXED_SET) if has_term is None else int(has_term)
    p.acceptable_iter = acceptable_iter
    p.Ts, p.dmin = float(Ts), float(dmin)
    p.ego[:] = [float(e) for e in np.asarray(ego, float).reshape(4)]
    p.Q[:] = np.asarray(Q, float).reshape(9).tolist()
    p.P[:] = np.asarray(P, float).reshape(9).tolist()
    p.R1[:] = np.asarray(R[0], float).reshape(4).tolist()
    p.R2[:] = np.asarray(R[1], float).reshape(4).tolist()
    p.xL[:] = np.asarray(xL, float).reshape(-1)[:2].tolist()
    p.xU[:] = np.asarray(xU, float).reshape(-1)[:2].tolist()
    p.uL[:] = np.asarray(uL, float).reshape(2).tolist()
    p.uU[:] = np.asarray(uU, float).reshape(2).tolist()
    p.acc_max[:] = [0.6, float(np.pi / 6)]          # obca.py:932-933
    p.time_cost[:] = [10.0, 1.0]                     # obca.py:888
    p.T_min = T_min                                  # 1e-4: obca.py:963
    p.tol = tol
    p.acceptable_tol = (1e-6 if free else 1e-8) if acceptable_tol is None else acceptable_tol
    p.mu_init, p.bound_push = mu_init, bound_push
    return p


def pack_obstacles(mode, N, nObs, vObs, AObs, bObs, atol=1e-9):
    """Time-stacked (AObs ((N+1)R, 2), bObs ((N+1)R, 1)) of closed_loop.py:488-500 -> (edge_ptr, A (R,2),
    b0 (R,), db (R,) or None).  ``rebuild_lObs`` only translates polygons (demo_setting.py:457-473), so A is
    constant in k and b is affine in k: b_k = b0 + k*db.  mpc4 reads the first block only (obca.py:969).
    A single block ((R,2)) is accepted too (static scene)."""
    edges = [int(v) - 1 for v in list(vObs)[:nObs]]
    R = int(sum(edges))
    edge_ptr = np.concatenate([[0], np.cumsum(edges)]).astype(np.int32)
    AObs = np.ascontiguousarray(AObs, dtype=np.float64).reshape(-1, 2)
    bObs = np.ascontiguousarray(bObs, dtype=np.float64).reshape(-1)
    if R == 0:
        return edge_ptr, np.zeros((0, 2)), np.zeros(0), None
    nblk = AObs.shape[0] // R
    if AObs.shape[0] % R or bObs.shape[0] != AObs.shape[0] or nblk < 1:
        raise ValueError("AObs/bObs rows (%d/%d) are not a multiple of R=%d" % (AObs.shape[0], bObs.shape[0], R))
    A = AObs[:R].copy(); b0 = bObs[:R].copy()
    if mode == MODE_FREE or nblk == 1:
        return edge_ptr, A, b0, None
    if nblk < N + 1:
        raise ValueError("time-stacked obstacle rows cover %d steps, horizon needs %d" % (nblk, N + 1))
    Ak = AObs[:(N + 1) * R].reshape(N + 1, R, 2); bk = bObs[:(N + 1) * R].reshape(N + 1, R)
    db = (bk[N] - bk[0]) / N
    scale = 1.0 + np.abs(bk).max()
    if np.abs(Ak - A).max() > atol * (1 + np.abs(A).max()) or \
            np.abs(bk - (b0 + np.arange(N + 1)[:, None] * db)).max() > atol * scale:
        raise ValueError("obstacle rows are not a pure constant-velocity translation of the first time block")
    if not np.any(db):
        db = None
    return edge_ptr, A, b0, db


def tmax_of(xref_N, x0, N, uU0, Ts):
    """obca.py:961-962 (signed sum of dx + dy)."""
    return ((xref_N[0] - x0[0]) + (xref_N[1] - x0[1])) / (N * uU0 * Ts) + 1.0


def term_of(terminal_set):
    """terminal_set [[xmin, _], [ymin, ymax]] -> (xmin, ymin, ymax)  (obca.py:1465-1466)."""
    ts = np.asarray(terminal_set, float)
    return np.array([ts[0, 0], ts[1, 0], ts[1, 1]])


def ptr(a, ctype=C.c_double):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


SOLVE_ARGTYPES_HOST = [C.c_int] + [C.POINTER(C.c_double)] * 7 + [C.POINTER(C.c_int32)] + [C.POINTER(C.c_double)] * 3 + \
    [C.c_int] + [C.POINTER(C.c_double)] * 6 + [C.POINTER(C.c_int32)] * 2

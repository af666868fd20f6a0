"""Drop-in for the reference's ``class obca`` (/root/reference/src/obca.py:10) backed by the B200 kernel.

Call surface kept (positional signatures and the 4-tuple return ``(x_Opt (3,N+1), u_Opt (2,N), feas, Ts_opt)``):

    obca().obca_mpc4(Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0)
                                                              obca.py:828-1071, closed_loop.py:118,382
    obca().obca_mpc6(..., u0, uOpt, terminal_set)             obca.py:1361-1562, closed_loop.py:131,269,389
    obca().obca_mpc8(..., u0, uOpt)                           obca.py:1564-1758, closed_loop.py:137,275,395
    obca().obca2(Ts, P, Q, R, N, x0, u0, xL, xU, uL, uU, xref, uref, nObs, vObs, AObs, bObs, dmin, ego,
                 fixtime, timeScale_size, terminal_set)       obca.py:338-629,  closed_loop.py:170,263
    obca().obca(...same without terminal_set)                 obca.py:12-336 (dead code in the reference)

Like the reference the methods never raise on a solver failure: ``feas`` is False and the last iterate is
returned (obca.py:1062-1065).  The duals, objective, status and iteration count the reference does not return
are kept on the instance (``.lam``, ``.mu``, ``.obj``, ``.status``, ``.iters``).

``BatchSolver`` / ``solve_batch`` is the batched entry (thousands of independent NLPs in one launch); the
single-problem methods are a batch of one through the same C-ABI (host buffers, copies included).
There is no CPU path: without the CUDA library or a device these calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi, _lib
from ._abi import (INIT_WARM, INIT_XREF, INIT_ZERO, MODE_FIXED_NOTERM, MODE_FIXED_OBCA2, MODE_FIXED_SET, MODE_FREE,  # noqa: F401
                   MODE_FREE_STACKED)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class BatchSolver:
    """One solver context = one (device, parameter set).  ``solve`` takes device tensors (torch, CUDA,
    float64, contiguous) and is asynchronous on the current torch stream; ``solve_host`` takes NumPy
    arrays and includes the host<->device copies.  Shapes follow include/obca_b200.h."""

    def __init__(self, params: _abi.ObcaParams, edge_ptr, max_batch, device=-1):
        self.params = params
        self.edge_ptr = np.ascontiguousarray(edge_ptr, dtype=np.int32)
        if self.edge_ptr.shape[0] != params.n_obs + 1 or int(self.edge_ptr[-1]) != params.rows:
            raise ValueError("edge_ptr does not match n_obs / rows")
        self.max_batch = int(max_batch)
        self._L = _lib.lib()
        self._ctx = C.c_void_p()
        _lib.check(self._L.obca_b200_create(C.byref(self._ctx), int(device), self.max_batch, C.byref(params)))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.obca_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self._L.obca_b200_launch_count(self._ctx))

    @property
    def bulk_timeouts(self):
        """input prefetches (cp.async.bulk) that timed out inside the kernel - diagnostics, expected 0"""
        return int(self._L.obca_b200_bulk_timeouts(self._ctx))

    @property
    def scratch_bytes(self):
        return int(self._L.obca_b200_scratch_bytes(self._ctx))

    def last_kernel_ms(self):
        return float(self._L.obca_b200_last_kernel_ms(self._ctx))

    # ---- device path -------------------------------------------------------------------------------
    def alloc_outputs(self, B, device):
        import torch
        p = self.params
        N, R, no = p.N, p.rows, p.n_obs
        kw = dict(dtype=torch.float64, device=device)
        return dict(x=torch.empty((B, N + 1, 3), **kw), u=torch.empty((B, N, 2), **kw),
                    lam=torch.empty((B, N + 1, R), **kw), mu=torch.empty((B, N + 1, 4 * no), **kw),
                    T=torch.empty((B,), **kw), obj=torch.empty((B,), **kw),
                    status=torch.empty((B,), dtype=torch.int32, device=device),
                    iters=torch.empty((B,), dtype=torch.int32, device=device))

    def solve(self, x0, u0, xref, A, b0, db=None, T_max=None, term=None, uref=None, out=None, stream=None, Ts=None,
              index=None, count=None):
        """``index`` (int32 CUDA tensor) and ``count`` (int32 CUDA tensor with one element): solve only the instances
        ``index[:count]``; arrays stay indexed by instance (``obca_b200_solve_indexed``)."""
        import torch

        def dp(t, dtype=torch.float64):
            if t is None:
                return None
            if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
                raise ValueError("device tensors must be CUDA, contiguous, %s" % dtype)
            return t.data_ptr()
        B = int(x0.shape[0])
        if out is None:
            out = self.alloc_outputs(B, x0.device)
        shared = int(A.dim() == 2)
        if stream is None:
            stream = torch.cuda.current_stream(x0.device).cuda_stream
        if (index is None) != (count is None):
            raise ValueError("index and count go together")
        rc = self._L.obca_b200_solve_indexed(self._ctx, B, dp(count, torch.int32), dp(index, torch.int32),
                                             dp(x0), dp(u0), dp(xref), dp(uref), dp(T_max), dp(term), dp(Ts),
                                             _abi.ptr(self.edge_ptr, C.c_int32), dp(A), dp(b0), dp(db), shared,
                                             dp(out["x"]), dp(out["u"]), dp(out["lam"]), dp(out["mu"]), dp(out["T"]),
                                             dp(out["obj"]), dp(out["status"], torch.int32),
                                             dp(out["iters"], torch.int32), C.c_void_p(stream))
        _lib.check(rc)
        return out

    # ---- host path ---------------------------------------------------------------------------------
    def alloc_host_outputs(self, B, pinned=False):
        """Host result arrays; ``pinned=True`` backs them with page-locked memory (torch) so the D2H copies of
        ``solve_host`` are true DMA transfers."""
        p = self.params
        N, R, no = p.N, p.rows, p.n_obs
        shapes = dict(x=(B, N + 1, 3), u=(B, N, 2), lam=(B, N + 1, R), mu=(B, N + 1, 4 * no), T=(B,), obj=(B,))
        if pinned:
            import torch
            out = {k: torch.empty(s, dtype=torch.float64).pin_memory().numpy() for k, s in shapes.items()}
            out["status"] = torch.empty((B,), dtype=torch.int32).pin_memory().numpy()
            out["iters"] = torch.empty((B,), dtype=torch.int32).pin_memory().numpy()
            return out
        out = {k: np.empty(s) for k, s in shapes.items()}
        out["status"] = np.empty(B, np.int32); out["iters"] = np.empty(B, np.int32)
        return out

    def solve_host(self, x0, u0, xref, A, b0, db=None, T_max=None, term=None, uref=None, out=None, Ts=None):
        p = self.params
        N, R, no = p.N, p.rows, p.n_obs
        x0, u0, xref, uref, T_max, term, A, b0, db, Ts = map(_f64, (x0, u0, xref, uref, T_max, term, A, b0, db, Ts))
        B = x0.shape[0]
        if xref.shape != (B, N + 1, 3) or u0.shape != (B, 2) or x0.shape != (B, 3):
            raise ValueError("x0 (B,3), u0 (B,2), xref (B,N+1,3) expected")
        shared = int(A.ndim == 2)
        if out is None:
            out = self.alloc_host_outputs(B)
        elif out["x"].shape != (B, N + 1, 3) or out["lam"].shape != (B, N + 1, R) or out["mu"].shape != (B, N + 1, 4 * no):
            raise ValueError("preallocated outputs do not match the batch")
        P = _abi.ptr
        rc = self._L.obca_b200_solve_host(self._ctx, B, P(x0), P(u0), P(xref), P(uref), P(T_max), P(term), P(Ts),
                                          P(self.edge_ptr, C.c_int32), P(A), P(b0), P(db), shared,
                                          P(out["x"]), P(out["u"]), P(out["lam"]), P(out["mu"]), P(out["T"]),
                                          P(out["obj"]), P(out["status"], C.c_int32), P(out["iters"], C.c_int32))
        _lib.check(rc)
        return out


def solve_batch(mode, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0,
                terminal_set=None, uref=None, init=INIT_WARM, device=-1, T_max=None, **opts):
    """Batched form of the reference call: x0 (B,3), u0 (B,2), xref (B,3,N+1) [the reference's layout],
    terminal_set (B,2,2) or None; one scene (AObs, bObs) shared by the batch.  Host arrays in, host arrays
    out (x (B,3,N+1), u (B,2,N), feas (B,), Ts_opt (B,), plus lam/mu/obj/status/iters)."""
    x0 = np.asarray(x0, float).reshape(-1, 3)
    B = x0.shape[0]
    u0 = np.asarray(u0, float).reshape(-1, 2)
    if u0.shape[0] == 1 and B > 1:
        u0 = np.repeat(u0, B, axis=0)
    xref = np.asarray(xref, float).reshape(B, 3, N + 1)
    edge_ptr, A, b0, db = _abi.pack_obstacles(mode, N, nObs, vObs, AObs, bObs)
    has_term = terminal_set is not None and np.size(terminal_set) > 0
    if mode == MODE_FIXED_SET and not has_term:
        raise ValueError("obca_mpc6 needs a terminal_set")
    prm = _abi.make_params(mode, N, nObs, int(edge_ptr[-1]), Ts, P, Q, R, xL, xU, uL, uU, dmin, ego, init=init,
                           has_term=has_term and mode in (MODE_FIXED_SET, MODE_FIXED_OBCA2), **opts)
    term = None
    if T_max is not None:
        T_max = np.broadcast_to(np.asarray(T_max, float).reshape(-1), (B,)).copy()
    elif _abi.is_free(mode):
        uU0 = float(np.asarray(uU, float).reshape(-1)[0])
        T_max = ((xref[:, 0, N] - x0[:, 0]) + (xref[:, 1, N] - x0[:, 1])) / (N * uU0 * Ts) + 1.0   # obca.py:961-962
    if prm.has_term:
        ts = np.asarray(terminal_set, float).reshape(-1, 2, 2)
        if ts.shape[0] == 1 and B > 1:
            ts = np.repeat(ts, B, axis=0)
        term = np.stack([ts[:, 0, 0], ts[:, 1, 0], ts[:, 1, 1]], axis=1)
    if uref is not None and np.size(uref):
        uref = np.ascontiguousarray(np.asarray(uref, float).reshape(B, 2, N).transpose(0, 2, 1))
    else:
        uref = None
    s = BatchSolver(prm, edge_ptr, B, device)
    try:
        o = s.solve_host(x0, u0, np.ascontiguousarray(xref.transpose(0, 2, 1)), A, b0, db, T_max=T_max, term=term, uref=uref)
    finally:
        s.close()
    return dict(x=np.ascontiguousarray(o["x"].transpose(0, 2, 1)), u=np.ascontiguousarray(o["u"].transpose(0, 2, 1)),
                feas=o["status"] >= 0, Ts_opt=(o["T"] * Ts if _abi.is_free(mode) else np.full(B, float(Ts))),
                T=o["T"], lam=o["lam"], mu=o["mu"], obj=o["obj"], status=o["status"], iters=o["iters"])


class obca:
    """Same name, same methods, same returns as the reference's solver object (no constructor arguments,
    closed_loop.py:22)."""

    # start point: None = automatic (A* warm start when xref is a path window; the reference's own IPOPT start -
    # every variable 0, Topt = 1, obca.py:856 - when xref is the degenerate start/goal-only reference of
    # closed_loop.py:535-544, whose poses are useless as an initial trajectory), or force INIT_ZERO / INIT_XREF / INIT_WARM
    init = None
    device = -1
    # what follows a failed attempt (the solver's own feasibility-restoration phase comes first): the other start points,
    # with the iteration budget counted per start point - a single solve has no batch to hold up.  OR-ed into the
    # start point; 0 switches it off.
    recover = _abi.INIT_RETRY | _abi.INIT_PATIENT
    # IPOPT's defaults, which the reference runs with: a single solve takes ~30 % more iterations with
    # them than with the batch defaults (10, 0.1) but the hard open-loop problems of closed_loop.py:113-120 need them
    mu_init = 0.1
    bound_push = 1e-2

    @staticmethod
    def _auto_init(xref, x0, N):
        P = np.concatenate([np.asarray(x0, float).reshape(3, 1)[:2], np.asarray(xref, float).reshape(3, N + 1)[:2, 1:]], axis=1)
        seg = np.sqrt((np.diff(P, axis=1) ** 2).sum(0))
        tot = seg.sum()
        if N > 2 and tot > 0 and seg.max() > 0.5 * tot:
            return INIT_ZERO
        return INIT_WARM

    def _one(self, mode, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0,
             terminal_set=None, uref=None, **opts):
        r = solve_batch(mode, Ts, P, Q, R, int(N), np.asarray(x0, float).reshape(1, 3), xL, xU, uL, uU,
                        np.asarray(xref, float).reshape(1, 3, int(N) + 1), int(nObs), vObs, AObs, bObs, dmin, ego,
                        np.asarray(u0, float).reshape(1, 2), terminal_set=terminal_set, uref=uref,
                        init=(self._auto_init(xref, x0, int(N)) if self.init is None else self.init) | self.recover,
                        device=self.device, **dict(dict(mu_init=self.mu_init, bound_push=self.bound_push), **opts))
        self.lam, self.mu = r["lam"][0], r["mu"][0]
        self.obj, self.status, self.iters, self.T = float(r["obj"][0]), int(r["status"][0]), int(r["iters"][0]), float(r["T"][0])
        return r["x"][0], r["u"][0], bool(r["feas"][0]), float(r["Ts_opt"][0])

    def obca_mpc4(self, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0):
        return self._one(MODE_FREE, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0)

    def obca_mpc6(self, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0, uOpt, terminal_set):
        # uOpt is accepted and ignored, as in the reference (obca.py:1402)
        return self._one(MODE_FIXED_SET, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego,
                         u0, terminal_set=terminal_set)

    def obca_mpc8(self, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego, u0, uOpt):
        return self._one(MODE_FIXED_NOTERM, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin,
                         ego, u0)

    def obca2(self, Ts, P, Q, R, N, x0, u0, xL, xU, uL, uU, xref, uref, nObs, vObs, AObs, bObs, dmin, ego, fixtime,
              timeScale_size, terminal_set):
        # timeScale_size is ignored by the reference's obca2
        if fixtime == 0:
            return self._one(MODE_FREE_STACKED, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs,
                             dmin, ego, u0, uref=uref)
        ts = terminal_set if (terminal_set is not None and np.size(terminal_set) > 0) else None
        return self._one(MODE_FIXED_OBCA2, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin,
                         ego, u0, terminal_set=ts, uref=uref)

    def obca(self, Ts, P, Q, R, N, x0, u0, xL, xU, uL, uU, xref, uref, nObs, vObs, AObs, bObs, dmin, ego, fixtime,
             timeScale_size):
        """The earliest variant (obca.py:12-336; no caller anywhere in the reference).  Free time only: obca2's NLP with the
        time-scale box of obca.py:234-240 - 'big': [1e-4, sqrt(dx^2 + dy)/(N uU Ts) + 1] (the missing square on dy is the
        reference's), 'small': [0.8, 1.2].  Its fixed-time form (terminal xy equality with a +-pi/4 heading
        band, obca.py:223-225) has no counterpart in the live code paths and is not implemented."""
        if fixtime != 0:
            raise NotImplementedError("obca.obca(fixtime=1) is dead code in the reference; use obca_mpc6 / obca_mpc8 / obca2")
        x0a = np.asarray(x0, float).reshape(3); xr = np.asarray(xref, float).reshape(3, int(N) + 1)
        if timeScale_size == 'small':
            opts = dict(T_min=0.8, T_max=1.2)
        else:
            uU0 = float(np.asarray(uU, float).reshape(-1)[0])
            dis = np.sqrt((xr[0, N] - x0a[0]) ** 2 + (xr[1, N] - x0a[1]))
            opts = dict(T_min=1e-4, T_max=dis / (N * uU0 * Ts) + 1)
        return self._one(MODE_FREE_STACKED, Ts, P, Q, R, N, x0, xL, xU, uL, uU, xref, nObs, vObs, AObs, bObs, dmin, ego,
                         u0, uref=uref, **opts)


OBCA = obca

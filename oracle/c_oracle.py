"""ORACLE (test infrastructure): build + ctypes binding of ``oracle/obca_oracle.c``.

Only tests/, __graft_entry__ (build/smoke) and bench.py's cpu_baseline / --impl reference legs use this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "obca_oracle.c")
LIB = os.path.join(HERE, "libobca_oracle.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
            os.path.getmtime(SRC), os.path.getmtime(os.path.join(HERE, "..", "include", "obca_b200.h"))):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, SRC, "-lm", "-lpthread"])
    return LIB


_lib = None
TRACE_FN = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.obca_oracle_solve.restype = C.c_int
        _lib.obca_oracle_solve.argtypes = [C.POINTER(_abi.ObcaParams)] + _abi.SOLVE_ARGTYPES_HOST + [C.c_int]
        _lib.obca_oracle_set_trace.argtypes = [TRACE_FN]
    return _lib


def solve(params, x0, u0, xref, edge_ptr, A, b0, db=None, T_max=None, term=None, uref=None, nthreads=1, trace=None,
          Ts=None, guess=None):
    """ABI-level arrays (see include/obca_b200.h) -> dict of outputs.  x0 (B,3), u0 (B,2), xref (B,N+1,3)."""
    L = lib()
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    x0, u0, xref, uref, T_max, term, A, b0, db, Ts = map(f64, (x0, u0, xref, uref, T_max, term, A, b0, db, Ts))
    B = x0.shape[0]
    N, R, no = params.N, params.rows, params.n_obs
    shared = int(A.ndim == 2)
    out = dict(x=np.zeros((B, N + 1, 3)) if guess is None else np.array(guess, dtype=np.float64).reshape(B, N + 1, 3),
               u=np.zeros((B, N, 2)), lam=np.zeros((B, N + 1, R)),
               mu=np.zeros((B, N + 1, 4 * no)), T=np.zeros(B), obj=np.zeros(B),
               status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
    ep = np.ascontiguousarray(edge_ptr, dtype=np.int32)
    cb = TRACE_FN(trace) if trace else C.cast(None, TRACE_FN)
    L.obca_oracle_set_trace(cb)
    rc = L.obca_oracle_solve(C.byref(params), B, _abi.ptr(x0), _abi.ptr(u0), _abi.ptr(xref), _abi.ptr(uref),
                             _abi.ptr(T_max), _abi.ptr(term), _abi.ptr(Ts), _abi.ptr(ep, C.c_int32), _abi.ptr(A), _abi.ptr(b0),
                             _abi.ptr(db), shared, _abi.ptr(out["x"]), _abi.ptr(out["u"]), _abi.ptr(out["lam"]),
                             _abi.ptr(out["mu"]), _abi.ptr(out["T"]), _abi.ptr(out["obj"]),
                             _abi.ptr(out["status"], C.c_int32), _abi.ptr(out["iters"], C.c_int32), nthreads)
    L.obca_oracle_set_trace(C.cast(None, TRACE_FN))
    if rc != 0:
        raise RuntimeError("obca_oracle_solve rc=%d" % rc)
    return out

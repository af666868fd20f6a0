"""Build + load the CUDA shared library (csrc/libobca_b200.so, C-ABI of include/obca_b200.h).

The library is built in-tree with nvcc for sm_100a.  There is no CPU fallback: if it cannot be built or
loaded, or no CUDA device is present at solve time, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("OBCA_B200_LIB") or os.path.join(CSRC, "libobca_b200.so")   # env override: developer builds
SOURCES = ["obca_b200.cu", "obca_loop.cu", "obca_planner.cpp"]
HEADERS = ["obca_cta.cuh", "obca_kernel.cuh", os.path.join("..", "..", "include", "obca_b200.h")]
# Kernel variants (obca_variant.cu compiled once per row, one translation unit each so that they build in parallel):
# symbol, max edges per obstacle, threads per block, blocks per SM, then horizon / obstacles / rows compiled in (0 = generic)
VARIANTS = [("obca_kv_cfg3", 4, 128, 3, 20, 4, 16), ("obca_kv_cfg5", 4, 192, 2, 20, 6, 24), ("obca_kv_cfg2", 4, 128, 3, 10, 2, 8),
            ("obca_kv_cfg4d", 4, 128, 3, 5, 6, 18), ("obca_kv_cfg4f", 4, 128, 3, 5, 5, 14),
            ("obca_kv_g4_128", 4, 128, 3, 0, 0, 0), ("obca_kv_g4_192", 4, 192, 2, 0, 0, 0), ("obca_kv_g4_416", 4, 416, 1, 0, 0, 0),
            ("obca_kv_g8_128", 8, 128, 2, 0, 0, 0), ("obca_kv_g8_416", 8, 416, 1, 0, 0, 0),
            # recovery kernels (the complete sequence: pass, restoration phase, fresh starts, other start points)
            ("obca_kv_r4_128", 4, 128, 3, 0, 0, 0), ("obca_kv_r4_192", 4, 192, 2, 0, 0, 0), ("obca_kv_r4_416", 4, 416, 1, 0, 0, 0),
            ("obca_kv_r8_128", 8, 128, 2, 0, 0, 0), ("obca_kv_r8_416", 8, 416, 1, 0, 0, 0)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-pthread"]


def _mtime(f):
    f = os.path.join(CSRC, f)
    return os.path.getmtime(f) if os.path.exists(f) else 0.0


def _stale():
    if os.environ.get("OBCA_B200_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_mtime(f) > t for f in SOURCES + HEADERS + ["obca_variant.cu"])


def build(force=False, verbose=False, extra_flags=(), out=None, objdir=None):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libobca_b200.so (cross-compiles without a GPU).

    Every kernel variant and every source file is its own object (csrc/_obj/), compiled in parallel on the host's
    cores and re-compiled only when a file it includes is newer; the objects are then linked into the library."""
    out = out or LIB
    if not force and out == LIB and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = objdir or os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
    newest_hdr = max(_mtime(f) for f in HEADERS)
    jobs = []   # (object, source, extra defines)
    for sym, e, t, b, n, o, r in VARIANTS:
        jobs.append((os.path.join(objdir, sym + ".o"), "obca_variant.cu",
                     ["-DKV_SYM=%s" % sym, "-DKV_E=%d" % e, "-DKV_T=%d" % t, "-DKV_B=%d" % b, "-DKV_N=%d" % n, "-DKV_O=%d" % o,
                      "-DKV_R=%d" % r, "-DKV_FULL=%d" % int(sym.startswith("obca_kv_r"))]))
    for src in SOURCES:
        jobs.append((os.path.join(objdir, os.path.splitext(src)[0] + ".o"), src, []))

    def compile_one(job):
        obj, src, defs = job
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(newest_hdr, _mtime(src)):
            return ""
        r = subprocess.run([nvcc] + flags + defs + ["-c", "-o", obj, src], cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s %s:\n%s" % (src, " ".join(defs), r.stderr[-4000:]))
        return r.stderr

    jobs.sort(key=lambda j: 0 if "KV_FULL=1" in " ".join(j[2]) else 1)   # the recovery kernels take longest: start them first
    workers = int(os.environ.get("OBCA_BUILD_JOBS", "0")) or max(1, (os.cpu_count() or 1))
    with ThreadPoolExecutor(workers) as pool:
        logs = list(pool.map(compile_one, jobs))
    if verbose:
        print("\n".join(l for l in logs if l))
    subprocess.check_call([nvcc, "-shared", "-o", out] + [j[0] for j in jobs] + ["-lpthread"], cwd=CSRC)
    return out


_lib = None
EXPORTS = ["obca_b200_bulk_timeouts", "obca_b200_fp64_peak", "obca_b200_abi_version", "obca_b200_create", "obca_b200_destroy", "obca_b200_scratch_bytes",
           "obca_b200_solve", "obca_b200_solve_host", "obca_b200_launch_count", "obca_b200_last_kernel_ms",
           "obca_b200_strerror", "obca_b200_astar_batch", "obca_b200_reference_windows",
           "obca_b200_solve_indexed", "obca_b200_build_rows", "obca_b200_loop_create", "obca_b200_loop_reset",
           "obca_b200_loop_run", "obca_b200_loop_read", "obca_b200_loop_launch_count", "obca_b200_loop_destroy"]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        try:
            build()
        except Exception as e:  # no nvcc on this box: a prebuilt .so is still fine if present
            if not os.path.exists(LIB):
                raise RuntimeError("libobca_b200.so is missing and could not be built: %r" % (e,))
    L = C.CDLL(LIB)
    L.obca_b200_abi_version.restype = C.c_int
    L.obca_b200_create.restype = C.c_int
    L.obca_b200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(_abi.ObcaParams)]
    L.obca_b200_destroy.restype = C.c_int
    L.obca_b200_destroy.argtypes = [C.c_void_p]
    L.obca_b200_scratch_bytes.restype = C.c_int64
    L.obca_b200_scratch_bytes.argtypes = [C.c_void_p]
    L.obca_b200_launch_count.restype = C.c_int64
    L.obca_b200_launch_count.argtypes = [C.c_void_p]
    L.obca_b200_last_kernel_ms.restype = C.c_float
    L.obca_b200_last_kernel_ms.argtypes = [C.c_void_p]
    L.obca_b200_bulk_timeouts.restype = C.c_int64
    L.obca_b200_bulk_timeouts.argtypes = [C.c_void_p]
    L.obca_b200_fp64_peak.restype = C.c_int
    L.obca_b200_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.obca_b200_strerror.restype = C.c_char_p
    L.obca_b200_strerror.argtypes = [C.c_int]
    vp = C.c_void_p
    L.obca_b200_solve.restype = C.c_int
    L.obca_b200_solve.argtypes = [vp, C.c_int] + [vp] * 7 + [C.POINTER(C.c_int32)] + [vp] * 3 + [C.c_int] + [vp] * 8 + [vp]
    L.obca_b200_solve_host.restype = C.c_int
    L.obca_b200_solve_host.argtypes = [vp] + _abi.SOLVE_ARGTYPES_HOST
    ip = C.POINTER(C.c_int32)
    L.obca_b200_astar_batch.restype = C.c_int
    L.obca_b200_astar_batch.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, ip, ip, ip, C.c_int, vp, ip, C.c_int]
    L.obca_b200_reference_windows.restype = C.c_int
    L.obca_b200_reference_windows.argtypes = [C.c_int, vp, ip, C.c_int, ip, vp, C.c_int, vp]
    L.obca_b200_solve_indexed.restype = C.c_int
    L.obca_b200_solve_indexed.argtypes = [vp, C.c_int, vp, vp] + [vp] * 7 + [ip] + [vp] * 3 + [C.c_int] + [vp] * 8 + [vp]
    L.obca_b200_build_rows.restype = C.c_int
    L.obca_b200_build_rows.argtypes = [C.c_int, C.c_int, ip, vp, vp, vp, C.c_double, vp, vp, vp, vp]
    pp = C.POINTER(_abi.ObcaParams)
    L.obca_b200_loop_create.restype = C.c_int
    L.obca_b200_loop_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(_abi.LoopParams), pp, pp, pp, ip, vp, vp, vp]
    L.obca_b200_loop_reset.restype = C.c_int
    L.obca_b200_loop_reset.argtypes = [vp, vp, vp, vp]
    L.obca_b200_loop_run.restype = C.c_int
    L.obca_b200_loop_run.argtypes = [vp, C.c_int, vp]
    L.obca_b200_loop_read.restype = C.c_int
    L.obca_b200_loop_read.argtypes = [vp] * 10
    L.obca_b200_loop_launch_count.restype = C.c_int64
    L.obca_b200_loop_launch_count.argtypes = [vp]
    L.obca_b200_loop_destroy.restype = C.c_int
    L.obca_b200_loop_destroy.argtypes = [vp]
    if L.obca_b200_abi_version() != 2:
        raise RuntimeError("libobca_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("obca_b200: %s (rc=%d)" % (lib().obca_b200_strerror(rc).decode(), rc))

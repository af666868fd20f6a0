"""CPU tests of the C-ABI boundary: the CUDA library builds/loads, exports every symbol include/obca_b200.h
declares, the ctypes struct mirrors the C struct, and every entry point fails loudly without a device (no CPU
fallback anywhere on the product path)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import obca_testlib as common
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _abi, _lib, obca as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "obca_b200.h")


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    src = open(HDR).read()
    declared = set(re.findall(r"\b(obca_b200_\w+)\s*\(", src))
    assert declared == set(_lib.EXPORTS)
    L = C.CDLL(_lib.build())
    for s in declared:
        assert hasattr(L, s), s
    assert _lib.lib().obca_b200_abi_version() == 2


def test_params_struct_layout_matches_header(tmp_path):
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu",'
                 'sizeof(obca_params),offsetof(obca_params,Ts),offsetof(obca_params,Q),offsetof(obca_params,tol),'
                 'offsetof(obca_params,bound_push));return 0;}' % HDR)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-o", str(exe), str(c)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    P = _abi.ObcaParams
    assert got == [C.sizeof(P), P.Ts.offset, P.Q.offset, P.tol.offset, P.bound_push.offset]


def test_no_device_fails_loudly():
    if _has_gpu():
        pytest.skip("needs a box without a GPU")
    prm, a, d = common.fixture_arrays("demo1_N6_astar_free")
    ctx = C.c_void_p()
    rc = _lib.lib().obca_b200_create(C.byref(ctx), -1, 4, C.byref(prm))
    assert rc == -2 and not ctx.value            # OBCA_E_NODEVICE
    assert b"no host fallback" in _lib.lib().obca_b200_strerror(rc)
    with pytest.raises(RuntimeError):
        om.BatchSolver(prm, a["edge_ptr"], 4)
    with pytest.raises(RuntimeError):
        s = om.obca()
        s.obca_mpc4(float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], int(d["N"]), d["x0"], d["xL"], d["xU"], d["uL"],
                    d["uU"], d["xref"], int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"], float(d["dmin"]), d["ego"], d["u0"])


def test_argument_validation_without_compute():
    L = _lib.lib()
    prm, a, d = common.fixture_arrays("demo1_N6_astar_free")
    assert L.obca_b200_create(None, -1, 4, C.byref(prm)) == -1
    ctx = C.c_void_p()
    big = _abi.ObcaParams.from_buffer_copy(prm); big.N = 40
    assert L.obca_b200_create(C.byref(ctx), -1, 4, C.byref(big)) == -5         # OBCA_E_SIZE
    assert L.obca_b200_destroy(None) == -1
    assert L.obca_b200_scratch_bytes(None) == 0 and L.obca_b200_launch_count(None) == 0
    assert L.obca_b200_strerror(-5).startswith(b"problem size")


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under the package may import it"""
    pkg = os.path.join(ROOT, "vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "obca_oracle" not in txt, f
    code = "import sys; sys.path.insert(0, %r); import vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200.obca, " \
           "vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200.closed_loop, " \
           "vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200.sharding; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT
    subprocess.check_call([sys.executable, "-c", code])


def test_solve_batch_packing_shapes(monkeypatch):
    """solve_batch hands the C-ABI batch-major arrays and maps results back to the reference's (3,N+1)/(2,N) layout"""
    monkeypatch.setattr(om, "BatchSolver", common.OracleSolver)
    mode, d = common.load_fixture("demo1_N6_astar_free")
    B = 3
    x0 = np.tile(d["x0"], (B, 1)); x0[:, 1] += [0.0, 0.2, -0.2]
    xref = np.tile(d["xref"][None], (B, 1, 1))
    r = om.solve_batch(mode, float(d["Ts"]), d["P"], d["Q"], [d["R1"], d["R2"]], int(d["N"]), x0, d["xL"], d["xU"], d["uL"],
                       d["uU"], xref, int(d["nObs"]), d["vObs"], d["AObs"], d["bObs"], float(d["dmin"]), d["ego"], d["u0"])
    assert r["x"].shape == (B, 3, 7) and r["u"].shape == (B, 2, 6) and r["feas"].all()
    assert np.allclose(r["x"][:, :, 0], x0) and abs(r["obj"][0] - 4334.19729465) < 1e-5
    assert np.allclose(r["Ts_opt"], r["T"] * float(d["Ts"]))

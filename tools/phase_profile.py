#!/usr/bin/env python
"""Developer tool (GPU box): per-phase cycle shares of the solver kernel.

Builds csrc/libobca_b200_prof.so with -DOBCA_PROFILE (clock64 around every phase of the interior-point loop),
runs one batch and prints cycles per phase summed over warps.  Not part of the product path or the bench.

    python tools/phase_profile.py [cfg] [batch]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import _lib  # noqa: E402

PROF = os.path.join(_lib.CSRC, "libobca_b200_prof.so")


def build():
    _lib.build(force=True, extra_flags=["-DOBCA_PROFILE"], out=PROF, objdir=os.path.join(_lib.CSRC, "_obj_prof"))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":      # (here, without a GPU: the library travels to the GPU box)
        return build()
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    if not os.path.exists(PROF) or os.path.getmtime(PROF) < max(
            os.path.getmtime(os.path.join(_lib.CSRC, f)) for f in _lib.SOURCES + _lib.HEADERS[:2]):
        build()
    _lib.LIB = PROF
    _lib._stale = lambda: False
    import torch
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import obca as om, scenario as sc
    b = sc.make_batch(cfg, B)
    prm, a = sc.batch_arrays(b)
    s = om.BatchSolver(prm, a["edge_ptr"], B)
    L = _lib.lib()
    t = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.float64, device="cuda").contiguous()
    dv = {k: t(a[k]) for k in ("x0", "u0", "xref", "A", "b0", "db", "T_max", "term")}
    out = s.alloc_outputs(B, "cuda")
    buf = (C.c_ulonglong * 48)()
    for i in range(2):
        L.obca_b200_prof_read(buf, 1)
        s.solve(dv["x0"], dv["u0"], dv["xref"], dv["A"], dv["b0"], dv["db"], T_max=dv["T_max"], term=dv["term"], out=out)
        torch.cuda.synchronize()
    L.obca_b200_prof_read(buf, 0)
    names = ["start", "assemble", "combine", "asm-reduce", "control", "riccati", "rollout", "steps", "steps-reduce",
             "ls-setup", "trial", "trial-reduce", "update", "exit"]
    tot = float(sum(buf[:14]))
    it = out["iters"].cpu().numpy()
    print("cfg %d B %d kernel %.2f ms -> %.0f solves/s, iters mean %.1f max %d" % (cfg, B, s.last_kernel_ms(), B / s.last_kernel_ms() * 1e3, it.mean(), it.max()))
    print("  %-10s %8s %12s %14s %14s" % ("phase", "share", "cycles/iter", "block-warp work", "stage-warp work"))
    print("  sweep sub-steps: %.1f per iteration, %.0f cycles each (stage warp, clock64 around body + warp barrier)" % (
        buf[32 + 14] / max(1, it.sum()), buf[32 + 15] / max(1, buf[32 + 14])))
    # work columns: cycles warp 0 (a block warp) / the stage warp spent inside the phase's par() bodies, before the barrier
    for i, (n, v) in enumerate(zip(names, buf[:14])):
        print("  %-12s %6.2f %%  %10.0f %14.0f %14.0f" % (n, 100.0 * v / tot, v / max(1, it.sum()), buf[16 + i] / max(1, it.sum()),
                                                      buf[32 + i] / max(1, it.sum()) if i < 14 else 0))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate golden input fixtures by RUNNING the reference's own host-side code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_reference_fixtures.py

The reference (pure Python) imports casadi / matplotlib / skimage / ttictoc at module top level and
relies on ``np.size(ragged_list, 0)`` (NumPy < 1.24).  None of those are in this image, so the modules
are stubbed (they are only needed for plotting / the IPOPT call, never for the input builders) and
``np.size`` is shimmed to ``len`` for ragged lists.  Nothing from the reference is copied: its modules are
imported from /root/reference/src and *executed*; only their numeric outputs are stored.

Outputs (tests/golden/*.npz): the exact arguments the reference's ``closedLoop`` would hand to
``obca.obca_mpc4 / obca_mpc6`` (closed_loop.py:118,131,382,389) for demo1 / demo9, plus the A* routes and
H-representations quoted in SURVEY.md Appendix C.
"""
import hashlib
import os
import sys
import types

import numpy as np

REF = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_stubs():
    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything(self.__name__ + "." + name)

        def __call__(self, *a, **k):
            return _Anything(self.__name__ + "()")

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches",
                 "matplotlib.collections", "matplotlib.transforms",
                 "skimage", "skimage.morphology", "skimage.draw", "ttictoc"]:
        sys.modules[name] = _Anything(name)
    # closed_loop.py does `from casadi import *` and gets `np` from it
    cas = types.ModuleType("casadi")
    cas.np = np
    cas.casadi = cas
    cas.__all__ = ["np", "casadi"]
    sys.modules["casadi"] = cas

    _size = np.size

    def size(a, axis=None):
        try:
            return _size(a, axis)
        except ValueError:
            if axis in (0, None):
                return len(a)
            raise
    np.size = size


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    from demo_setting import problemSetting  # noqa
    from closed_loop import closedLoop  # noqa
    from a_star import a_star  # noqa
    from model_obstacle import obstacleModel  # noqa

    out = {}

    # ---- A* known answer (a_star.py:202-232)
    grid, start, goal = a_star(np.zeros((2, 2)), (0, 0), (1, 1)).demo_data()
    route = a_star(grid, start, goal).solve(grid, start, goal)
    np.savez(os.path.join(OUT, "astar_demo_data.npz"), grid=grid, start=np.array(start), goal=np.array(goal),
             route=np.array(route))

    # ---- per-demo MPC inputs
    def pack_free(mpc, N, ref_type):
        mpc.N_free = N
        mpc.update_obstacle_constraint(mpc.N_free, mpc.Ts, 0)
        if ref_type == "A_star":
            path = mpc.update_path(0, mpc.x0, mpc.xF, 0, "A_star")
            xref = mpc.update_reference_trajectory(mpc.N_free, path, mpc.x0)
        else:
            path = np.zeros((3, 0))
            xref = mpc.update_path(mpc.N_free, mpc.x0, mpc.xF, allAviable=0, type="startGoal_only")
        return dict(Ts=mpc.Ts, P=np.asarray(mpc.P_free), Q=np.asarray(mpc.Q_free),
                    R1=np.asarray(mpc.R_free[0]), R2=np.asarray(mpc.R_free[1]), N=N,
                    x0=np.asarray(mpc.x0, float), xL=np.asarray(mpc.xL, float), xU=np.asarray(mpc.xU, float),
                    uL=np.asarray(mpc.uL, float), uU=np.asarray(mpc.uU, float), xref=np.asarray(xref, float),
                    nObs=mpc.nObs, vObs=np.asarray(mpc.vObs, int), AObs=np.asarray(mpc.AObs, float),
                    bObs=np.asarray(mpc.bObs, float), dmin=mpc.dmin, ego=np.asarray(mpc.ego, float),
                    u0=np.asarray(mpc.u0, float), path=np.asarray(path, float))

    for demo, N, ref_type in [("demo1", 6, "A_star"), ("demo1", 5, "A_star"), ("demo9", 5, "A_star"),
                              ("demo9", 6, "A_star"), ("demo1", 6, "startGoal_only"),
                              ("demo9", 10, "startGoal_only"), ("demo6", 6, "A_star"), ("demo2", 6, "A_star")]:
        try:
            mpc = closedLoop(problemSetting(demo))
        except Exception as e:  # some demos are broken in the reference itself
            print("skip", demo, repr(e))
            continue
        if demo == "demo9":      # simulation.py:68-70 recommended settings
            mpc.Q_free = 0.5 * np.eye(3)
            mpc.P_free = mpc.Q_free
        d = pack_free(mpc, N, ref_type)
        tag = "%s_N%d_%s_free" % (demo, N, "astar" if ref_type == "A_star" else "sg")
        np.savez(os.path.join(OUT, tag + ".npz"), **d)
        h = hashlib.sha256(np.ascontiguousarray(d["path"]).tobytes()).hexdigest()[:16] if d["path"].size else "-"
        print(tag, "AObs", d["AObs"].shape, "vObs", d["vObs"], "path sha", h)

    # ---- fixed-time inputs with the dynamic obstacle time-stacked (closed_loop.py:122-134 with a given
    #      free-time result replaced by an A* window; Ts_opt = 2.0 is a typical inherited value)
    for demo, N in [("demo1", 6), ("demo9", 5)]:
        mpc = closedLoop(problemSetting(demo))
        path = mpc.update_path(0, mpc.x0, mpc.xF, 0, "A_star")
        mpc.N_fix = N
        xref = mpc.update_reference_trajectory(mpc.N_fix, path, mpc.x0)
        Ts_opt = 2.0
        mpc.update_obstacle_constraint(mpc.N_fix, Ts_opt, 1)
        ts = np.array([[mpc.x0[0] + 5, 99], [1, 9]], float) if demo == "demo1" else np.asarray(mpc.setting.terminal_set, float)
        if demo == "demo9":
            ts = np.array([[0.0, 99.0], [mpc.x0[1] + 4, 60.0]])   # simulation.py:72 recommended
        d = dict(Ts=Ts_opt, P=np.asarray(mpc.P_fix), Q=np.asarray(mpc.Q_fix), R1=np.asarray(mpc.R_fix[0]),
                 R2=np.asarray(mpc.R_fix[1]), N=N, x0=np.asarray(mpc.x0, float), xL=np.asarray(mpc.xL, float),
                 xU=np.asarray(mpc.xU, float), uL=np.asarray(mpc.uL, float), uU=np.asarray(mpc.uU, float),
                 xref=np.asarray(xref, float), nObs=mpc.nObs, vObs=np.asarray(mpc.vObs, int),
                 AObs=np.asarray(mpc.AObs, float), bObs=np.asarray(mpc.bObs, float), dmin=mpc.dmin,
                 ego=np.asarray(mpc.ego, float), u0=np.asarray(mpc.u0, float), terminal_set=ts)
        np.savez(os.path.join(OUT, "%s_N%d_fixed.npz" % (demo, N)), **d)
        print(demo, "fixed AObs", d["AObs"].shape, "vObs", d["vObs"])

    # ---- H-rep of a rotated rectangle (demo_setting.py:405-429 + model_obstacle.py:37-102)
    ps = problemSetting("demo1")
    verts = ps.get_obstacle(20, 5, np.pi / 6, 4, 2)
    A, b = obstacleModel().obstacle_H_Represent(1, np.array([5]), [verts])
    np.savez(os.path.join(OUT, "hrep_rotated_rect.npz"), verts=np.asarray(verts), A=A, b=b,
             args=np.array([20, 5, np.pi / 6, 4, 2]))
    print("hrep", np.hstack([A, b]))


if __name__ == "__main__":
    main()

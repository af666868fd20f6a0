"""Receding-horizon MPC orchestration on top of the batched B200 OBCA solver.

Two layers:

* ``closedLoop`` - host-side drop-in for the reference's orchestrator
  (/root/reference/src/closed_loop.py:16-630): same constructor argument (a ``problemSetting``), same
  attributes (``Ts, xL, xU, uL, uU, x0, xF, u0, Q_free, R_free, P_free, N_free, Q_fix, R_fix, P_fix, N_fix,
  terminal_set, ego, dmin, xOpt, uOpt, feas, Ts_opt``) and the same methods: ``mpc_openLoop_freeTime``
  (113-120), ``mpc_openLoop_fixTime`` (122-140), ``closed_loop_mpc`` (142-209), ``closed_loop_mpc3``
  (211-321), ``closed_loop_mpc4`` (323-441), ``update_obstacle`` (445-486), ``update_obstacle_constraint``
  (488-500), ``update_reference_trajectory`` (502-528), ``update_path`` (530-589), ``sensor`` (591-630).
  Every solve goes through ``obca.obca_mpc4 / obca_mpc6 / obca_mpc8 / obca2`` (a batch of one on the GPU).
  Plotting (``draw``) is out of scope: the loops return their logs instead of animating them.

* ``ClosedLoopBatch`` - the same ``closed_loop_mpc4`` logic for thousands of Monte-Carlo scenarios advanced in
  lock-step (SURVEY.md 8(d) cfg 4): per step one FREE launch on the scenarios whose lidar sees nothing, one
  FIXED_SET launch on the others and one FIXED_NOTERM launch on that launch's failures
  (closed_loop.py:381-398).  Scenario state is NumPy (struct of arrays); the solver calls take the
  whole subset at once.

The reference's quirks are kept on purpose: the "plant" is the prediction itself (Q11),
``update_path(allAviable=1)`` overwrites ``Ts`` with ``Ts_opt`` (Q6), the reference yaw is recomputed by
``atan2`` in the fixed phase, the terminal set of the closed loop is ``[[x0.x + 5, 99], [1, 9]]``
(closed_loop.py:371), and ``sensor`` filters the info rows but not the polygons (Q8).
"""
from __future__ import annotations

import numpy as np

from . import _abi, model_obstacle as mo
from .a_star import a_star
from .obca import BatchSolver, obca
from .scenario import update_reference_trajectory as _window


def _yaw_path(xy):
    """``a_star.create_reference_path`` on an (n, 2) point list -> (3, n): yaw to the next point, last copied."""
    xy = np.asarray(xy, float)
    yaw = np.arctan2(xy[1:, 1] - xy[:-1, 1], xy[1:, 0] - xy[:-1, 0])
    yaw = np.concatenate([yaw, yaw[-1:]])
    return np.vstack([xy[:, 0], xy[:, 1], yaw])


class closedLoop:
    def __init__(self, problem_setting, solver=None, verbose=False):
        s = self.setting = problem_setting
        self.verbose = verbose
        self.obs_model = mo.obstacleModel()
        self.obca_solver = solver if solver is not None else obca()
        self.path_solver = a_star(s.org_gridMap, (s.startPose[1], s.startPose[0]), (s.goalPose[1], s.goalPose[0]))
        self.Ts = 0.1
        self.nx, self.nu = 3, 2
        self.xL = [s.xL[0], s.xL[1], -np.pi]; self.xU = [s.xU[0], s.xU[1], np.pi]
        self.uL = [-0.6, -np.pi / 6]; self.uU = [0.6, np.pi / 6]
        self.x0 = s.startPose; self.xF = s.goalPose; self.u0 = [0, 0]
        self.fixtime = 0
        self.nObs = 0; self.vObs = np.ones(0, dtype=int); self.AObs = []; self.bObs = []
        self.xref = []; self.uref = []
        self.ego = [1.7, 0.75, 1.7, 0.75]; self.dmin = 0.05
        self.xOpt = []; self.uOpt = []; self.feas = False; self.Ts_opt = self.Ts
        self.Q_free = 0.1 * np.eye(3); self.R_free = [0.01 * np.eye(2), 0.1 * np.eye(2)]; self.P_free = self.Q_free
        self.N_free = 6
        self.Q_fix = 0.001 * np.eye(3); self.R_fix = [0.01 * np.eye(2), 1.0 * np.eye(2)]; self.P_fix = self.Q_fix
        self.N_fix = 6
        self.terminal_set = []
        self.dyn_orignal_info = s.dyn_obs_info
        self.dyn_fulltime_info = []
        self.dyn_loc = []
        self.max_steps = 30            # closed_loop.py:431-432

    def _say(self, *a):
        if self.verbose:
            print(*a)

    # ---- open loop (closed_loop.py:113-140) -------------------------------------------------------------
    def mpc_openLoop_freeTime(self):
        self.update_obstacle_constraint(self.N_free, self.Ts, 0)
        self.xref = self.update_path(self.N_free, self.x0, self.xF, allAviable=0, type='startGoal_only')
        self.xOpt, self.uOpt, self.feas, self.Ts_opt = self._free()

    def mpc_openLoop_fixTime(self):
        self.xref = self.xOpt
        self.xref = self.update_path(0, 0, 0, allAviable=1, type='')
        self.update_obstacle_constraint(self.N_fix, self.Ts_opt, 1)
        self.terminal_set = self.setting.terminal_set
        self.fixtime = 1
        self.xOpt, self.uOpt, self.feas, self.Ts_opt = self._fixed()

    def _free(self):
        return self.obca_solver.obca_mpc4(self.Ts, self.P_free, self.Q_free, self.R_free, self.N_free, self.x0,
                                          self.xL, self.xU, self.uL, self.uU, self.xref, self.nObs, self.vObs,
                                          self.AObs, self.bObs, self.dmin, self.ego, self.u0)

    def _fixed(self):
        args = (self.Ts, self.P_fix, self.Q_fix, self.R_fix, self.N_fix, self.x0, self.xL, self.xU, self.uL, self.uU,
                self.xref, self.nObs, self.vObs, self.AObs, self.bObs, self.dmin, self.ego, self.u0, self.uOpt)
        r = self.obca_solver.obca_mpc6(*args, self.terminal_set)
        if r[2] is False or r[2] == False:  # noqa: E712  (drop the terminal set: closed_loop.py:135-140)
            r = self.obca_solver.obca_mpc8(*args)
        return r

    def _obca2_free(self):
        return self.obca_solver.obca2(self.Ts, self.P_free, self.Q_free, self.R_free, self.N_free, self.x0, self.u0,
                                      self.xL, self.xU, self.uL, self.uU, self.xref, self.uref, self.nObs, self.vObs,
                                      self.AObs, self.bObs, self.dmin, self.ego, self.fixtime, '', [])

    # ---- receding-horizon loops -------------------------------------------------------------------------
    def _loop(self, variant):
        """Shared body of closed_loop_mpc (variant 1), closed_loop_mpc3 (3) and closed_loop_mpc4 (4)."""
        k = 0
        path = self.update_path(0, self.x0, self.xF, 0, 'A_star')
        x_opt = [self.x0]; u_opt = []; T_opt = []; x_openLoop = []
        goal = self.setting.goalPose
        while (self.x0[0] - goal[0]) ** 2 + (self.x0[1] - goal[1]) ** 2 >= 0.1:
            self.update_obstacle(k, self.Ts_opt)
            if variant != 1:
                self.sensor()
            use_free = variant == 1 or self.fixtime == 0 or (variant == 4 and k == 0)
            if use_free:
                self.update_obstacle_constraint(self.N_free, self.Ts, 0)
                self.xref = self.update_reference_trajectory(self.N_free, path, self.x0)
                if variant == 4:
                    self.xOpt, self.uOpt, self.feas, self.Ts_opt = self._free()
                else:
                    self.xOpt, self.uOpt, self.feas, self.Ts_opt = self._obca2_free()
            else:
                self.xref = self.update_reference_trajectory(self.N_fix, path, self.x0)
                if variant == 4:
                    for i in range(self.N_fix - 5):                       # closed_loop.py:362-363
                        self.xref[:, i] = self.xOpt[:, i + 1]
                self.xref = self.update_path(0, 0, 0, allAviable=1, type='')
                if variant == 4:
                    self.terminal_set = np.array([[self.x0[0] + 5, 99], [1, 9]])   # closed_loop.py:371
                else:
                    self.terminal_set = self.setting.terminal_set
                self.update_obstacle_constraint(self.N_fix, self.Ts_opt, 1)
                self.xOpt, self.uOpt, self.feas, self.Ts_opt = self._fixed()
            self._say('MPC -- %s, fixtime = %i' % ('Success' if self.feas else 'Failed', self.fixtime))
            if not self.feas:
                break
            self.u0 = self.uOpt[:, 0].T
            self.x0 = self.xOpt[:, 1].T
            x_opt.append(self.x0); u_opt.append(self.u0); T_opt.append(self.Ts_opt); x_openLoop.append(self.xOpt.T)
            k += 1
            if variant == 4 and k == self.max_steps:
                break
        self.xOpt = np.asarray(x_opt).T
        self.xref = path
        self.Ts_opt = T_opt
        return x_openLoop, x_opt, u_opt, T_opt

    def closed_loop_mpc(self):
        return self._loop(1)

    def closed_loop_mpc3(self):
        return self._loop(3)

    def closed_loop_mpc4(self):
        return self._loop(4)

    # ---- dynamic obstacles (closed_loop.py:445-486) ------------------------------------------------------
    def update_obstacle(self, k, Ts_opt):
        live = []; polys = []
        for row in self.dyn_orignal_info:
            if k > row[9]:
                row[0] += Ts_opt * row[5] * np.cos(row[2])
                row[1] += Ts_opt * row[5] * np.sin(row[2])
            if k >= row[9]:
                live.append(row)
                polys.append(mo.get_obstacle(row[0], row[1], row[2], row[3], row[4]))
        self.setting.dyn_obs_info = live
        self.dyn_loc.append(polys)
        self.setting.add_dynamic_obstacle(live)

    def update_obstacle_constraint(self, N, Ts, dynobs_exist):
        self.setting.rebuild_lObs(N, Ts, dynObs_exist=dynobs_exist)
        self.lObs = self.setting.lObs
        self.nObs = self.setting.nObs
        self.vObs = self.setting.vObs
        fv = [len(p) for p in self.lObs]
        self.AObs, self.bObs = self.obs_model.obstacle_H_Represent(len(self.lObs), fv, self.lObs)

    def update_reference_trajectory(self, N, ref_trajectory, current_state):
        return _window(N, np.asarray(ref_trajectory, float), np.asarray(current_state, float))

    def update_path(self, N, x0, xF, allAviable, type):
        if allAviable == 0:
            ref = np.zeros((self.nx, N + 1))
            if type == 'startGoal_only':
                ref[:, 0] = np.asarray(x0, float)[:3]
                ref[:, 1:] = np.asarray(xF, float)[:3, None]
            elif type == 'startGoal_smooth':
                x0 = np.asarray(x0, float); xF = np.asarray(xF, float)
                for k in range(N + 1):
                    ref[0, k] = ((xF[0] - x0[0]) / N) * k + x0[0]
                    ref[1, k] = ((xF[1] - x0[1]) / N) * k + x0[1]
                    if k >= 1:
                        ref[2, k - 1] = np.arctan2(ref[1, k] - ref[1, k - 1], ref[0, k] - ref[0, k - 1])
                ref[2, N] = ref[2, N - 1]
            elif type == 'A_star':
                s = self.setting
                start = (s.startPose[1], s.startPose[0]); goal = (s.goalPose[1], s.goalPose[0])
                route = self.path_solver.solve(s.org_gridMap, start, goal)
                ref = np.asarray(self.path_solver.create_reference_path(self.path_solver.rebuild_path(route)), float).T
            return ref
        # allAviable == 1: re-sample the current reference (N_fix/N_free points per segment), recompute the yaw,
        # and inherit the time step (closed_loop.py:571-587)
        xr = np.asarray(self.xref, float)
        per = int(self.N_fix / self.N_free)
        pts = []
        for i in range(self.N_free):
            xx = np.linspace(xr[0][i], xr[0][i + 1], num=per, endpoint=False)
            yy = np.linspace(xr[1][i], xr[1][i + 1], num=per, endpoint=False)
            pts += [[xx[j], yy[j]] for j in range(per)]
        pts.append([xr[0][-1], xr[1][-1]])
        ref = _yaw_path(pts)
        self.N_fix = ref.shape[1] - 1
        self.Ts_opt = (self.N_free * self.Ts_opt) / self.N_fix
        self.Ts = self.Ts_opt
        return ref

    # ---- lidar gate (closed_loop.py:591-630) ---------------------------------------------------------------
    def sensor(self):
        cx, cy, th = self.x0[0], self.x0[1], self.x0[2]
        front = (cx + self.ego[0] * np.cos(th), cy + self.ego[0] * np.sin(th))
        seen = []
        self.fixtime = 0
        for i, poly in enumerate(self.dyn_loc[-1]):
            hit = any(np.hypot(front[0] - poly[j][0], front[1] - poly[j][1]) <= self.setting.senseDis for j in range(4))
            if hit:
                self.fixtime = 1
                seen.append(self.setting.dyn_obs_info[i])
            poly.append(1 if hit else 0)
        # polygons, vertex counts and info rows of the DETECTED obstacles only, kept paired (with two live obstacles and
        # only the second one in range the reference's combine_obstacle, demo_setting.py:445-452, would pair the first
        # polygon with the second velocity - SURVEY Q8 - or raise; the detected obstacle must be the constrained one)
        self.setting.add_dynamic_obstacle(seen)


# ==========================================================================================================
# lock-step batch of closed loops
# ==========================================================================================================
def box_hrep_batch(cx, cy, th, length, width):
    """Vectorised ``get_obstacle`` + ``obstacle_H_Represent`` for B rectangles: (B,) arrays -> A (B,4,2), b (B,4).
    Same branch rules as model_obstacle.obstacle_H_Represent (exact equality picks the axis-aligned rows,
    slanted edges stay unnormalised)."""
    cx, cy, th = (np.asarray(v, float) for v in (cx, cy, th))
    l = np.asarray(length, float) / 2; w = np.asarray(width, float) / 2
    c, s = np.cos(th), np.sin(th)
    V = np.stack([np.stack([cx - l * c - w * s, cy - l * s + w * c], -1),
                  np.stack([cx + l * c - w * s, cy + l * s + w * c], -1),
                  np.stack([cx + l * c + w * s, cy + l * s - w * c], -1),
                  np.stack([cx - l * c + w * s, cy - l * s - w * c], -1)], 1)       # (B,4,2)
    V = np.concatenate([V, V[:, :1]], 1)
    x1, y1, x2, y2 = V[:, :-1, 0], V[:, :-1, 1], V[:, 1:, 0], V[:, 1:, 1]
    vert = x1 == x2; horz = (y1 == y2) & ~vert
    with np.errstate(divide="ignore", invalid="ignore"):
        a = (y2 - y1) / (x2 - x1)
        cc = y1 - a * x1
    right = x1 < x2
    A0 = np.where(vert, np.where(y2 < y1, 1.0, -1.0), np.where(horz, 0.0, np.where(right, -a, a)))
    A1 = np.where(vert, 0.0, np.where(horz, np.where(right, 1.0, -1.0), np.where(right, 1.0, -1.0)))
    b = np.where(vert, np.where(y2 < y1, x1, -x1), np.where(horz, np.where(right, y1, -y1), np.where(right, cc, -cc)))
    return np.stack([A0, A1], -1), b, V


class ClosedLoopBatch:
    """``closed_loop_mpc4`` for B Monte-Carlo scenarios in lock-step on one GPU.

    All scenarios share the static map, start/goal and hence the A* path; each has its own moving boxes: ``dyn`` rows
    ``[cx, cy, theta, l, w, v, start_step]``, (B,7) for one box per scenario or (B,D,7) for D (the reference's demos 6, 7,
    8 and 11 carry two, demo_setting.py:221-338).  As in the single-scenario loop (``sensor``), a fixed-time solve is
    built from the static obstacles plus the boxes the lidar sees at that step, in their order, each with its own
    velocity; scenarios are grouped by the number of boxes seen (one solver context per count).  ``run`` returns the
    closed-loop logs."""

    def __init__(self, setting, dyn, N=5, Q_free=0.5, sense=8.0, device=-1, init=_abi.INIT_WARM | _abi.RECOVER, max_steps=30,
                 solver_factory=None):
        self.s = setting
        self.dyn = np.array(dyn, float)
        if self.dyn.ndim == 2:
            self.dyn = self.dyn[:, None, :]
        self.B, self.D = self.dyn.shape[0], self.dyn.shape[1]
        self.N = int(N)
        self.sense = float(sense)
        self.max_steps = max_steps
        self.device = device
        self.init = init
        self.uL = np.array([-0.6, -np.pi / 6]); self.uU = -self.uL
        self.ego = np.array([1.7, 0.75, 1.7, 0.75]); self.dmin = 0.05
        self.Q_free = Q_free * np.eye(3); self.R_free = [0.01 * np.eye(2), 0.1 * np.eye(2)]
        self.Q_fix = 0.001 * np.eye(3); self.R_fix = [0.01 * np.eye(2), 1.0 * np.eye(2)]
        pl = a_star(setting.org_gridMap, (setting.startPose[1], setting.startPose[0]), (setting.goalPose[1], setting.goalPose[0]))
        route = pl.solve(setting.org_gridMap, (setting.startPose[1], setting.startPose[0]),
                         (setting.goalPose[1], setting.goalPose[0]))
        self.path = np.asarray(pl.create_reference_path(pl.rebuild_path(route)), float).T
        # static rows (one time block; they never move)
        sv = [int(v) for v in setting.static_vObs]
        self.A_s, b_s = mo.obstacleModel().obstacle_H_Represent(len(sv), sv, setting.static_lObs)
        self.b_s = b_s.reshape(-1)
        self.edges_s = [v - 1 for v in sv]
        self._solvers = {}
        self._factory = solver_factory or (lambda prm, ep, cap: BatchSolver(prm, ep, cap, device))
        self.launches = 0
        self.solves = 0

    def _solver(self, mode, with_dyn):
        """solver context for `mode` with `with_dyn` moving boxes in the NLP (False / True count as 0 / 1)"""
        key = (mode, int(with_dyn))
        if key not in self._solvers:
            edges = self.edges_s + [4] * int(with_dyn)
            ep = np.concatenate([[0], np.cumsum(edges)]).astype(np.int32)
            free = _abi.is_free(mode)
            prm = _abi.make_params(mode, self.N, len(edges), int(ep[-1]), 0.1, self.Q_free if free else self.Q_fix,
                                   self.Q_free if free else self.Q_fix, self.R_free if free else self.R_fix,
                                   self.s.xL, self.s.xU, self.uL, self.uU, self.dmin, self.ego, init=self._init_of(mode))
            self._solvers[key] = (self._factory(prm, ep, self.B), ep)
        return self._solvers[key]

    def _init_of(self, mode):
        """The terminal-set solve has its own fallback (the solve without the set, closed_loop.py:389-395) and is often
        truly infeasible: it runs without the recovery rules; the other two modes keep them."""
        return ((self.init & 15) | _abi.INIT_NORESTO) if mode == _abi.MODE_FIXED_SET else self.init

    def close(self):
        for s, _ in self._solvers.values():
            if hasattr(s, "close"):
                s.close()
        self._solvers = {}

    def _windows(self, x0):
        d = (x0[:, 0:1] - self.path[0][None]) ** 2 + (x0[:, 1:2] - self.path[1][None]) ** 2
        i0 = np.argmin(d, axis=1)
        idx = np.minimum(i0[:, None] + np.arange(self.N + 1)[None], self.path.shape[1] - 1)
        return self.path[:, idx].transpose(1, 2, 0).copy()                    # (n, N+1, 3)

    def run(self, terminal_rule="shipped"):
        B, N = self.B, self.N
        s = self.s
        x0 = np.tile(np.asarray(s.startPose, float), (B, 1)); u0 = np.zeros((B, 2))
        Ts = np.full(B, 0.1); Ts_opt = np.full(B, 0.1)
        alive = np.ones(B, bool); failed = np.zeros(B, bool)
        steps = np.zeros(B, int)
        dyn = self.dyn.copy()
        traj = np.full((B, self.max_steps + 1, 3), np.nan); traj[:, 0] = x0
        mode_log = np.full((B, self.max_steps), -1, int)
        goal = np.asarray(s.goalPose, float)
        xprev = np.zeros((B, N + 1, 3))
        self.launches = 0; self.solves = 0
        for k in range(self.max_steps):
            alive &= ((x0[:, 0] - goal[0]) ** 2 + (x0[:, 1] - goal[1]) ** 2 >= 0.1)
            if not alive.any():
                break
            # update_obstacle: appear at k == start step, move afterwards by the last optimal step
            mv = alive[:, None] & (k > dyn[:, :, 6])
            step = np.where(mv, Ts_opt[:, None] * dyn[:, :, 5], 0.0)
            dyn[:, :, 0] += step * np.cos(dyn[:, :, 2])
            dyn[:, :, 1] += step * np.sin(dyn[:, :, 2])
            live = alive[:, None] & (k >= dyn[:, :, 6])
            flat = dyn.reshape(B * self.D, 7)
            A_d, b_d, V = box_hrep_batch(flat[:, 0], flat[:, 1], flat[:, 2], flat[:, 3], flat[:, 4])
            A_d = A_d.reshape(B, self.D, 4, 2); b_d = b_d.reshape(B, self.D, 4); V = V.reshape(B, self.D, 5, 2)
            # sensor: car-front midpoint to the four vertices of every live box
            fx = x0[:, 0] + self.ego[0] * np.cos(x0[:, 2]); fy = x0[:, 1] + self.ego[0] * np.sin(x0[:, 2])
            dist = np.hypot(fx[:, None, None] - V[:, :, :4, 0], fy[:, None, None] - V[:, :, :4, 1]).min(2)
            seen = live & (dist <= self.sense)
            fix = seen.any(1) & (k > 0)
            free = alive & ~fix
            xopt = np.zeros((B, N + 1, 3)); uopt = np.zeros((B, N, 2)); feas = np.zeros(B, bool)
            newT = Ts_opt.copy()
            if free.any():
                i = np.where(free)[0]
                sol, ep = self._solver(_abi.MODE_FREE, False)
                xr = self._windows(x0[i])
                Tm = ((xr[:, N, 0] - x0[i, 0]) + (xr[:, N, 1] - x0[i, 1])) / (N * self.uU[0] * Ts[i]) + 1.0
                o = sol.solve_host(x0[i], u0[i], xr, self.A_s, self.b_s, None, T_max=Tm, Ts=Ts[i])
                self.launches += 1; self.solves += len(i)
                xopt[i] = o["x"]; uopt[i] = o["u"]; feas[i] = o["status"] >= 0
                newT[i] = o["T"] * Ts[i]
                mode_log[i, k] = _abi.MODE_FREE
            if fix.any():
                i = np.where(fix)[0]
                xr = self._windows(x0[i])
                for c in range(N - 5):                                             # closed_loop.py:362-363
                    xr[:, c] = xprev[i, c + 1]
                # update_path(allAviable=1) with N_fix == N_free: same points, yaw recomputed by atan2
                yaw = np.arctan2(xr[:, 1:, 1] - xr[:, :-1, 1], xr[:, 1:, 0] - xr[:, :-1, 0])
                xr[:, :-1, 2] = yaw; xr[:, -1, 2] = yaw[:, -1]
                Ts[i] = Ts_opt[i]                                                  # Q6
                if terminal_rule == "shipped":
                    term = np.stack([x0[i, 0] + 5, np.full(len(i), 1.0), np.full(len(i), 9.0)], 1)
                else:   # the demo9 recommendation of simulation.py:72
                    term = np.stack([np.full(len(i), 5.0), x0[i, 1] + 4, np.full(len(i), 60.0)], 1)
                # rows of the boxes each scenario sees, in their order; scenarios grouped by how many that is
                nseen = seen[i].sum(1)
                xo = np.zeros((len(i), N + 1, 3)); uo = np.zeros((len(i), N, 2)); ok = np.zeros(len(i), bool)
                for c in np.unique(nseen):
                    g = np.where(nseen == c)[0]                                  # positions within i
                    ig = i[g]
                    pick = np.argsort(~seen[ig], axis=1, kind="stable")[:, :c]        # indices of the seen boxes, ascending
                    rows = np.arange(len(ig))[:, None]
                    Ad = A_d[ig][rows, pick].reshape(len(ig), 4 * c, 2); bd = b_d[ig][rows, pick].reshape(len(ig), 4 * c)
                    dd = dyn[ig][rows, pick]                                          # (n, c, 7)
                    sh = (Ts_opt[ig][:, None] * dd[:, :, 5])[:, :, None] * (
                        A_d[ig][rows, pick][..., 0] * np.cos(dd[:, :, 2])[:, :, None] + A_d[ig][rows, pick][..., 1] * np.sin(dd[:, :, 2])[:, :, None])
                    A = np.concatenate([np.tile(self.A_s[None], (len(ig), 1, 1)), Ad], 1)
                    b0 = np.concatenate([np.tile(self.b_s[None], (len(ig), 1)), bd], 1)
                    db = np.concatenate([np.zeros((len(ig), self.b_s.shape[0])), sh.reshape(len(ig), 4 * c)], 1)
                    sol, ep = self._solver(_abi.MODE_FIXED_SET, int(c))
                    o = sol.solve_host(x0[ig], u0[ig], xr[g], A, b0, db, term=term[g], Ts=Ts[ig])
                    self.launches += 1; self.solves += len(ig)
                    okg = o["status"] >= 0
                    xg, ug = o["x"], o["u"]
                    if (~okg).any():
                        j = np.where(~okg)[0]
                        sol8, _ = self._solver(_abi.MODE_FIXED_NOTERM, int(c))
                        o8 = sol8.solve_host(x0[ig[j]], u0[ig[j]], xr[g[j]], A[j], b0[j], db[j], Ts=Ts[ig[j]])
                        self.launches += 1; self.solves += len(j)
                        xg[j] = o8["x"]; ug[j] = o8["u"]; okg[j] = o8["status"] >= 0
                        mode_log[ig[j], k] = _abi.MODE_FIXED_NOTERM
                    xo[g] = xg; uo[g] = ug; ok[g] = okg
                mode_log[i[ok & (mode_log[i, k] < 0)], k] = _abi.MODE_FIXED_SET
                xopt[i] = xo; uopt[i] = uo; feas[i] = ok
                newT[i] = Ts[i]
            bad = alive & ~feas
            failed |= bad
            alive &= feas
            Ts_opt = np.where(alive, newT, Ts_opt)
            u0[alive] = uopt[alive, 0]; x0[alive] = xopt[alive, 1]
            xprev[alive] = xopt[alive]
            steps[alive] += 1
            traj[alive, k + 1] = x0[alive]
        reached = ((x0[:, 0] - goal[0]) ** 2 + (x0[:, 1] - goal[1]) ** 2 < 0.1)
        return dict(traj=traj, steps=steps, failed=failed, reached=reached, mode=mode_log, x=x0, u=u0, Ts_opt=Ts_opt,
                    launches=self.launches, solves=self.solves)


class ClosedLoopDevice(ClosedLoopBatch):
    """``ClosedLoopBatch`` with the whole receding-horizon loop resident on the GPU (``obca_b200_loop_*`` of
    include/obca_b200.h): scenario state, input builders, work lists and the three solver modes stay in HBM; the host
    issues the launches of all steps without waiting and reads the logs once at the end."""

    def __init__(self, setting, dyn, N=5, Q_free=0.5, sense=8.0, device=-1, init=_abi.INIT_WARM | _abi.RECOVER, max_steps=30,
                 speculative=False):
        """``speculative``: solve without the terminal set beside the solve with it on every detected scenario (same
        results as the sequential fallback; pays off where the terminal-set solve mostly fails)."""
        super().__init__(setting, dyn, N=N, Q_free=Q_free, sense=sense, device=device, init=init, max_steps=max_steps)
        if self.D != 1:
            raise ValueError("the device-resident loop carries one moving box per scenario (obca_b200_loop_*); "
                             "ClosedLoopBatch runs scenarios with several")
        self.speculative = bool(speculative)
        self._loop = None
        self._rule = None

    def _params(self, mode, with_dyn):
        edges = self.edges_s + ([4] if with_dyn else [])
        free = _abi.is_free(mode)
        return _abi.make_params(mode, self.N, len(edges), int(sum(edges)), 0.1, self.Q_free if free else self.Q_fix,
                                self.Q_free if free else self.Q_fix, self.R_free if free else self.R_fix,
                                self.s.xL, self.s.xU, self.uL, self.uU, self.dmin, self.ego, init=self._init_of(mode))

    def _create(self, rule):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        lp = _abi.LoopParams(N=self.N, max_steps=self.max_steps, n_static=len(self.edges_s),
                             rows_static=int(sum(self.edges_s)), path_len=self.path.shape[1], terminal_rule=rule,
                             sense=self.sense, goal_tol=0.1, Ts0=0.1, speculative=int(self.speculative))
        lp.goal[0], lp.goal[1] = float(self.s.goalPose[0]), float(self.s.goalPose[1])
        for j in range(3):
            lp.start[j] = float(self.s.startPose[j])
        pf, ps, pn = (self._params(_abi.MODE_FREE, False), self._params(_abi.MODE_FIXED_SET, True),
                      self._params(_abi.MODE_FIXED_NOTERM, True))
        edges = np.asarray(self.edges_s, np.int32)
        A_s = np.ascontiguousarray(self.A_s, float); b_s = np.ascontiguousarray(self.b_s, float)
        path = np.ascontiguousarray(self.path.T, float)                                  # (M, 3)
        h = C.c_void_p()
        _lib.check(L.obca_b200_loop_create(C.byref(h), self.device, self.B, C.byref(lp), C.byref(pf), C.byref(ps),
                                           C.byref(pn), edges.ctypes.data_as(C.POINTER(C.c_int32)), A_s.ctypes.data,
                                           b_s.ctypes.data, path.ctypes.data))
        self._loop, self._rule = h, rule

    def close(self):
        if self._loop is not None:
            from . import _lib
            _lib.lib().obca_b200_loop_destroy(self._loop)
            self._loop = None
        super().close()

    def start(self, terminal_rule="shipped", steps=None):
        """Reset the scenarios and issue ``steps`` (default: all) receding-horizon steps; returns without waiting."""
        from . import _lib
        rule = 0 if terminal_rule == "shipped" else 1
        if self._loop is None or self._rule != rule:
            self.close()
            self._create(rule)
        L = _lib.lib()
        dyn = np.ascontiguousarray(self.dyn[:, 0, :], float)
        cs = np.ascontiguousarray(np.stack([np.cos(dyn[:, 2]), np.sin(dyn[:, 2])], 1))
        self._launches0 = int(L.obca_b200_loop_launch_count(self._loop))      # (the library's count is cumulative)
        _lib.check(L.obca_b200_loop_reset(self._loop, dyn.ctypes.data, cs.ctypes.data, None))
        _lib.check(L.obca_b200_loop_run(self._loop, self.max_steps if steps is None else int(steps), None))

    def read(self):
        import ctypes as C
        from . import _lib
        B, K = self.B, self.max_steps
        traj = np.empty((B, K + 1, 3)); steps = np.empty(B, np.int32); failed = np.empty(B, np.int32)
        mode = np.empty((B, K), np.int32); x = np.empty((B, 3)); u = np.empty((B, 2)); ts = np.empty(B)
        solves = np.zeros(3, np.int64)
        L = _lib.lib()
        _lib.check(L.obca_b200_loop_read(self._loop, traj.ctypes.data, steps.ctypes.data, failed.ctypes.data,
                                         mode.ctypes.data, x.ctypes.data, u.ctypes.data, ts.ctypes.data,
                                         solves.ctypes.data, None))
        goal = np.asarray(self.s.goalPose, float)
        self.solves = int(solves.sum()); self.launches = int(L.obca_b200_loop_launch_count(self._loop)) - getattr(self, "_launches0", 0)
        return dict(traj=traj, steps=steps.astype(int), failed=failed.astype(bool),
                    reached=((x[:, 0] - goal[0]) ** 2 + (x[:, 1] - goal[1]) ** 2 < 0.1), mode=mode.astype(int), x=x, u=u,
                    Ts_opt=ts, launches=self.launches, solves=self.solves, solves_by_mode=solves.tolist())

    def run(self, terminal_rule="shipped"):
        self.start(terminal_rule)
        return self.read()


def build_rows_device(verts, vObs, vel=None, Ts=0.1, with_db=True):
    """``obstacle_H_Represent`` + the time-stacking of ``rebuild_lObs`` on the GPU (``obca_b200_build_rows``).

    ``verts`` CUDA float64 tensor (B, sum(vObs), 2): the polygons' vertices back to back, as the reference lists
    them; ``vel`` (B, nObs, 3) = speed, cos, sin of each obstacle's heading, or None (static); ``Ts`` float or (B,)
    tensor.  Returns ``edge_ptr`` (host), ``A`` (B,R,2), ``b0`` (B,R), ``db`` (B,R) with b_k = b0 + k*db."""
    import ctypes as C
    import torch
    from . import _lib
    vObs = [int(v) for v in vObs]
    ep = np.concatenate([[0], np.cumsum([v - 1 for v in vObs])]).astype(np.int32)
    R = int(ep[-1]); B = verts.shape[0]
    if not verts.is_cuda or verts.dtype != torch.float64 or tuple(verts.shape[1:]) != (sum(vObs), 2):
        raise ValueError("verts must be a CUDA float64 tensor of shape (B, %d, 2)" % sum(vObs))
    verts = verts.contiguous()
    A = torch.empty((B, R, 2), dtype=torch.float64, device=verts.device)
    b0 = torch.empty((B, R), dtype=torch.float64, device=verts.device)
    db = torch.empty((B, R), dtype=torch.float64, device=verts.device) if with_db else None
    tsv = Ts.contiguous() if torch.is_tensor(Ts) else None
    if vel is not None:
        vel = vel.contiguous()
    with torch.cuda.device(verts.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().obca_b200_build_rows(
            B, len(vObs), ep.ctypes.data_as(C.POINTER(C.c_int32)), verts.data_ptr(),
            None if vel is None else vel.data_ptr(), None if tsv is None else tsv.data_ptr(),
            0.0 if tsv is not None else float(Ts), A.data_ptr(), b0.data_ptr(), None if db is None else db.data_ptr(), st))
    return ep, A, b0, db


def demo9_monte_carlo(B, seed=20221209 + 4):
    """SURVEY 8(d) cfg 4: demo9 map, one 2x2 box starting at (8, U[35,55]) heading -pi/2 with speed U[0.2,0.8],
    appearing at step U{0..5}."""
    rng = np.random.default_rng(seed)
    dyn = np.zeros((B, 7))
    dyn[:, 0] = 8.0; dyn[:, 1] = rng.uniform(35, 55, B); dyn[:, 2] = -np.pi / 2
    dyn[:, 3] = 2.0; dyn[:, 4] = 2.0; dyn[:, 5] = rng.uniform(0.2, 0.8, B); dyn[:, 6] = rng.integers(0, 6, B)
    return dyn

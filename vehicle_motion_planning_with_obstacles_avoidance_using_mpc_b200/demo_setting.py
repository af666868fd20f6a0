"""Scenario definitions: the reference's hard-coded demos as data, plus obstacle bookkeeping.

Host-side drop-in for ``problemSetting`` (/root/reference/src/demo_setting.py:11-488): same attribute names
(``xL, xU, startPose, goalPose, static_lObs, static_vObs, dyn_obs_info, dyn_lObs, terminal_set, org_gridMap,
senseDis, nObs, vObs, lObs``) and the same methods the MPC layer calls (``add_dynamic_obstacle`` 374-403,
``get_obstacle`` 405-429, ``combine_obstacle`` 431-455, ``rebuild_lObs`` 457-473).  The eleven demos of
``set_problem`` (82-341) are a table here instead of an if/elif chain.  Everything is plain Python/NumPy; this
layer stays on the host.
"""
from __future__ import annotations

import copy

import numpy as np

from . import model_obstacle as mo

_PI = np.pi


def _corridor(xmax, ymax=10):
    """The recurring corridor: wall half-spaces y >= ymax-1 and y <= 1 (two-vertex 'polygons', one edge each)."""
    top = [[xmax, ymax - 1], [0, ymax - 1]]
    bot = [[0, 1], [xmax, 1]]
    gtop = [[xmax, ymax - 1], [0, ymax - 1], [0, ymax], [xmax, ymax]]
    gbot = [[0, 1], [xmax, 1], [xmax, 0], [0, 0]]
    return top, bot, gtop, gbot


def _box(x0, y0, x1, y1, order="cw_from_ll"):
    """Axis-aligned box as 5 clockwise vertices starting at the lower-left corner."""
    return [[x0, y0], [x0, y1], [x1, y1], [x1, y0], [x0, y0]]


def _corridor_demo(xmax, start, goal, boxes, dyn, term):
    top, bot, gtop, gbot = _corridor(xmax)
    return dict(xL=[0, 0], xU=[xmax, 10], start=start, goal=goal, static=[top] + boxes + [bot],
                grid=[gtop] + boxes + [gbot], dyn=dyn, term=term)


def _dyn(cx, cy, th, l, w, v, ex, ey, eth, t0, t1):
    return [cx, cy, th, l, w, v, ex, ey, eth, t0, t1]


def _box_hi(x0, y0, x1, y1):
    """Box listed from its upper-right corner going down (the vertex order demo2-4 use)."""
    return [[x1, y1], [x1, y0], [x0, y0], [x0, y1], [x1, y1]]


_UP, _DOWN = _PI / 2, -_PI / 2

DEMOS = {
    # demo_setting.py:82-105
    "demo1": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [_box(10, 1, 15, 5)],
                                    [_dyn(22.5, 0, _UP, 3, 3, 0.2, 22.5, 9, _UP, 0, 55)], [[25, 39], [1, 9]]),
    # 106-128, 129-151, 152-174: same map, obstacle speed 0.2 / 0.15 / 0.1
    "demo2": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [_box_hi(20, 3, 25, 8)],
                                    [_dyn(18.5, 0, _UP, 3, 3, 0.2, 18.5, 9, _UP, 0, 55)], [[25, 39], [1, 9]]),
    "demo3": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [_box_hi(20, 3, 25, 8)],
                                    [_dyn(18.5, 0, _UP, 3, 3, 0.15, 18.5, 9, _UP, 0, 55)], [[25, 39], [1, 9]]),
    "demo4": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [_box_hi(20, 3, 25, 8)],
                                    [_dyn(18.5, 0, _UP, 3, 3, 0.1, 18.5, 9, _UP, 0, 55)], [[25, 39], [1, 9]]),
    # 175-197
    "demo5": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [_box(10, 1, 15, 5)],
                                    [_dyn(22.5, 0, _UP, 3, 3, 0.1, 22.5, 9, _UP, 0, 55)], [[25, 39], [1, 9]]),
    # 198-216, 217-235: empty corridor, two crossing boxes
    "demo6": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [],
                                    [_dyn(13.5, 0, _UP, 3, 3, 0.2, 13.5, 9, _UP, 0, 100),
                                     _dyn(22.5, 0, _UP, 3, 3, 0.1, 22.5, 9, _UP, 0, 200)], [[25, 39], [1, 9]]),
    "demo7": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [],
                                    [_dyn(13.5, 0, _UP, 3, 3, 0.1, 13.5, 9, _UP, 0, 100),
                                     _dyn(22.5, 0, _UP, 3, 3, 0.05, 22.5, 9, _UP, 0, 200)], [[28, 39], [1, 9]]),
    # 322-341
    "demo8": lambda: _corridor_demo(39, [3, 4, 0], [38, 4, 0], [],
                                    [_dyn(13.5, 0, _UP, 3, 3, 0.1, 13.5, 9, _UP, 0, 100),
                                     _dyn(22.5, 9, _DOWN, 3, 3, 0.1, 22.5, 0, _DOWN, 0, 200)], [[25, 39], [2, 6]]),
    # 270-297: the L-shaped yard
    "demo9": lambda: dict(xL=[0, 0], xU=[40, 60], start=[1, 5, 0], goal=[37, 58, _PI / 2],
                          static=[[[8, 0], [8, 6], [40, 6]],
                                  [[12, 30], [34, 30], [34, 14], [12, 14], [12, 30]],
                                  [[13, 49], [34, 49], [34, 34], [13, 34], [13, 49]],
                                  [[4, 60], [4, 10], [0, 10]],
                                  [[33, 60], [33, 55], [4, 55]]],
                          grid=[[[8, 6], [40, 6], [40, 0], [8, 0]],
                                [[12, 30], [34, 30], [34, 14], [12, 14]],
                                [[12, 50], [34, 50], [34, 34], [12, 34]],
                                [[0, 60], [4, 60], [4, 10], [0, 10]],
                                [[4, 60], [34, 60], [34, 54], [4, 54]]],
                          dyn=[_dyn(8, 50, _DOWN, 2, 2, 0.5, 8, 10, _DOWN, 0, 100)], term=[[34, 40], [54, 60]]),
    # 298-321
    "demo10": lambda: _corridor_demo(99, [3, 4, 0], [98, 4, 0], [],
                                     [_dyn(99, 5, -_PI, 3, 3, 0.5, 0, 5, -_PI, 0, 100)], [[60, 99], [1, 9]]),
    # 236-269
    "demo11": lambda: _corridor_demo(80, [3, 4, 0], [77, 4, 0], [],
                                     [_dyn(30.5, 0, _UP, 3, 3, 0.1, 30.5, 9, _UP, 0, 100),
                                      _dyn(39.5, 9, _DOWN, 3, 3, 0.1, 39.5, 0, _DOWN, 0, 200)], [[25, 39], [2, 6]]),
}


class problemSetting:
    def __init__(self, demo_name, custom=None):
        """``custom`` (dict with the keys of a DEMOS entry) defines a scenario that is not in the table."""
        self.demo_name = demo_name
        if custom is None:
            if demo_name not in DEMOS:
                raise KeyError("unknown demo %r (known: %s)" % (demo_name, ", ".join(sorted(DEMOS))))
            d = DEMOS[demo_name]()
        else:
            d = copy.deepcopy(custom)
        self.xL = list(d["xL"]); self.xU = list(d["xU"])
        self.map_size = [self.xU[0] - self.xL[0] + 1, self.xU[1] - self.xL[1] + 1]
        self.startPose = list(d["start"]); self.goalPose = list(d["goal"])
        self.static_lObs = d["static"]
        self.static_gridlObs = d["grid"]
        self.static_nObs = len(self.static_lObs)
        self.static_vObs = np.array([len(p) for p in self.static_lObs], dtype=int)
        self.terminal_set = np.asarray(d["term"], float)
        self.nObs = 0; self.vObs = []; self.lObs = []; self.obs_info = []
        self.add_dynamic_obstacle([list(r) for r in d["dyn"]])
        self.resolution = 1
        self.org_gridMap = mo.shape2grid(self.map_size, self.static_gridlObs, self.resolution)
        self.grid_map = self.org_gridMap
        self.senseDis = 10          # demo_setting.py:70

    def get_obstacle(self, center_x, center_y, theta, length, width):
        return mo.get_obstacle(center_x, center_y, theta, length, width)

    def add_dynamic_obstacle(self, dyn_obs_info):
        """Rectangles (length along the heading) from the info rows; demo_setting.py:374-403."""
        self.dyn_lObs = [mo.get_obstacle(r[0], r[1], r[2], r[3], r[4]) for r in dyn_obs_info]
        self.dyn_nObs = len(self.dyn_lObs)
        self.dyn_vObs = np.full(self.dyn_nObs, 5, dtype=int)
        self.dyn_obs_info = dyn_obs_info

    def combine_obstacle(self, dynObs_exist):
        """Static obstacles first (zero velocity rows), then the dynamic ones; demo_setting.py:431-455."""
        vObs = [int(v) for v in self.static_vObs]
        lObs = list(self.static_lObs)
        info = [[0] * 11 for _ in range(self.static_nObs)]
        nObs = self.static_nObs
        if dynObs_exist == 1:
            vObs += [int(v) for v in self.dyn_vObs]
            lObs += list(self.dyn_lObs)
            info += [list(r) for r in self.dyn_obs_info[:self.dyn_nObs]]
            nObs += self.dyn_nObs
        return nObs, vObs, lObs, info

    def rebuild_lObs(self, N, Ts, dynObs_exist):
        """Time-stack every polygon N+1 times (time-major, obstacle-minor); demo_setting.py:457-473."""
        nObs, vObs, lObs, info = self.combine_obstacle(dynObs_exist)
        self.nObs, self.vObs, self.obs_info = nObs, vObs, info
        # only the first nObs polygons are stacked; with a filtered info list the extra polygons are ignored (Q8)
        self.lObs = mo.rebuild_lObs(lObs[:nObs], vObs[:nObs], info, N, Ts)

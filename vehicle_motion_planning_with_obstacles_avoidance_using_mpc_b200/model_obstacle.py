"""Host-side obstacle builders: vertices -> half-space rows, rectangles, grid rasterisation, time-stacking.

Mirrors ``obstacleModel.obstacle_H_Represent`` (/root/reference/src/model_obstacle.py:37-102),
``problemSetting.get_obstacle`` / ``rebuild_lObs`` (demo_setting.py:405-429, 457-473) and
``mapModel.shape2grid`` (model_map.py:21-56).  Stays on the host.
"""
from __future__ import annotations

import numpy as np


class obstacleModel:
    def obstacle_H_Represent(self, nOb, vOb, obstacle_vertex):
        """Clockwise vertex lists -> stacked rows A (sum(vOb)-nOb, 2), b (.., 1) with A p <= b inside.

        Vertical edges give [+-1, 0], horizontal [0, +-1] (exact float equality picks the branch, as in
        the reference 63-76); slanted edges give the UNNORMALISED rows [-a, 1 | b] / [a, -1 | -b]
        (77-89)."""
        vOb = [int(v) for v in vOb]
        rows = sum(vOb) - nOb
        A = np.zeros((rows, 2)); b = np.zeros((rows, 1))
        r = 0
        for i in range(nOb):
            P = obstacle_vertex[i]
            for j in range(vOb[i] - 1):
                (x1, y1), (x2, y2) = P[j][:2], P[j + 1][:2]
                if x1 == x2:
                    if y2 < y1: A[r] = (1, 0); b[r] = x1
                    else: A[r] = (-1, 0); b[r] = -x1
                elif y1 == y2:
                    if x1 < x2: A[r] = (0, 1); b[r] = y1
                    else: A[r] = (0, -1); b[r] = -y1
                else:
                    a = (y2 - y1) / (x2 - x1)
                    c = y1 - a * x1
                    if x1 < x2: A[r] = (-a, 1); b[r] = c
                    else: A[r] = (a, -1); b[r] = -c
                r += 1
        return A, b


def get_obstacle(cx, cy, theta, length, width):
    """Rectangle -> 5 clockwise vertices (first repeated), demo_setting.py:405-429."""
    l = length / 2; w = width / 2
    c, s = np.cos(theta), np.sin(theta)
    v1 = [cx - l * c - w * s, cy - l * s + w * c]
    v2 = [cx + l * c - w * s, cy + l * s + w * c]
    v3 = [cx + l * c + w * s, cy + l * s - w * c]
    v4 = [cx - l * c + w * s, cy - l * s - w * c]
    return [v1, v2, v3, v4, v1]


def rebuild_lObs(lObs, vObs, obs_info, N, Ts):
    """Time-stack polygons N+1 times, time-major / obstacle-minor, translating polygon i by
    Ts*v_i*(cos th_i, sin th_i)*k (demo_setting.py:457-473). obs_info rows: [cx,cy,theta,l,w,v,...]."""
    out = []
    for k in range(N + 1):
        for i in range(len(lObs)):
            dx = Ts * obs_info[i][5] * np.cos(obs_info[i][2]) * k
            dy = Ts * obs_info[i][5] * np.sin(obs_info[i][2]) * k
            out.append([[lObs[i][j][0] + dx, lObs[i][j][1] + dy] for j in range(int(vObs[i]))])
    return out


def stacked_H_rep(lObs, vObs, obs_info, N, Ts):
    """closedLoop.update_obstacle_constraint (closed_loop.py:488-500): AObs ((N+1)R, 2), bObs ((N+1)R, 1)."""
    full = rebuild_lObs(lObs, vObs, obs_info, N, Ts)
    fv = [len(p) for p in full]
    return obstacleModel().obstacle_H_Represent(len(full), fv, full)


def shape2grid(map_size, polygons, resolution=1):
    """Occupancy grid (H, W) with each polygon's AABB filled (model_map.py:16-56)."""
    g = np.zeros((int((map_size[1] - 1) / resolution) + 1, int((map_size[0] - 1) / resolution) + 1))
    for poly in polygons:
        xs = [p[0] / resolution for p in poly]; ys = [p[1] / resolution for p in poly]
        x0, x1, y0, y1 = min(xs), max(xs), min(ys), max(ys)
        for i in range(int(x1 - x0) + 1):
            for j in range(int(y1 - y0) + 1):
                yy, xx = int(y0) + j, int(x0) + i
                if 0 <= yy < g.shape[0] and 0 <= xx < g.shape[1]:
                    g[yy, xx] = 1
    return g

"""Host-side global planner: 8-connected grid A* producing the (x, y, yaw) reference path.

Mirrors the call surface of the reference's ``class a_star`` (/root/reference/src/a_star.py:16-200):
``a_star(grid, start, goal).solve(grid, start, goal)`` -> route (goal -> first cell after start, start
excluded, a_star.py:56-61), ``rebuild_path`` (reverse + (row,col)->(x,y), 137-147) and
``create_reference_path`` (yaw = atan2 to the next point, last yaw copied, 189-200).

Behavioural contract reproduced exactly (it decides which of several equal-cost routes is returned, hence
``xref``): neighbour order (a_star.py:20), Euclidean step cost and heuristic (35-37, 69), heap entries
``(f, (row, col))`` with tuple tie-breaking (32, 100), no decrease-key (stale entries stay in the heap and
count as "open", 93), closed nodes re-opened only on a strictly better g (90-93), and the
``gscore.get(n, 0)`` default (90, 93).  The linear heap scan of the reference (93) is replaced by a
multiset counter - same truth value, O(1).
"""
from __future__ import annotations

import heapq
import math
from collections import Counter

import numpy as np

_NEIGHBORS = ((0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (1, -1), (-1, 1), (-1, -1))


def _h(a, b):
    return math.sqrt((b[0] - a[0]) ** 2 + (b[1] - a[1]) ** 2)


class a_star:
    def __init__(self, array=None, start=None, goal=None):
        self.neighbors = list(_NEIGHBORS)
        self._reset(start, goal)

    def _reset(self, start, goal):
        self.close_set = set()
        self.came_from = {}
        self.gscore = {}
        self.fscore = {}
        self.oheap = []
        self._open = Counter()
        if start is not None:
            start = (int(start[0]), int(start[1]))
            self.gscore[start] = 0
            self.fscore[start] = _h(start, goal)
            heapq.heappush(self.oheap, (self.fscore[start], start))
            self._open[start] += 1

    def heuristic(self, a, b):
        return _h(a, b)

    def solve(self, array, start, goal):
        """Route as a list of (row, col) from goal back to the first cell after start; False if none.
        Unlike the reference object (single use: its open set is built in __init__ for one start), a solved
        or fresh instance may be re-used - state is rebuilt when needed."""
        start = (int(start[0]), int(start[1])); goal = (int(goal[0]), int(goal[1]))
        if self.close_set or not self.oheap or self.oheap[0][1] != start:
            self._reset(start, goal)
        grid = np.asarray(array)
        H, W = grid.shape
        occ = (grid == 1)
        while self.oheap:
            current = heapq.heappop(self.oheap)[1]
            self._open[current] -= 1
            if current == goal:
                data = []
                while current in self.came_from:
                    data.append(current)
                    current = self.came_from[current]
                return data
            self.close_set.add(current)
            gc = self.gscore[current]
            for i, j in _NEIGHBORS:
                nb = (current[0] + i, current[1] + j)
                if not (0 <= nb[0] < H and 0 <= nb[1] < W) or occ[nb[0], nb[1]]:
                    continue
                tg = gc + _h(current, nb)
                gn = self.gscore.get(nb, 0)
                if nb in self.close_set and tg >= gn:
                    continue
                if tg < gn or self._open[nb] <= 0:
                    self.came_from[nb] = current
                    self.gscore[nb] = tg
                    self.fscore[nb] = tg + _h(nb, goal)
                    heapq.heappush(self.oheap, (self.fscore[nb], nb))
                    self._open[nb] += 1
        return False

    def rebuild_path(self, route):
        route = np.asarray(route)
        n = route.shape[0]
        return [[route[n - 1 - i][1], route[n - 1 - i][0]] for i in range(n)]

    def create_reference_path(self, path):
        n = len(path)
        ref = []
        for i in range(n - 1):
            yaw = np.arctan2(path[i + 1][1] - path[i][1], path[i + 1][0] - path[i][0])
            ref.append([path[i][0], path[i][1], yaw])
        ref.append([path[n - 1][0], path[n - 1][1], ref[-1][2]])
        return ref

    def demo_data(self):
        """Known-answer grid of the reference (a_star.py:202-232): rows given as run-lengths of obstacles."""
        g = np.zeros((11, 20))
        for r in (0, 1, 2):
            g[r, 6:9] = 1; g[r, 10:13] = 1
        for r in (3, 4, 5, 6, 7, 8):
            g[r, 8] = 1
        g[5, 3:6] = 1
        g[6, 13:18] = 1; g[6, 19] = 1
        g[7, 13:16] = 1; g[8, 13:16] = 1
        g[9, 8:14] = 1
        return g, (0, 0), (0, 19)


def plan_reference(grid, start_pose, goal_pose):
    """closedLoop.update_path(type='A_star') (closed_loop.py:555-563): (3, M) reference or None."""
    start = (int(start_pose[1]), int(start_pose[0])); goal = (int(goal_pose[1]), int(goal_pose[0]))
    pl = a_star(grid, start, goal)
    route = pl.solve(grid, start, goal)
    if route is False or len(route) < 2:
        return None
    return np.asarray(pl.create_reference_path(pl.rebuild_path(route)), float).T


def plan_batch(grids, starts, goals, grid_index=None, max_len=None, threads=0):
    """Many ``plan_reference`` queries at once through the native planner (``obca_b200_astar_batch``,
    include/obca_b200.h): same routes as ``a_star.solve`` cell for cell, on all host cores.

    ``grids`` (H,W) or (G,H,W) occupancy (1 = occupied); ``starts``/``goals`` (n,3) poses ``[x, y, yaw]`` (or one
    goal for all); ``grid_index`` (n,) picks the grid of each query.  Returns ``ref`` (n, max_len, 3) and ``ref_len``
    (n,), 0 where ``plan_reference`` would return None."""
    import ctypes as C
    from . import _lib
    g = np.asarray(grids)
    if g.ndim == 2:
        g = g[None]
    occ = np.ascontiguousarray(g == 1, dtype=np.uint8)
    G, H, W = occ.shape
    starts = np.atleast_2d(np.asarray(starts, float)); n = starts.shape[0]
    goals = np.broadcast_to(np.atleast_2d(np.asarray(goals, float)), (n, np.atleast_2d(goals).shape[1]))
    s_rc = np.ascontiguousarray(np.stack([starts[:, 1], starts[:, 0]], 1).astype(np.int32))
    g_rc = np.ascontiguousarray(np.stack([goals[:, 1], goals[:, 0]], 1).astype(np.int32))
    gi = None if grid_index is None else np.ascontiguousarray(grid_index, dtype=np.int32)
    max_len = int(max_len or H * W)
    ref = np.zeros((n, max_len, 3)); ref_len = np.zeros(n, np.int32)
    ip = C.POINTER(C.c_int32)
    _lib.check(_lib.lib().obca_b200_astar_batch(
        n, occ.ctypes.data, G, H, W, None if gi is None else gi.ctypes.data_as(ip), s_rc.ctypes.data_as(ip),
        g_rc.ctypes.data_as(ip), max_len, ref.ctypes.data, ref_len.ctypes.data_as(ip), int(threads)))
    return ref, ref_len


def reference_windows(ref, ref_len, x0, N, path_index=None):
    """``update_reference_trajectory`` (closed_loop.py:502-528) for n poses: (n, N+1, 3) windows
    (``obca_b200_reference_windows``)."""
    import ctypes as C
    from . import _lib
    ref = np.ascontiguousarray(ref, float); ref_len = np.ascontiguousarray(ref_len, np.int32)
    x0 = np.ascontiguousarray(np.atleast_2d(x0), float); n = x0.shape[0]
    pi = None if path_index is None else np.ascontiguousarray(path_index, dtype=np.int32)
    out = np.empty((n, N + 1, 3))
    ip = C.POINTER(C.c_int32)
    _lib.check(_lib.lib().obca_b200_reference_windows(
        n, ref.ctypes.data, ref_len.ctypes.data_as(ip), ref.shape[1], None if pi is None else pi.ctypes.data_as(ip),
        x0.ctypes.data, int(N), out.ctypes.data))
    return out

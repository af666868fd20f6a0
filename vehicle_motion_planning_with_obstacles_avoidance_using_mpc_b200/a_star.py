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

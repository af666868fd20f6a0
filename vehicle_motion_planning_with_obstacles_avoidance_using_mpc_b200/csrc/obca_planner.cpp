// Batched host planner: 8-connected grid A* -> (x, y, yaw) reference paths, many queries on all host cores.
//
// Reproduces the reference planner's choices exactly, because they decide which of several equal-cost routes is
// returned and therefore xref (reference: src/a_star.py):
//   * neighbour order (a_star.py:20), Euclidean step cost and heuristic (35-37, 69);
//   * heap entries ordered by (f, row, col) (32, 100): a total order, so any binary heap pops the same sequence;
//   * no decrease-key: stale entries stay in the heap and still count as "open" (93);
//   * an unseen cell reads g = 0 (gscore.get(n, 0), 90 and 93), closed cells re-open only on a strictly better g;
//   * route = goal back to the first cell after the start (56-61); reversed and swapped to (x, y) (137-147);
//     yaw_i = atan2 towards the next point, the last yaw copied (189-200).
// The reference's linear scan of the heap for membership (93) is a per-cell counter here.
#include <stdint.h>
#include <math.h>
#include <atomic>
#include <thread>
#include <vector>
#include <algorithm>

#include "../../include/obca_b200.h"

namespace {

struct HeapItem { double f; int32_t r, c; };
struct HeapAfter {      // std::push_heap keeps the largest on top: "a after b" makes it a min-heap on (f, r, c)
  bool operator()(const HeapItem& a, const HeapItem& b) const {
    if (a.f != b.f) return a.f > b.f;
    if (a.r != b.r) return a.r > b.r;
    return a.c > b.c;
  }
};

struct Planner {
  int H, W;
  std::vector<double> g;
  std::vector<int32_t> parent, open_n;
  std::vector<uint8_t> seen, closed;
  std::vector<HeapItem> heap;
  std::vector<int32_t> route;

  Planner(int h, int w) : H(h), W(w), g(h * w), parent(h * w), open_n(h * w), seen(h * w), closed(h * w) {}

  static double dist(int r0, int c0, int r1, int c1) {
    return sqrt((double)((r1 - r0) * (r1 - r0) + (c1 - c0) * (c1 - c0)));
  }

  // returns the number of cells of the route (goal ... first after start) left in `route`, 0 if none
  int solve(const uint8_t* occ, int sr, int sc, int gr, int gc) {
    static const int dr[8] = {0, 0, 1, -1, 1, 1, -1, -1};
    static const int dc[8] = {1, -1, 0, 0, 1, -1, 1, -1};
    std::fill(seen.begin(), seen.end(), 0); std::fill(closed.begin(), closed.end(), 0);
    std::fill(open_n.begin(), open_n.end(), 0); std::fill(parent.begin(), parent.end(), -1);
    heap.clear(); route.clear();
    HeapAfter after;
    int s = sr * W + sc;
    g[s] = 0.0; seen[s] = 1; open_n[s] = 1;
    heap.push_back({dist(sr, sc, gr, gc), sr, sc});
    while (!heap.empty()) {
      std::pop_heap(heap.begin(), heap.end(), after);
      HeapItem cur = heap.back(); heap.pop_back();
      int ci = cur.r * W + cur.c;
      open_n[ci] -= 1;
      if (cur.r == gr && cur.c == gc) {
        for (int p = ci; parent[p] >= 0; p = parent[p]) route.push_back(p);
        return (int)route.size();
      }
      closed[ci] = 1;
      double gcur = g[ci];
      for (int n = 0; n < 8; ++n) {
        int r = cur.r + dr[n], c = cur.c + dc[n];
        if (r < 0 || r >= H || c < 0 || c >= W) continue;
        int ni = r * W + c;
        if (occ[ni] == 1) continue;
        double tg = gcur + dist(cur.r, cur.c, r, c);
        double gn = seen[ni] ? g[ni] : 0.0;
        if (closed[ni] && tg >= gn) continue;
        if (tg < gn || open_n[ni] <= 0) {
          parent[ni] = ci; g[ni] = tg; seen[ni] = 1;
          heap.push_back({tg + dist(r, c, gr, gc), r, c});
          std::push_heap(heap.begin(), heap.end(), after);
          open_n[ni] += 1;
        }
      }
    }
    return 0;
  }
};

}  // namespace

extern "C" int obca_b200_astar_batch(int n, const uint8_t* grids, int n_grids, int H, int W, const int32_t* grid_index,
                                     const int32_t* start_rc, const int32_t* goal_rc, int max_len, double* ref,
                                     int32_t* ref_len, int n_threads) {
  if (n < 0 || !grids || n_grids < 1 || H < 1 || W < 1 || !start_rc || !goal_rc || max_len < 2 || !ref || !ref_len)
    return OBCA_E_ARG;
  for (int q = 0; q < n; ++q) {
    int gi = grid_index ? grid_index[q] : 0;
    if (gi < 0 || gi >= n_grids) return OBCA_E_ARG;
    const int32_t* s = start_rc + 2 * q; const int32_t* t = goal_rc + 2 * q;
    if (s[0] < 0 || s[0] >= H || s[1] < 0 || s[1] >= W || t[0] < 0 || t[0] >= H || t[1] < 0 || t[1] >= W) return OBCA_E_ARG;
  }
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  n_threads = std::max(1, std::min(n_threads, std::max(1, n / 4)));
  std::atomic<int> next(0), overflow(0);
  auto work = [&]() {
    Planner pl(H, W);
    for (;;) {
      int q = next.fetch_add(1);
      if (q >= n) break;
      const uint8_t* occ = grids + (size_t)(grid_index ? grid_index[q] : 0) * H * W;
      int len = pl.solve(occ, start_rc[2 * q], start_rc[2 * q + 1], goal_rc[2 * q], goal_rc[2 * q + 1]);
      double* out = ref + (size_t)q * max_len * 3;
      if (len < 2) { ref_len[q] = 0; continue; }             // closed_loop.py:555-563 needs two points for a yaw
      if (len > max_len) { ref_len[q] = -len; overflow.store(1); continue; }
      for (int i = 0; i < len; ++i) {                         // reversed: first cell after the start comes first
        int cell = pl.route[len - 1 - i];
        out[3 * i] = (double)(cell % W); out[3 * i + 1] = (double)(cell / W);
      }
      for (int i = 0; i + 1 < len; ++i) out[3 * i + 2] = atan2(out[3 * i + 4] - out[3 * i + 1], out[3 * i + 3] - out[3 * i]);
      out[3 * (len - 1) + 2] = out[3 * (len - 2) + 2];
      ref_len[q] = len;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return overflow.load() ? OBCA_E_SIZE : OBCA_OK;
}

// closedLoop.update_reference_trajectory (closed_loop.py:502-528) for n poses: first closest path point, N + 1
// consecutive points clamped to the last.  ref is [n_paths, max_len, 3]; path_index NULL = pose q uses path q.
extern "C" int obca_b200_reference_windows(int n, const double* ref, const int32_t* ref_len, int max_len,
                                           const int32_t* path_index, const double* x0, int N, double* xref) {
  if (n < 0 || !ref || !ref_len || !x0 || !xref || N < 1 || max_len < 1) return OBCA_E_ARG;
  for (int q = 0; q < n; ++q) {
    int p = path_index ? path_index[q] : q;
    int M = ref_len[p];
    if (M < 1 || M > max_len) return OBCA_E_ARG;
    const double* r = ref + (size_t)p * max_len * 3;
    double best = 0.0; int i0 = 0;
    for (int i = 0; i < M; ++i) {
      double dx = x0[3 * q] - r[3 * i], dy = x0[3 * q + 1] - r[3 * i + 1];
      double d = dx * dx + dy * dy;
      if (i == 0 || d < best) { best = d; i0 = i; }
    }
    double* o = xref + (size_t)q * (N + 1) * 3;
    for (int k = 0; k <= N; ++k) {
      int i = std::min(i0 + k, M - 1);
      o[3 * k] = r[3 * i]; o[3 * k + 1] = r[3 * i + 1]; o[3 * k + 2] = r[3 * i + 2];
    }
  }
  return OBCA_OK;
}

// obca_loop.cu - the callers' side of the solve on the device:
//   * obca_b200_build_rows: vertices + velocities -> half-space rows (A, b0, db), the compact form of the reference's
//     time-stacked H-representation (src/model_obstacle.py:37-102 over src/demo_setting.py:457-473);
//   * obca_loop_*: closedLoop.closed_loop_mpc4 (src/closed_loop.py:323-441) for B scenarios in lock-step with all state
//     resident in HBM: per step one kernel builds every scenario's inputs (goal test, obstacle propagation 445-486,
//     lidar gate 591-630, reference window 502-528, terminal set 371, obstacle rows) and the work lists of the three
//     solver modes, the solver runs over those lists (obca_b200_solve_indexed), one kernel applies the first input.
//     Nothing crosses PCIe between reset and read.
// Arithmetic that decides a branch in the reference (exact == tests of the H-representation, the strict '<' of the
// closest-point scan) is written with explicitly rounded operations so that no FMA contraction changes the outcome.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/obca_b200.h"

namespace {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// one edge (x1,y1) -> (x2,y2) of a clockwise polygon: model_obstacle.py:63-89 (rows are not normalised)
__device__ __forceinline__ void edge_row(double x1, double y1, double x2, double y2, double& a0, double& a1, double& b) {
  if (x1 == x2) {
    if (y2 < y1) { a0 = 1.0; a1 = 0.0; b = x1; } else { a0 = -1.0; a1 = 0.0; b = -x1; }
  } else if (y1 == y2) {
    if (x1 < x2) { a0 = 0.0; a1 = 1.0; b = y1; } else { a0 = 0.0; a1 = -1.0; b = -y1; }
  } else {
    const double a = __ddiv_rn(sub(y2, y1), sub(x2, x1));
    const double c = sub(y1, mul(a, x1));
    if (x1 < x2) { a0 = -a; a1 = 1.0; b = c; } else { a0 = a; a1 = -1.0; b = -c; }
  }
}

struct RowsArgs {
  int32_t eptr[OBCA_MAX_OBS + 1];
  int batch, n_obs, rows;
  const double *verts, *vel, *Ts;
  double ts;
  double *A, *b0, *db;
};

__global__ void build_rows_kernel(const RowsArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)a.batch * a.rows) return;
  const int b = (int)(t / a.rows), r = (int)(t % a.rows);
  int i = 0;
  while (a.eptr[i + 1] <= r) ++i;
  const double* v = a.verts + ((size_t)b * (a.rows + a.n_obs) + r + i) * 2;   // polygon i starts at vertex eptr[i] + i
  double a0, a1, bb;
  edge_row(v[0], v[1], v[2], v[3], a0, a1, bb);
  a.A[2 * t] = a0; a.A[2 * t + 1] = a1; a.b0[t] = bb;
  if (a.db) {
    double d = 0.0;
    if (a.vel) {
      const double* w = a.vel + ((size_t)b * a.n_obs + i) * 3;                  // speed, cos(heading), sin(heading)
      const double ts = a.Ts ? a.Ts[b] : a.ts;
      d = mul(mul(ts, w[0]), add(mul(a0, w[1]), mul(a1, w[2])));
    }
    a.db[t] = d;
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct LoopDev {
  int B, N, S1, M, Rs, R, max_steps, terminal_rule;
  double sense, goal[2], goal_tol, ego0, uU0;
  // scenario state
  double *x0, *u0, *Ts, *Ts_opt, *dyn, *dcs, *xprev, *traj;
  int32_t *alive, *failed, *steps, *mode_log, *cur;
  // solver inputs / outputs (indexed by scenario)
  double *xref, *Tmax, *term, *A, *b0, *db, *x, *u, *T;
  int32_t* status;
  // speculative fallback (lp.speculative): the solve without the terminal set runs beside the one with it, on every
  // detected scenario, into its own result arrays; loop_fallback takes them where the terminal-set solve failed
  int speculative;
  double *x2, *u2;
  int32_t* status2;
  // shared constants
  const double *path, *A_s, *b_s;
  // work lists
  int32_t *idx_free, *idx_set, *idx_fall, *counts;
  unsigned long long* totals;
};

__global__ void loop_prepare(const LoopDev L, int k) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= L.B) return;
  L.cur[b] = -1;
  double* x0 = L.x0 + 3 * b;
  {
    const double dx = x0[0] - L.goal[0], dy = x0[1] - L.goal[1];
    if (!(add(mul(dx, dx), mul(dy, dy)) >= L.goal_tol)) L.alive[b] = 0;
  }
  if (!L.alive[b]) return;
  double* dy_ = L.dyn + 7 * b;
  const double c = L.dcs[2 * b], s = L.dcs[2 * b + 1];
  const double tso = L.Ts_opt[b];
  if ((double)k > dy_[6]) {                                                     // update_obstacle: closed_loop.py:468-471
    dy_[0] = add(dy_[0], mul(mul(tso, dy_[5]), c));
    dy_[1] = add(dy_[1], mul(mul(tso, dy_[5]), s));
  }
  const bool live = (double)k >= dy_[6];
  // rectangle -> vertices (demo_setting.py:405-429)
  const double cx = dy_[0], cy = dy_[1], l = dy_[3] / 2, w = dy_[4] / 2;
  const double lc = mul(l, c), ls = mul(l, s), wc = mul(w, c), ws = mul(w, s);
  double vx[5], vy[5];
  vx[0] = sub(sub(cx, lc), ws); vy[0] = add(sub(cy, ls), wc);
  vx[1] = sub(add(cx, lc), ws); vy[1] = add(add(cy, ls), wc);
  vx[2] = add(add(cx, lc), ws); vy[2] = sub(add(cy, ls), wc);
  vx[3] = add(sub(cx, lc), ws); vy[3] = sub(sub(cy, ls), wc);
  vx[4] = vx[0]; vy[4] = vy[0];
  // sensor: car-front midpoint to the four vertices (closed_loop.py:601-618)
  const double fx = x0[0] + L.ego0 * cos(x0[2]), fy = x0[1] + L.ego0 * sin(x0[2]);
  double dist = hypot(fx - vx[0], fy - vy[0]);
  for (int j = 1; j < 4; ++j) dist = fmin(dist, hypot(fx - vx[j], fy - vy[j]));
  const bool fix = live && dist <= L.sense && k > 0;
  // reference window (closed_loop.py:502-528): first closest path point, clamped to the last
  int i0 = 0;
  {
    double best = 0.0;
    for (int i = 0; i < L.M; ++i) {
      const double dx = sub(x0[0], L.path[3 * i]), dy = sub(x0[1], L.path[3 * i + 1]);
      const double d = add(mul(dx, dx), mul(dy, dy));
      if (i == 0 || d < best) { best = d; i0 = i; }
    }
  }
  double* xr = L.xref + (size_t)b * L.S1 * 3;
  for (int q = 0; q <= L.N; ++q) {
    const int i = min(i0 + q, L.M - 1);
    xr[3 * q] = L.path[3 * i]; xr[3 * q + 1] = L.path[3 * i + 1]; xr[3 * q + 2] = L.path[3 * i + 2];
  }
  if (!fix) {
    L.Tmax[b] = add(__ddiv_rn(add(sub(xr[3 * L.N], x0[0]), sub(xr[3 * L.N + 1], x0[1])), mul(mul((double)L.N, L.uU0), L.Ts[b])), 1.0);
    L.cur[b] = OBCA_MODE_FREE;
    L.idx_free[atomicAdd(&L.counts[0], 1)] = b;
    return;
  }
  // fixed-time phase: keep the tail of the previous plan (closed_loop.py:362-363), headings from the path segments
  const double* xp = L.xprev + (size_t)b * L.S1 * 3;
  for (int q = 0; q < L.N - 5; ++q) { xr[3 * q] = xp[3 * (q + 1)]; xr[3 * q + 1] = xp[3 * (q + 1) + 1]; xr[3 * q + 2] = xp[3 * (q + 1) + 2]; }
  for (int q = 0; q < L.N; ++q) xr[3 * q + 2] = atan2(xr[3 * (q + 1) + 1] - xr[3 * q + 1], xr[3 * (q + 1)] - xr[3 * q]);
  xr[3 * L.N + 2] = xr[3 * (L.N - 1) + 2];
  L.Ts[b] = tso;                                                                 // the inherited step (closed_loop.py:586-587)
  double* tm = L.term + 3 * b;
  if (L.terminal_rule == 0) { tm[0] = x0[0] + 5; tm[1] = 1.0; tm[2] = 9.0; }     // closed_loop.py:371
  else { tm[0] = 5.0; tm[1] = x0[1] + 4; tm[2] = 60.0; }                         // simulation.py:72
  double* A = L.A + (size_t)b * L.R * 2; double* b0 = L.b0 + (size_t)b * L.R; double* db = L.db + (size_t)b * L.R;
  for (int r = 0; r < L.Rs; ++r) { A[2 * r] = L.A_s[2 * r]; A[2 * r + 1] = L.A_s[2 * r + 1]; b0[r] = L.b_s[r]; db[r] = 0.0; }
  const double tv = mul(tso, dy_[5]);
  for (int j = 0; j < 4; ++j) {
    double a0, a1, bb;
    edge_row(vx[j], vy[j], vx[j + 1], vy[j + 1], a0, a1, bb);
    const int r = L.Rs + j;
    A[2 * r] = a0; A[2 * r + 1] = a1; b0[r] = bb;
    db[r] = mul(tv, add(mul(a0, c), mul(a1, s)));
  }
  L.cur[b] = OBCA_MODE_FIXED_SET;
  L.idx_set[atomicAdd(&L.counts[1], 1)] = b;
  if (L.speculative) L.idx_fall[atomicAdd(&L.counts[2], 1)] = b;
}

// failures of the terminal-set solve go to the solve without it (closed_loop.py:389-395)
__global__ void loop_fallback(const LoopDev L) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= L.B) return;
  if (L.cur[b] == OBCA_MODE_FIXED_SET && L.status[b] < 0) {
    L.cur[b] = OBCA_MODE_FIXED_NOTERM;
    if (!L.speculative) { L.idx_fall[atomicAdd(&L.counts[2], 1)] = b; return; }
    atomicAdd(&L.counts[3], 1);                                   // fallback results actually used
    L.status[b] = L.status2[b];
    const size_t nx = (size_t)L.S1 * 3, nu = (size_t)L.N * 2;
    for (size_t q = 0; q < nx; ++q) L.x[b * nx + q] = L.x2[b * nx + q];
    for (size_t q = 0; q < nu; ++q) L.u[b * nu + q] = L.u2[b * nu + q];
  }
}

// apply the first input / take the predicted state (closed_loop.py:416-419)
__global__ void loop_advance(const LoopDev L, int k) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) {
    L.totals[0] += (unsigned long long)L.counts[0]; L.totals[1] += (unsigned long long)L.counts[1];
    L.totals[2] += (unsigned long long)L.counts[L.speculative ? 3 : 2];
  }
  if (b >= L.B) return;
  const int mode = L.cur[b];
  if (mode < 0) return;
  L.mode_log[(size_t)b * L.max_steps + k] = mode;
  if (L.status[b] < 0) { L.failed[b] = 1; L.alive[b] = 0; return; }
  L.Ts_opt[b] = (mode == OBCA_MODE_FREE) ? mul(L.T[b], L.Ts[b]) : L.Ts[b];
  const double* x = L.x + (size_t)b * L.S1 * 3;
  const double* u = L.u + (size_t)b * L.N * 2;
  double* xp = L.xprev + (size_t)b * L.S1 * 3;
  for (int q = 0; q < 3 * L.S1; ++q) xp[q] = x[q];
  L.u0[2 * b] = u[0]; L.u0[2 * b + 1] = u[1];
  double* tr = L.traj + ((size_t)b * (L.max_steps + 1) + k + 1) * 3;
  for (int j = 0; j < 3; ++j) { L.x0[3 * b + j] = x[3 + j]; tr[j] = x[3 + j]; }
  L.steps[b] += 1;
}

__global__ void loop_reset(const LoopDev L, double sx, double sy, double sth, double ts0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= L.B) return;
  const double nan_ = __longlong_as_double(0x7ff8000000000000LL);
  L.x0[3 * b] = sx; L.x0[3 * b + 1] = sy; L.x0[3 * b + 2] = sth;
  L.u0[2 * b] = 0.0; L.u0[2 * b + 1] = 0.0;
  L.Ts[b] = ts0; L.Ts_opt[b] = ts0;
  L.alive[b] = 1; L.failed[b] = 0; L.steps[b] = 0; L.cur[b] = -1; L.status[b] = 0;
  for (int q = 0; q < 3 * L.S1; ++q) L.xprev[(size_t)b * L.S1 * 3 + q] = 0.0;
  double* tr = L.traj + (size_t)b * (L.max_steps + 1) * 3;
  tr[0] = sx; tr[1] = sy; tr[2] = sth;
  for (int q = 3; q < 3 * (L.max_steps + 1); ++q) tr[q] = nan_;
  for (int q = 0; q < L.max_steps; ++q) L.mode_log[(size_t)b * L.max_steps + q] = -1;
  if (b == 0) for (int m = 0; m < 3; ++m) L.totals[m] = 0ull;
}

}  // namespace

struct obca_loop {
  int device;
  obca_loop_params lp;
  LoopDev d;
  obca_ctx* ctx[3];            // FREE, FIXED_SET, FIXED_NOTERM
  int32_t eptr_free[OBCA_MAX_OBS + 1], eptr_fix[OBCA_MAX_OBS + 1];
  double *lam, *mu, *lam_free, *mu_free, *obj;
  int32_t* iters;
  void* arena;
  int k;                       // steps issued since the last reset
  cudaStream_t side;           // the free-time solve of a step runs here, beside the fixed-time chain
  cudaStream_t side2;          // speculative fallback solve
  cudaEvent_t ev_ready, ev_free, ev_spec;
  double *lam2, *mu2, *obj2, *T2;
  int32_t* iters2;
};

extern "C" {

int obca_b200_build_rows(int batch, int n_obs, const int32_t* edge_ptr, const double* verts, const double* vel,
                         const double* Ts_inst, double ts, double* A, double* b0, double* db, void* cuda_stream) {
  if (batch < 0 || n_obs < 1 || n_obs > OBCA_MAX_OBS || !edge_ptr || !verts || !A || !b0) return OBCA_E_ARG;
  if (batch == 0) return OBCA_OK;
  RowsArgs a;
  memset(&a, 0, sizeof(a));
  if (edge_ptr[0] != 0) return OBCA_E_ARG;
  for (int i = 0; i <= n_obs; ++i) {
    if (i > 0 && edge_ptr[i] <= edge_ptr[i - 1]) return OBCA_E_ARG;
    a.eptr[i] = edge_ptr[i];
  }
  a.batch = batch; a.n_obs = n_obs; a.rows = edge_ptr[n_obs];
  a.verts = verts; a.vel = vel; a.Ts = Ts_inst; a.ts = ts; a.A = A; a.b0 = b0; a.db = db;
  const long long total = (long long)batch * a.rows;
  build_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? OBCA_OK : OBCA_E_CUDA;
}

int obca_b200_loop_destroy(obca_loop* l) {
  if (!l) return OBCA_E_ARG;
  for (int m = 0; m < 3; ++m) if (l->ctx[m]) obca_b200_destroy(l->ctx[m]);
  if (l->side) cudaStreamDestroy(l->side);
  if (l->side2) cudaStreamDestroy(l->side2);
  if (l->ev_spec) cudaEventDestroy(l->ev_spec);
  if (l->ev_ready) cudaEventDestroy(l->ev_ready);
  if (l->ev_free) cudaEventDestroy(l->ev_free);
  if (l->arena) cudaFree(l->arena);
  free(l);
  return OBCA_OK;
}

int obca_b200_loop_create(obca_loop** out, int device, int n_scenarios, const obca_loop_params* lp,
                          const obca_params* p_free, const obca_params* p_set, const obca_params* p_noterm,
                          const int32_t* edges_static, const double* A_static, const double* b_static,
                          const double* path) {
  if (!out || !lp || !p_free || !p_set || !p_noterm || !edges_static || !A_static || !b_static || !path || n_scenarios < 1)
    return OBCA_E_ARG;
  const int N = lp->N, ns = lp->n_static, Rs = lp->rows_static, M = lp->path_len, B = n_scenarios;
  if (N < 5 || N + 1 > OBCA_MAX_STAGES || ns < 0 || ns + 1 > OBCA_MAX_OBS || Rs + 4 > OBCA_MAX_ROWS || M < 1 || lp->max_steps < 1)
    return OBCA_E_ARG;
  if (p_free->N != N || p_set->N != N || p_noterm->N != N || p_free->n_obs != ns || p_free->rows != Rs ||
      p_set->n_obs != ns + 1 || p_set->rows != Rs + 4 || p_noterm->n_obs != ns + 1 || p_noterm->rows != Rs + 4 ||
      p_free->mode != OBCA_MODE_FREE || p_set->mode != OBCA_MODE_FIXED_SET || p_noterm->mode != OBCA_MODE_FIXED_NOTERM)
    return OBCA_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return OBCA_E_NODEVICE; }
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return OBCA_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return OBCA_E_CUDA;
  obca_loop* l = (obca_loop*)calloc(1, sizeof(obca_loop));
  if (!l) return OBCA_E_NOMEM;
  l->device = device; l->lp = *lp;
  int acc = 0;
  l->eptr_free[0] = 0; l->eptr_fix[0] = 0;
  for (int i = 0; i < ns; ++i) { acc += edges_static[i]; l->eptr_free[i + 1] = acc; l->eptr_fix[i + 1] = acc; }
  if (acc != Rs) { free(l); return OBCA_E_ARG; }
  l->eptr_fix[ns + 1] = Rs + 4;
  const obca_params* ps[3] = {p_free, p_set, p_noterm};
  for (int m = 0; m < 3; ++m) {
    int rc = obca_b200_create(&l->ctx[m], device, B, ps[m]);
    if (rc != OBCA_OK) { obca_b200_loop_destroy(l); return rc; }
  }
  // one arena, 256-byte aligned slices
  const size_t S1 = N + 1, R = Rs + 4, no = ns + 1, T1 = lp->max_steps + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_x0 = take(B * 3 * 8), o_u0 = take(B * 2 * 8), o_Ts = take(B * 8), o_Tso = take(B * 8), o_dyn = take(B * 7 * 8),
               o_dcs = take(B * 2 * 8), o_xprev = take(B * S1 * 3 * 8), o_traj = take(B * T1 * 3 * 8), o_alive = take(B * 4),
               o_failed = take(B * 4), o_steps = take(B * 4), o_mlog = take((size_t)B * lp->max_steps * 4), o_cur = take(B * 4),
               o_xref = take(B * S1 * 3 * 8), o_Tmax = take(B * 8), o_term = take(B * 3 * 8), o_A = take(B * R * 2 * 8),
               o_b0 = take(B * R * 8), o_db = take(B * R * 8), o_x = take(B * S1 * 3 * 8), o_u = take((size_t)B * N * 2 * 8),
               o_T = take(B * 8), o_status = take(B * 4), o_path = take((size_t)M * 3 * 8), o_As = take((size_t)(Rs + 1) * 2 * 8),
               o_bs = take((size_t)(Rs + 1) * 8), o_i0 = take(B * 4), o_i1 = take(B * 4), o_i2 = take(B * 4), o_cnt = take(16),
               o_tot = take(32), o_lam = take(B * S1 * R * 8), o_mu = take(B * S1 * 4 * no * 8), o_lamf = take(B * S1 * R * 8), o_lam2 = take(B * S1 * R * 8), o_mu2 = take(B * S1 * 4 * no * 8), o_x2 = take(B * S1 * 3 * 8),
               o_u2 = take((size_t)B * N * 2 * 8), o_st2 = take(B * 4), o_obj2 = take(B * 8), o_it2 = take(B * 4), o_T2 = take(B * 8),
               o_muf = take(B * S1 * 4 * no * 8), o_obj = take(B * 8),
               o_it = take(B * 4);
  if (cudaMalloc(&l->arena, off) != cudaSuccess) { cudaGetLastError(); obca_b200_loop_destroy(l); return OBCA_E_NOMEM; }
  char* base = (char*)l->arena;
  LoopDev& d = l->d;
  d.B = B; d.N = N; d.S1 = (int)S1; d.M = M; d.Rs = Rs; d.R = (int)R; d.max_steps = lp->max_steps; d.terminal_rule = lp->terminal_rule;
  d.sense = lp->sense; d.goal[0] = lp->goal[0]; d.goal[1] = lp->goal[1]; d.goal_tol = lp->goal_tol;
  d.ego0 = p_free->ego[0]; d.uU0 = p_free->uU[0];
  d.x0 = (double*)(base + o_x0); d.u0 = (double*)(base + o_u0); d.Ts = (double*)(base + o_Ts); d.Ts_opt = (double*)(base + o_Tso);
  d.dyn = (double*)(base + o_dyn); d.dcs = (double*)(base + o_dcs); d.xprev = (double*)(base + o_xprev); d.traj = (double*)(base + o_traj);
  d.alive = (int32_t*)(base + o_alive); d.failed = (int32_t*)(base + o_failed); d.steps = (int32_t*)(base + o_steps);
  d.mode_log = (int32_t*)(base + o_mlog); d.cur = (int32_t*)(base + o_cur);
  d.xref = (double*)(base + o_xref); d.Tmax = (double*)(base + o_Tmax); d.term = (double*)(base + o_term); d.A = (double*)(base + o_A);
  d.b0 = (double*)(base + o_b0); d.db = (double*)(base + o_db); d.x = (double*)(base + o_x); d.u = (double*)(base + o_u);
  d.T = (double*)(base + o_T); d.status = (int32_t*)(base + o_status);
  d.path = (double*)(base + o_path); d.A_s = (double*)(base + o_As); d.b_s = (double*)(base + o_bs);
  d.idx_free = (int32_t*)(base + o_i0); d.idx_set = (int32_t*)(base + o_i1); d.idx_fall = (int32_t*)(base + o_i2);
  d.counts = (int32_t*)(base + o_cnt); d.totals = (unsigned long long*)(base + o_tot);
  l->lam = (double*)(base + o_lam); l->mu = (double*)(base + o_mu);
  l->lam_free = (double*)(base + o_lamf); l->mu_free = (double*)(base + o_muf);
  l->lam2 = (double*)(base + o_lam2); l->mu2 = (double*)(base + o_mu2); l->obj2 = (double*)(base + o_obj2); l->iters2 = (int32_t*)(base + o_it2); l->T2 = (double*)(base + o_T2);
  d.x2 = (double*)(base + o_x2); d.u2 = (double*)(base + o_u2); d.status2 = (int32_t*)(base + o_st2); d.speculative = lp->speculative != 0; l->obj = (double*)(base + o_obj); l->iters = (int32_t*)(base + o_it);
  bool ok = cudaMemcpy(base + o_path, path, (size_t)M * 3 * 8, cudaMemcpyHostToDevice) == cudaSuccess;
  if (Rs > 0) {
    ok = ok && cudaMemcpy(base + o_As, A_static, (size_t)Rs * 2 * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(base + o_bs, b_static, (size_t)Rs * 8, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (!ok) { cudaGetLastError(); obca_b200_loop_destroy(l); return OBCA_E_CUDA; }
  if (cudaStreamCreateWithFlags(&l->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&l->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&l->ev_free, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&l->side2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&l->ev_spec, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError(); obca_b200_loop_destroy(l); return OBCA_E_CUDA;
  }
  l->k = -1;
  *out = l;
  return OBCA_OK;
}

// dyn [B,7] = cx, cy, heading, length, width, speed, first step; heading_cs [B,2] = cos, sin of the heading (HOST)
int obca_b200_loop_reset(obca_loop* l, const double* dyn, const double* heading_cs, void* cuda_stream) {
  if (!l || !dyn || !heading_cs) return OBCA_E_ARG;
  if (cudaSetDevice(l->device) != cudaSuccess) return OBCA_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const LoopDev& d = l->d;
  if (cudaMemcpyAsync(d.dyn, dyn, (size_t)d.B * 7 * 8, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(d.dcs, heading_cs, (size_t)d.B * 2 * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) return OBCA_E_CUDA;
  loop_reset<<<(d.B + 127) / 128, 128, 0, st>>>(d, l->lp.start[0], l->lp.start[1], l->lp.start[2], l->lp.Ts0);
  l->k = 0;
  return cudaGetLastError() == cudaSuccess ? OBCA_OK : OBCA_E_CUDA;
}

// issue n_steps receding-horizon steps on the stream (asynchronous, no host synchronisation)
int obca_b200_loop_run(obca_loop* l, int n_steps, void* cuda_stream) {
  if (!l || n_steps < 0 || l->k < 0 || l->k + n_steps > l->lp.max_steps) return OBCA_E_ARG;
  if (cudaSetDevice(l->device) != cudaSuccess) return OBCA_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const LoopDev& d = l->d;
  const int blocks = (d.B + 127) / 128;
  for (int s = 0; s < n_steps; ++s, ++l->k) {
    if (cudaMemsetAsync(d.counts, 0, 16, st) != cudaSuccess) return OBCA_E_CUDA;
    loop_prepare<<<blocks, 128, 0, st>>>(d, l->k);
    // the free-time solve (its own list, its own instances) runs beside the fixed-time chain: two short launches
    // share the GPU instead of each ending in its own tail.  The two sides write disjoint instances of x, u, T, status;
    // their dual arrays have different row counts per instance, so the free side has its own.
    if (cudaEventRecord(l->ev_ready, st) != cudaSuccess || cudaStreamWaitEvent(l->side, l->ev_ready, 0) != cudaSuccess) return OBCA_E_CUDA;
    int rc = obca_b200_solve_indexed(l->ctx[0], d.B, d.counts + 0, d.idx_free, d.x0, d.u0, d.xref, nullptr, d.Tmax, nullptr, d.Ts,
                                     l->eptr_free, d.A_s, d.b_s, nullptr, 1, d.x, d.u, l->lam_free, l->mu_free, d.T, l->obj, d.status,
                                     l->iters, l->side);
    if (rc != OBCA_OK) return rc;
    if (cudaEventRecord(l->ev_free, l->side) != cudaSuccess) return OBCA_E_CUDA;
    if (d.speculative) {   // solve without the terminal set on every detected scenario, beside the solve with it
      if (cudaStreamWaitEvent(l->side2, l->ev_ready, 0) != cudaSuccess) return OBCA_E_CUDA;
      rc = obca_b200_solve_indexed(l->ctx[2], d.B, d.counts + 2, d.idx_fall, d.x0, d.u0, d.xref, nullptr, nullptr, nullptr, d.Ts,
                                   l->eptr_fix, d.A, d.b0, d.db, 0, d.x2, d.u2, l->lam2, l->mu2, l->T2, l->obj2, d.status2, l->iters2,
                                   l->side2);
      if (rc != OBCA_OK) return rc;
      if (cudaEventRecord(l->ev_spec, l->side2) != cudaSuccess) return OBCA_E_CUDA;
    }
    rc = obca_b200_solve_indexed(l->ctx[1], d.B, d.counts + 1, d.idx_set, d.x0, d.u0, d.xref, nullptr, nullptr, d.term, d.Ts,
                                 l->eptr_fix, d.A, d.b0, d.db, 0, d.x, d.u, l->lam, l->mu, d.T, l->obj, d.status, l->iters, st);
    if (rc != OBCA_OK) return rc;
    if (d.speculative && cudaStreamWaitEvent(st, l->ev_spec, 0) != cudaSuccess) return OBCA_E_CUDA;
    loop_fallback<<<blocks, 128, 0, st>>>(d);
    if (!d.speculative) {
      rc = obca_b200_solve_indexed(l->ctx[2], d.B, d.counts + 2, d.idx_fall, d.x0, d.u0, d.xref, nullptr, nullptr, nullptr, d.Ts,
                                   l->eptr_fix, d.A, d.b0, d.db, 0, d.x, d.u, l->lam, l->mu, d.T, l->obj, d.status, l->iters, st);
      if (rc != OBCA_OK) return rc;
    }
    if (cudaStreamWaitEvent(st, l->ev_free, 0) != cudaSuccess) return OBCA_E_CUDA;
    loop_advance<<<blocks, 128, 0, st>>>(d, l->k);
    if (cudaGetLastError() != cudaSuccess) return OBCA_E_CUDA;
  }
  return OBCA_OK;
}

// copy the logs to HOST buffers (any may be NULL) and synchronise the stream:
// traj [B,max_steps+1,3] (NaN after the last step taken), steps [B], failed [B], mode_log [B,max_steps] (-1 = no solve),
// x [B,3], u [B,2], Ts_opt [B], solves [3] (FREE, FIXED_SET, FIXED_NOTERM totals since reset)
int obca_b200_loop_read(obca_loop* l, double* traj, int32_t* steps, int32_t* failed, int32_t* mode_log, double* x, double* u,
                        double* Ts_opt, int64_t* solves, void* cuda_stream) {
  if (!l) return OBCA_E_ARG;
  if (cudaSetDevice(l->device) != cudaSuccess) return OBCA_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const LoopDev& d = l->d;
  const size_t B = d.B;
  bool ok = true;
  auto get = [&](void* dst, const void* src, size_t bytes) {
    if (dst) ok = ok && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  };
  get(traj, d.traj, B * (d.max_steps + 1) * 3 * 8); get(steps, d.steps, B * 4); get(failed, d.failed, B * 4);
  get(mode_log, d.mode_log, B * d.max_steps * 4); get(x, d.x0, B * 3 * 8); get(u, d.u0, B * 2 * 8); get(Ts_opt, d.Ts_opt, B * 8);
  get(solves, d.totals, 3 * 8);
  ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
  if (!ok) { cudaGetLastError(); return OBCA_E_CUDA; }
  return OBCA_OK;
}

int64_t obca_b200_loop_launch_count(const obca_loop* l) {
  if (!l) return 0;
  return obca_b200_launch_count(l->ctx[0]) + obca_b200_launch_count(l->ctx[1]) + obca_b200_launch_count(l->ctx[2]);
}

}  // extern "C"

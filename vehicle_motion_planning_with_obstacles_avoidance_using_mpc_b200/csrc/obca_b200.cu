// obca_b200.cu - kernel entry + C-ABI (include/obca_b200.h) of the batched OBCA-MPC solver for sm_100a.
//
// Replaces the CasADi/IPOPT work behind obca.obca_mpc4 / obca_mpc6 / obca_mpc8 / obca2 of the reference
// (src/obca.py:828-1071, 1361-1562, 1564-1758, 338-629; called at src/closed_loop.py:118,131,137,170,...).
// No host fallback: every entry point fails with OBCA_E_NODEVICE / OBCA_E_CUDA when there is no GPU.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "obca_phases.cuh"

namespace obca {

__device__ void init_slacks(const Warp& w, Glob& G) {
  const obca_params& P = w.kp.P;
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  const double bp = P.bound_push;
  double z[3], u[2], up[2], zn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) z[j] = w.W(L.Z, j);
#pragma unroll
  for (int j = 0; j < 2; ++j) u[j] = w.W(L.U, j);
#pragma unroll
  for (int j = 0; j < 3; ++j) zn[j] = sh_dn(z[j]);
#pragma unroll
  for (int j = 0; j < 2; ++j) { up[j] = sh_up(u[j]); if (k == 0) up[j] = G.u0[j]; }
  const double T = w.free_ ? G.T : 1.0;
  StageVals sv;
  stage_vals(w, G, z, u, up, zn, T, sv);
#pragma unroll
  for (int j = 0; j < 4; ++j) { w.W(L.SXY, j) = fmax(sv.dxy[j], bp); w.W(L.ZXY, j) = 1.0; }
#pragma unroll
  for (int j = 0; j < 8; ++j) { w.W(L.SUB, j) = (k < N) ? fmax(sv.dub[j], bp) : 1.0; w.W(L.ZUB, j) = 1.0; }
  if (w.free_) {
    G.STb[0] = fmax(T - P.T_min, bp); G.STb[1] = fmax(G.Tmax - T, bp);
    G.ZTb[0] = G.ZTb[1] = 1.0;
  }
  if (w.has_term) {
    const double zN0 = bcast(z[0], N), zN1 = bcast(z[1], N);
    G.Stm[0] = fmax(zN0 - G.term[0], bp); G.Stm[1] = fmax(zN1 - G.term[1], bp); G.Stm[2] = fmax(G.term[2] - zN1, bp);
    G.Ztm[0] = G.Ztm[1] = G.Ztm[2] = 1.0;
  }
  double st, ct;
  sincos(z[2], &st, &ct);
  const double tx = z[0] + G.off * ct, ty = z[1] + G.off * st;
  for (int i = 0; i < w.no; ++i) {
    const int r0 = w.kp.eptr[i], E = w.kp.eptr[i + 1] - r0;
    double a1 = 0, a2 = 0, bl = 0;
    for (int r = r0; r < r0 + E; ++r) {
      const double l = w.W(L.LAM, r);
      a1 += w.A[2 * r] * l; a2 += w.A[2 * r + 1] * l; bl += w.bk(r) * l;
      w.W(L.SL, r) = fmax(l, bp); w.W(L.ZL, r) = 1.0;
    }
    double m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m[q] = w.W(L.MU, 4 * i + q);
      w.W(L.SM, 4 * i + q) = fmax(m[q], bp); w.W(L.ZM, 4 * i + q) = 1.0;
    }
    w.W(L.SN, i) = fmax(1.0 - a1 * a1 - a2 * a2, bp); w.W(L.ZN, i) = 1.0;
    w.W(L.SD, i) = fmax(-(G.g[0] * m[0] + G.g[1] * m[1] + G.g[2] * m[2] + G.g[3] * m[3]) + tx * a1 + ty * a2 - bl - P.dmin, bp);
    w.W(L.ZD, i) = 1.0;
  }
}

// One warp solves one instance; returns status, iteration count and the objective of the final point
__device__ int solve_instance(const Warp& w, Glob& G, int& iters_out, double& obj_out) {
  const obca_params& P = w.kp.P;
  const double s_max = 100.0, kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
  const double dw_first = 1e-4, dw_min = 1e-20, dw_max = 1e20, kw_plus_first = 100.0, kw_plus = 8.0, kw_minus = 1.0 / 3.0;
  const double dc_min = 1e-8, lm_cap = 1e4, stall_alpha = 1e-3;
  const int stall_iters = 10;
  const double g_th = 1e-5, g_ph = 1e-8, s_th = 1.1, s_ph = 2.3, eta_ph = 1e-8;
  const double tol = P.tol;
  const int N = w.N;
  const int m_eq = 3 * N + (w.free_ ? 3 : 0) + 2 * w.no * (N + 1);
  const int q_in = 12 * N + (w.free_ ? 2 : 0) + (w.has_term ? 3 : 0) + (w.R + 6 * w.no) * (N + 1);

  start_point(w, G);
  __syncwarp();
  init_slacks(w, G);
  __syncwarp();

  double mu = P.mu_init;
  double f_th = 0.0, f_ph = 0.0;  // lane i holds filter entry i
  int f_n = 0, f_wr = 0;
  bool f_active = false;
  double thmax = 0, thmin = 0;
  int nstall = 0, acc_count = 0, iter = 0, status = OBCA_ST_MAXITER;
  double dw_last = 0.0, E0 = 0.0, fcur = 0.0;

  for (;;) {
    Err E;
    assemble(w, G, E);
    __syncwarp();
    fcur = E.f;
    if (!E.ok) { status = OBCA_ST_REGFAIL; break; }
    const double sd = fmax(s_max, (E.sumy + E.sumz) / (m_eq + q_in)) / s_max, sc = fmax(s_max, E.sumz / q_in) / s_max;
    E0 = fmax(fmax(E.e1 / sd, E.e2), E.szmax / sc);
    if (E0 <= tol) { status = OBCA_ST_OK; break; }
    if (E0 <= P.acceptable_tol) {
      if (++acc_count >= P.acceptable_iter) { status = OBCA_ST_ACCEPTABLE; break; }
    } else
      acc_count = 0;
    if (iter >= P.max_iter) { status = OBCA_ST_MAXITER; break; }
    // barrier update (monotone Fiacco-McCormick)
    bool changed = false;
    for (;;) {
      const double e3 = fmax(E.szmax - mu, mu - E.szmin) / sc;
      const double Emu = fmax(fmax(E.e1 / sd, E.e2), e3);
      if (Emu <= kappa_eps * mu && mu > tol / 10) {
        mu = fmax(tol / 10, fmin(kappa_mu * mu, pow(mu, theta_mu)));
        changed = true;
      } else
        break;
    }
    if (changed && f_active) { f_n = 0; f_wr = 0; }
    const double th = E.th, ph0 = E.f - mu * E.lgS;
    const double tau = fmax(tau_min, 1 - mu);
    const double dc = w.free_ ? fmax(dc_min, E.ctmax / lm_cap) : 0.0;
    // inertia correction: the Riccati pivots are the inertia test
    double dw = 0.0;
    bool regfail = false;
    for (;;) {
      if (riccati(w, G, mu, dw, dc)) break;
      if (dw == 0.0) dw = (dw_last == 0.0) ? dw_first : fmax(dw_min, kw_minus * dw_last);
      else dw = dw * ((dw_last == 0.0) ? kw_plus_first : kw_plus);
      if (dw > dw_max) { regfail = true; break; }
    }
    if (regfail) { status = OBCA_ST_REGFAIL; break; }
    if (dw > 0) dw_last = dw;
    forward(w, G, dc);
    StepInfo si;
    backsub(w, G, mu, tau, si);
    __syncwarp();
    const double Dphi = si.Dphi;
    if (!f_active) {
      thmax = 1e4 * fmax(1.0, th); thmin = 1e-4 * fmax(1.0, th);
      f_active = true; f_n = 0; f_wr = 0;
    }
    double a_min;
    if (Dphi < 0 && th <= thmin) a_min = fmin(g_th, fmin(g_ph * th / (-Dphi), (th > 0) ? pow(th, s_th) / pow(-Dphi, s_ph) : g_th));
    else if (Dphi < 0) a_min = fmin(g_th, g_ph * th / (-Dphi));
    else a_min = g_th;
    a_min *= 0.05;
    double a = si.a_max;
    int accepted = 0;
    while (a >= a_min * (1 - 1e-12)) {
      double tht, pht;
      trial(w, G, a, mu, tht, pht);
      accepted = 0;
      if (isfinite(pht) && tht < thmax) {
        const bool dom = __any_sync(FULL, (w.lane < f_n) && tht >= f_th && pht >= f_ph);
        if (!dom) {
          const bool sw = (Dphi < 0) && (a * pow(-Dphi, s_ph) > pow(th, s_th));
          if (th <= thmin && sw) {
            if (pht <= ph0 + eta_ph * a * Dphi + 10 * 2.220446049250313e-16 * fabs(ph0)) accepted = 2;
          } else if (tht <= (1 - g_th) * th || pht <= ph0 - g_ph * th)
            accepted = 1;
        }
      }
      if (accepted) break;
      a *= 0.5;
    }
    // IPOPT ends with Solved_To_Acceptable_Level when it cannot progress from an acceptable point; the second
    // clause is the rounding-noise floor of a degenerate vertex of the OBCA dual polytope: final barrier
    // parameter, primal feasible to 1e-6, complementary, only the dual infeasibility above tol
    const bool at_floor = (E0 <= P.acceptable_tol) || (mu <= tol / 10 * (1 + 1e-12) && th <= 1e-6 && E0 <= 1e-3);
    if (!accepted) { status = at_floor ? OBCA_ST_ACCEPTABLE : OBCA_ST_LSFAIL; break; }
    nstall = (a < stall_alpha) ? nstall + 1 : 0;
    if (nstall >= stall_iters) { status = at_floor ? OBCA_ST_ACCEPTABLE : OBCA_ST_STALL; break; }
    if (accepted == 1) {
      const int slot = (f_n < FILT_MAX) ? f_n++ : (f_wr % FILT_MAX);
      if (w.lane == slot) { f_th = (1 - g_th) * th; f_ph = ph0 - g_ph * th; }
      f_wr++;
    }
    update(w, G, a, si.a_z, mu);
    iter++;
  }
  iters_out = iter;
  obj_out = fcur;
  return status;
}

__global__ void __launch_bounds__(128) obca_solve_kernel(const __grid_constant__ KParams kp) {
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int gwarp = blockIdx.x * warps_per_cta + warp_in_cta;
  const obca_params& P = kp.P;
  const int N = P.N, no = P.n_obs, R = P.rows, S = kp.S;
  const Lay L = layout(R, no);
  const bool act = lane <= N;
  const int col = act ? lane : N + 1;
  Warp w{kp, L, kp.ws + (size_t)gwarp * kp.ws_elems * S + col, lane, lane < N ? lane : N, N, S, no, R, act,
         kp.free_ != 0, kp.has_term != 0, kp.stacked != 0, nullptr, nullptr, nullptr, nullptr, nullptr};
  Glob G;
  {
    const double Lc = P.ego[0] + P.ego[2], Wc = P.ego[1] + P.ego[3];
    G.g[0] = Lc / 2; G.g[1] = Wc / 2; G.g[2] = Lc / 2; G.g[3] = Wc / 2;
    G.off = Lc / 2 - P.ego[2];
  }
  for (;;) {
    unsigned int inst = 0;
    if (lane == 0) inst = atomicAdd(kp.counter, 1u);
    inst = __shfl_sync(FULL, inst, 0);
    if (inst >= (unsigned)kp.batch) break;
    const size_t b = inst, ob = kp.shared_obs ? 0 : b;
    w.A = kp.A + ob * 2 * R; w.b0 = kp.b0 + ob * R; w.db = kp.db ? kp.db + ob * R : nullptr;
    w.xref = kp.xref + b * 3 * (N + 1);
    w.uref = kp.uref ? kp.uref + b * 2 * N : nullptr;
#pragma unroll
    for (int j = 0; j < 3; ++j) G.x0[j] = kp.x0[3 * b + j];
#pragma unroll
    for (int j = 0; j < 2; ++j) G.u0[j] = kp.u0[2 * b + j];
    G.Tmax = (w.free_ && kp.Tmax) ? kp.Tmax[b] : 1.0;
    G.T = 1.0; G.dT = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      G.term[j] = (w.has_term && kp.term) ? kp.term[3 * b + j] : 0.0;
      G.Stm[j] = G.Ztm[j] = 1.0; G.dStm[j] = 0.0; G.yt[j] = 0.0; G.dyt[j] = 0.0;
    }
    G.STb[0] = G.STb[1] = G.ZTb[0] = G.ZTb[1] = 1.0; G.dSTb[0] = G.dSTb[1] = 0.0;

    int iters = 0;
    double obj = 0.0;
    const int status = solve_instance(w, G, iters, obj);

    // outputs: x [B,N+1,3]  u [B,N,2]  lam [B,N+1,R]  mu [B,N+1,4 no]  T  obj  status  iters
    if (act) {
      const int k = w.k;
      double* xo = kp.x + (b * (N + 1) + k) * 3;
#pragma unroll
      for (int j = 0; j < 3; ++j) xo[j] = w.W(L.Z, j);
      if (k < N) {
        double* uo = kp.u + (b * N + k) * 2;
        uo[0] = w.W(L.U, 0); uo[1] = w.W(L.U, 1);
      }
      double* lo = kp.lam + (b * (N + 1) + k) * R;
      for (int r = 0; r < R; ++r) lo[r] = w.W(L.LAM, r);
      double* mo = kp.mu + (b * (N + 1) + k) * 4 * no;
      for (int r = 0; r < 4 * no; ++r) mo[r] = w.W(L.MU, r);
    }
    if (lane == 0) {
      kp.T[b] = w.free_ ? G.T : 1.0;
      kp.obj[b] = obj;
      kp.status[b] = status;
      kp.iters[b] = iters;
    }
    __syncwarp();
  }
}

}  // namespace obca

// ======================================================================================================
// C-ABI
// ======================================================================================================
struct obca_ctx {
  int device;
  int max_batch;
  obca_params P;
  obca::Lay L;
  int S, n_warps, grid, block;
  double* ws;
  size_t ws_bytes;
  unsigned int* counter;
  int64_t launches;
  cudaEvent_t ev0, ev1;
  bool timed;
  // host-path staging
  void* stage;
  size_t stage_bytes;
};

static int sm_count_of(int device) {
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

extern "C" {

int obca_b200_abi_version(void) { return OBCA_B200_ABI_VERSION; }

const char* obca_b200_strerror(int rc) {
  switch (rc) {
    case OBCA_OK: return "ok";
    case OBCA_E_ARG: return "invalid argument";
    case OBCA_E_NODEVICE: return "no CUDA device (this library has no host fallback)";
    case OBCA_E_CUDA: return "CUDA runtime error";
    case OBCA_E_NOMEM: return "out of device memory";
    case OBCA_E_SIZE: return "problem size outside compiled limits (N+1 <= 32, obstacles <= 12, rows <= 48)";
    default: return "unknown error";
  }
}

int obca_b200_create(obca_ctx** out, int device, int max_batch, const obca_params* p) {
  if (!out || !p || max_batch < 1) return OBCA_E_ARG;
  *out = nullptr;
  if (p->N < 1 || p->N + 1 > OBCA_MAX_STAGES || p->n_obs < 0 || p->n_obs > OBCA_MAX_OBS || p->rows < 0 || p->rows > OBCA_MAX_ROWS)
    return OBCA_E_SIZE;
  if (p->mode < 0 || p->mode > OBCA_MODE_FIXED_OBCA2) return OBCA_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return OBCA_E_NODEVICE; }
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return OBCA_E_CUDA;
  if (device >= ndev) return OBCA_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return OBCA_E_CUDA;
  obca_ctx* c = (obca_ctx*)calloc(1, sizeof(obca_ctx));
  if (!c) return OBCA_E_NOMEM;
  c->device = device; c->max_batch = max_batch; c->P = *p;
  c->L = obca::layout(p->rows, p->n_obs);
  c->S = p->N + 2;
  c->block = 128;
  const int warps_per_cta = c->block / 32;
  int ctas_per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, obca::obca_solve_kernel, c->block, 0);
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const char* env = getenv("OBCA_CTAS_PER_SM");
  if (env && atoi(env) > 0 && atoi(env) < ctas_per_sm) ctas_per_sm = atoi(env);
  int grid = sm_count_of(device) * ctas_per_sm;
  const int need = (max_batch + warps_per_cta - 1) / warps_per_cta;
  if (grid > need) grid = need;
  c->grid = grid; c->n_warps = grid * warps_per_cta;
  c->ws_bytes = (size_t)c->n_warps * c->L.total * c->S * sizeof(double);
  if (cudaMalloc(&c->ws, c->ws_bytes) != cudaSuccess) { cudaGetLastError(); free(c); return OBCA_E_NOMEM; }
  if (cudaMalloc(&c->counter, sizeof(unsigned int)) != cudaSuccess) { cudaFree(c->ws); free(c); return OBCA_E_NOMEM; }
  cudaMemset(c->ws, 0, c->ws_bytes);
  cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
  *out = c;
  return OBCA_OK;
}

int obca_b200_destroy(obca_ctx* c) {
  if (!c) return OBCA_E_ARG;
  cudaSetDevice(c->device);
  cudaFree(c->ws); cudaFree(c->counter);
  if (c->stage) cudaFree(c->stage);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  free(c);
  return OBCA_OK;
}

int64_t obca_b200_scratch_bytes(const obca_ctx* c) { return c ? (int64_t)c->ws_bytes : 0; }
int64_t obca_b200_launch_count(const obca_ctx* c) { return c ? c->launches : 0; }

float obca_b200_last_kernel_ms(obca_ctx* c) {
  if (!c || !c->timed) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
  return ms;
}

int obca_b200_solve(obca_ctx* c, int batch, const double* x0, const double* u0, const double* xref, const double* uref,
                    const double* T_max, const double* term, const int32_t* edge_ptr, const double* A, const double* b0,
                    const double* db, int obstacles_shared, double* x, double* u, double* lam, double* mu, double* T,
                    double* obj, int32_t* status, int32_t* iters, void* cuda_stream) {
  if (!c || batch < 0 || batch > c->max_batch) return OBCA_E_ARG;
  if (batch == 0) return OBCA_OK;
  if (!x0 || !u0 || !xref || !edge_ptr || !x || !u || !lam || !mu || !T || !obj || !status || !iters) return OBCA_E_ARG;
  const obca_params& P = c->P;
  if (P.n_obs > 0 && (!A || !b0)) return OBCA_E_ARG;
  const bool free_ = (P.mode == OBCA_MODE_FREE || P.mode == OBCA_MODE_FREE_STACKED);
  const bool has_term = (P.mode == OBCA_MODE_FIXED_SET) || (P.mode == OBCA_MODE_FIXED_OBCA2 && P.has_term);
  if (free_ && !T_max) return OBCA_E_ARG;
  if (has_term && !term) return OBCA_E_ARG;
  if (edge_ptr[0] != 0 || edge_ptr[P.n_obs] != P.rows) return OBCA_E_ARG;
  if (cudaSetDevice(c->device) != cudaSuccess) return OBCA_E_CUDA;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  obca::KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.P = P;
  for (int i = 0; i <= P.n_obs; ++i) kp.eptr[i] = edge_ptr[i];
  kp.batch = batch; kp.shared_obs = obstacles_shared ? 1 : 0;
  kp.free_ = free_; kp.has_term = has_term; kp.stacked = (P.mode != OBCA_MODE_FREE);
  kp.S = c->S; kp.ws_elems = c->L.total;
  kp.x0 = x0; kp.u0 = u0; kp.xref = xref; kp.uref = uref; kp.Tmax = T_max; kp.term = term;
  kp.A = A; kp.b0 = b0; kp.db = db;
  kp.x = x; kp.u = u; kp.lam = lam; kp.mu = mu; kp.T = T; kp.obj = obj; kp.status = status; kp.iters = iters;
  kp.ws = c->ws; kp.counter = c->counter;
  if (cudaMemsetAsync(c->counter, 0, sizeof(unsigned int), st) != cudaSuccess) return OBCA_E_CUDA;
  cudaEventRecord(c->ev0, st);
  obca::obca_solve_kernel<<<c->grid, c->block, 0, st>>>(kp);
  cudaEventRecord(c->ev1, st);
  c->timed = true;
  c->launches += 1;
  if (cudaGetLastError() != cudaSuccess) return OBCA_E_CUDA;
  return OBCA_OK;
}

int obca_b200_solve_host(obca_ctx* c, int batch, const double* x0, const double* u0, const double* xref, const double* uref,
                         const double* T_max, const double* term, const int32_t* edge_ptr, const double* A,
                         const double* b0, const double* db, int obstacles_shared, double* x, double* u, double* lam,
                         double* mu, double* T, double* obj, int32_t* status, int32_t* iters) {
  if (!c || batch < 0 || batch > c->max_batch) return OBCA_E_ARG;
  if (batch == 0) return OBCA_OK;
  const obca_params& P = c->P;
  if (cudaSetDevice(c->device) != cudaSuccess) return OBCA_E_CUDA;
  const size_t B = batch, N = P.N, R = P.rows, no = P.n_obs, Bo = obstacles_shared ? 1 : B;
  // one staging buffer: inputs then outputs, all 8-byte aligned
  const size_t n_in[9] = {B * 3, B * 2, B * (N + 1) * 3, uref ? B * N * 2 : 0, T_max ? B : 0, term ? B * 3 : 0,
                          Bo * R * 2, Bo * R, db ? Bo * R : 0};
  const double* h_in[9] = {x0, u0, xref, uref, T_max, term, A, b0, db};
  const size_t n_out[6] = {B * (N + 1) * 3, B * N * 2, B * (N + 1) * R, B * (N + 1) * 4 * no, B, B};
  double* h_out[6] = {x, u, lam, mu, T, obj};
  size_t tot = 0;
  for (int i = 0; i < 9; ++i) tot += n_in[i];
  for (int i = 0; i < 6; ++i) tot += n_out[i];
  const size_t bytes = tot * sizeof(double) + 2 * B * sizeof(int32_t);
  if (bytes > c->stage_bytes) {
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_bytes = 0;
    if (cudaMalloc(&c->stage, bytes) != cudaSuccess) { cudaGetLastError(); return OBCA_E_NOMEM; }
    c->stage_bytes = bytes;
  }
  double* d = (double*)c->stage;
  double* d_in[9];
  double* d_out[6];
  for (int i = 0; i < 9; ++i) {
    d_in[i] = n_in[i] ? d : nullptr;
    if (n_in[i] && !h_in[i]) return OBCA_E_ARG;
    if (n_in[i] && cudaMemcpyAsync(d, h_in[i], n_in[i] * sizeof(double), cudaMemcpyHostToDevice, 0) != cudaSuccess) return OBCA_E_CUDA;
    d += n_in[i];
  }
  for (int i = 0; i < 6; ++i) { d_out[i] = d; d += n_out[i]; }
  int32_t* d_status = (int32_t*)d;
  int32_t* d_iters = d_status + B;
  int rc = obca_b200_solve(c, batch, d_in[0], d_in[1], d_in[2], d_in[3], d_in[4], d_in[5], edge_ptr, d_in[6], d_in[7], d_in[8],
                           obstacles_shared, d_out[0], d_out[1], d_out[2], d_out[3], d_out[4], d_out[5], d_status, d_iters, 0);
  if (rc != OBCA_OK) return rc;
  for (int i = 0; i < 6; ++i) {
    if (!h_out[i]) return OBCA_E_ARG;
    if (cudaMemcpyAsync(h_out[i], d_out[i], n_out[i] * sizeof(double), cudaMemcpyDeviceToHost, 0) != cudaSuccess) return OBCA_E_CUDA;
  }
  if (cudaMemcpyAsync(status, d_status, B * sizeof(int32_t), cudaMemcpyDeviceToHost, 0) != cudaSuccess) return OBCA_E_CUDA;
  if (cudaMemcpyAsync(iters, d_iters, B * sizeof(int32_t), cudaMemcpyDeviceToHost, 0) != cudaSuccess) return OBCA_E_CUDA;
  if (cudaStreamSynchronize(0) != cudaSuccess) return OBCA_E_CUDA;
  return OBCA_OK;
}

}  // extern "C"

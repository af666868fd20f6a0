// obca_emu.cpp - developer/test tool: runs the device phase code of csrc/obca_cta.cuh on the HOST, the threads
// of a block executed one after the other (every phase boundary of the kernel is a loop boundary here).  Lets
// the kernel LOGIC (mapping, shared-memory layout, phase ordering, reductions) be checked against the oracle on
// a box without a GPU; built by tests/test_kernel_emulation.py with g++.  Not part of the product path.
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200/csrc/obca_cta.cuh"

using namespace obca;

template <int EMAX>
struct HostExec {
  std::vector<BlockRegs<EMAX>> brs;
  std::vector<std::array<double, NPART_X>> parts;
  double red[NPART_X];
  int T;
  explicit HostExec(int T_) : brs(T_), parts(T_), T(T_) {
    for (auto& p : parts) p.fill(0.0);
    memset(brs.data(), 0, sizeof(BlockRegs<EMAX>) * T_);
  }
  template <class F> void par(F&& f) { for (int t = 0; t < T; ++t) f(t, brs[t], parts[t].data()); }
  template <class F> void all(F&& f) { for (int t = 0; t < T; ++t) f(t); }
  SweepRegs srs[32];
  template <class F> void sweep(F&& f) { for (int l = 0; l < 32; ++l) f(l, srs[l]); }
  template <class SolverT> void sweep_stages(const SolverT& S, double mu, double dw) {
    const obca::SwMemPtr m = S.sweep_mem();
    for (int s = S.N - 1; s >= 0; --s) {
      for (int l = 0; l < 32; ++l) obca::SweepOps<obca::SwMemPtr>::w(m, s, srs[l]);
      for (int l = 0; l < 32; ++l) obca::SweepOps<obca::SwMemPtr>::f(m, s, mu, dw, srs[l]);
      for (int l = 0; l < 32; ++l) obca::SweepOps<obca::SwMemPtr>::b(m, l, s, srs[l]);
    }
  }
  template <class F> void stage(F&& f) { for (int l = 0; l < 32; ++l) f(l); }
  void stage_end() {}
  void align() {}
  template <class F> void once(F&& f) { f(); }
  void tick(int) {}
  void trace(int it, double f, double th, double E0, double mu, double dw, double a) {
    if (getenv("OBCA_EMU_TRACE")) printf("  it %3d f %.6e th %.3e E0 %.3e mu %.1e dw %.1e a %.3e\n", it, f, th, E0, mu, dw, a);
  }
  template <int S0, int NS, int M0, int NM, int N0, int NN> void reduce(double* scratch) {
    // the device version stages the partials in `scratch`, which aliases step / Hessian storage: clobber it here so
    // that a wrong liveness assumption shows up in the emulation as well
    for (int i = 0; i < (NS + NM + NN) * red_stride(T); ++i) scratch[i] = 0.0 / 0.0;
    for (int q = 0; q < NS; ++q) { double a = 0; for (int t = 0; t < T; ++t) a += parts[t][S0 + q]; red[S0 + q] = a; }
    for (int q = 0; q < NM; ++q) { double a = parts[0][M0 + q]; for (int t = 1; t < T; ++t) a = fmax(a, parts[t][M0 + q]); red[M0 + q] = a; }
    for (int q = 0; q < NN; ++q) { double a = parts[0][N0 + q]; for (int t = 1; t < T; ++t) a = fmin(a, parts[t][N0 + q]); red[N0 + q] = a; }
  }
};

template <int EMAX>
static void run(const KParams& kp, int nwarps, int has_uref) {
  const obca_params& P = kp.P;
  Sm sm;
  size_t nd = sm_carve(sm, nullptr, P.N, P.n_obs, P.rows, nwarps, has_uref);
  std::vector<double> mem(nd + 8, 0.0);
  sm_carve(sm, mem.data(), P.N, P.n_obs, P.rows, nwarps, has_uref);
  Solver<EMAX> S(kp, sm);
  for (int b = 0; b < kp.batch; ++b) {
    std::fill(mem.begin(), mem.end(), 0.0);
    HostExec<EMAX> ex(32 * nwarps);
    for (int t = 0; t < ex.T; ++t) S.load(t, (size_t)b, true);
    int iters = 0;
    double obj = 0;
    const int wdn = Solver<EMAX>::wd_doubles(ex.T, P.N + 1);
    std::vector<double> wd(2 * (size_t)wdn, 0.0);   // watchdog reference | point of failure
    const int st = solve_with_recovery(S, ex, (size_t)b, wd.data(), wd.data() + wdn, iters, obj);
    if (st != OBCA_ST_STORED)
      for (int t = 0; t < ex.T; ++t) S.store(t, ex.brs[t], (size_t)b, st, iters, obj);
    else { kp.obj[b] = obj; kp.iters[b] = iters; }
  }
}

extern "C" int obca_emu_solve(const obca_params* P, int batch, const double* x0, const double* u0, const double* xref,
                              const double* uref, const double* T_max, const double* term, const double* Ts_inst,
                              const int32_t* edge_ptr, const double* A, const double* b0, const double* db,
                              int obstacles_shared, double* x, double* u, double* lam, double* mu, double* T, double* obj,
                              int32_t* status, int32_t* iters, int nthreads) {
  (void)nthreads;
  KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.P = *P;
  int emax = 0;
  for (int i = 0; i <= P->n_obs; ++i) kp.eptr[i] = edge_ptr[i];
  for (int i = 0; i < P->n_obs; ++i) emax = emax > edge_ptr[i + 1] - edge_ptr[i] ? emax : edge_ptr[i + 1] - edge_ptr[i];
  kp.batch = batch; kp.shared_obs = obstacles_shared ? 1 : 0;
  kp.free_ = (P->mode == OBCA_MODE_FREE || P->mode == OBCA_MODE_FREE_STACKED);
  kp.has_term = (P->mode == OBCA_MODE_FIXED_SET) || (P->mode == OBCA_MODE_FIXED_OBCA2 && P->has_term);
  kp.stacked = (P->mode != OBCA_MODE_FREE);
  kp.x0 = x0; kp.u0 = u0; kp.xref = xref; kp.uref = uref; kp.Tmax = T_max; kp.term = term; kp.Ts_inst = Ts_inst;
  kp.A = A; kp.b0 = b0; kp.db = db;
  kp.x = x; kp.u = u; kp.lam = lam; kp.mu = mu; kp.T = T; kp.obj = obj; kp.status = status; kp.iters = iters;
  const int nb = P->n_obs * (P->N + 1);
  const int nwarps = (nb + 31) / 32 + 1;
  if (emax <= 4) run<4>(kp, nwarps, uref != nullptr);
  else if (emax <= 8) run<8>(kp, nwarps, uref != nullptr);
  else return OBCA_E_SIZE;
  return OBCA_OK;
}

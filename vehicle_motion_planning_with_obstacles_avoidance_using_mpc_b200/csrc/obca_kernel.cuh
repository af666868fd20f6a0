// obca_kernel.cuh - batched OBCA-MPC interior-point solver, one warp per NLP instance (sm_100a).
//
// The NLP is the reference's (src/obca.py: obca_mpc4 828-1071, obca_mpc6 1361-1562, obca_mpc8 1564-1758,
// obca2 338-629) in the compact variable set of SURVEY.md Appendix A.  The algorithm is the primal-dual
// interior-point method specified by oracle/ipm_dense.py; nothing here is shared with oracle/ (the oracle is
// scalar C and only the tests call it).
//
// Mapping.  Lane k of the warp owns stage k (N + 1 <= 32): pose z_k, input u_k, the OBCA duals (lambda, mu)
// of every obstacle at step k, their slacks and multipliers.  Per iteration:
//   1. assemble   (lane-parallel)  residuals, Lagrangian gradient, stage Hessian; every (stage, obstacle) dual
//                                  block is eliminated through a 5x5 square-root (Givens) factorisation held
//                                  in registers and condensed onto the 3x3 pose block
//   2. riccati    (lane-sequential, state handed from lane k+1 to lane k) over the augmented stage state
//                                  (x, y, theta, v_prev, w_prev, T); pivots double as the inertia test
//   3. forward    (lane-sequential) roll-out of the step
//   4. backsub    (lane-parallel)  dual-block steps, slack steps, fraction-to-boundary (warp min-reduce)
//   5. line search (lane-parallel) filter test on warp-reduced (theta, phi)
// All arithmetic is fp64.  The iterate lives in a per-warp workspace laid out [element][stage] so that the 32
// lanes of a warp touch consecutive doubles (coalesced, L2-resident: the workspace is per resident warp, not
// per instance); inputs are read once and outputs written once per instance.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/obca_b200.h"

namespace obca {

#define FULL 0xffffffffu
constexpr int FILT_MAX = 32;
constexpr double SIG_MIN = 1e-8;  // primal regularisation of the OBCA duals (curvature floor of a sign row)

struct KParams {
  obca_params P;
  int32_t eptr[OBCA_MAX_OBS + 1];
  int32_t batch, shared_obs, free_, has_term, stacked, S;  // S = workspace stride (N + 2)
  int32_t ws_elems;                                         // elements (x S doubles) per warp workspace
  const double *x0, *u0, *xref, *uref, *Tmax, *term, *A, *b0, *db;
  double *x, *u, *lam, *mu, *T, *obj;
  int32_t *status, *iters;
  double* ws;
  unsigned int* counter;  // persistent-warp work queue
};

// ---- workspace element offsets (units of S doubles), filled by layout()
struct Lay {
  int Z, U, YD, SXY, ZXY, SUB, ZUB, LAM, SL, ZL, MU, SM, ZM, YE, SN, ZN, SD, ZD;
  int DZ, DU, DYD, DSXY, DSUB, DLAM, DMU, DYE, DSN, DSD;
  int H, RA, RB, CD, K, KAP, PM, PV, ETA, total;
};
__host__ __device__ inline Lay layout(int R, int no) {
  Lay L; int o = 0;
  L.Z = o; o += 3; L.U = o; o += 2; L.YD = o; o += 3;
  L.SXY = o; o += 4; L.ZXY = o; o += 4; L.SUB = o; o += 8; L.ZUB = o; o += 8;
  L.LAM = o; o += R; L.SL = o; o += R; L.ZL = o; o += R;
  L.MU = o; o += 4 * no; L.SM = o; o += 4 * no; L.ZM = o; o += 4 * no;
  L.YE = o; o += 2 * no; L.SN = o; o += no; L.ZN = o; o += no; L.SD = o; o += no; L.ZD = o; o += no;
  L.DZ = o; o += 3; L.DU = o; o += 2; L.DYD = o; o += 3; L.DSXY = o; o += 4; L.DSUB = o; o += 8;
  L.DLAM = o; o += R; L.DMU = o; o += 4 * no; L.DYE = o; o += 2 * no; L.DSN = o; o += no; L.DSD = o; o += no;
  L.H = o; o += 36; L.RA = o; o += 8; L.RB = o; o += 8; L.CD = o; o += 3;
  L.K = o; o += 12; L.KAP = o; o += 2; L.PM = o; o += 21; L.PV = o; o += 6;
  L.ETA = o; o += 25 * no;  // per block: eta0a(5) eta0b(5) eta_x(5) eta_y(5) eta_th(5)
  L.total = o;
  return L;
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double sh_up(double v) { return __shfl_up_sync(FULL, v, 1); }
__device__ __forceinline__ double sh_dn(double v) { return __shfl_down_sync(FULL, v, 1); }
__device__ __forceinline__ int symi(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// warp-uniform scalar state of one instance
struct Glob {
  double T, STb[2], ZTb[2], Stm[3], Ztm[3], yt[3];
  double dT, dSTb[2], dStm[3], dyt[3];
  double Tmax, x0[3], u0[2], term[3];
  double off, g[4];
};

struct Warp {
  const KParams& kp;
  const Lay& L;
  double* ws;   // this warp's workspace, already offset by the lane's column
  int lane, k, N, S, no, R;
  bool act, free_, has_term, stacked;
  const double *A, *b0, *db, *xref, *uref;
  __device__ __forceinline__ double& W(int off, int j) const { return ws[(size_t)(off + j) * S]; }
  __device__ __forceinline__ double bk(int r) const { return b0[r] + ((stacked && db) ? k * db[r] : 0.0); }
};

// 5x5 square-root factor R^T (lower triangular, packed) with Givens row insertion
struct Tri5 {
  double l[15];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 15; ++i) l[i] = 0.0;
  }
  __device__ __forceinline__ double& at(int r, int c) { return l[r * (r + 1) / 2 + c]; }
  __device__ __forceinline__ void insert(double row[5]) {
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      double a = at(c, c), b = row[c];
      if (b != 0.0) {
        double rr = sqrt(a * a + b * b), cs = a / rr, sn = b / rr;
        at(c, c) = rr;
#pragma unroll
        for (int q = c + 1; q < 5; ++q) {
          double u = at(q, c), w = row[q];
          at(q, c) = cs * u + sn * w;
          row[q] = -sn * u + cs * w;
        }
      }
    }
  }
  __device__ __forceinline__ bool finish() {
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      if (at(a, a) < 0) {
#pragma unroll
        for (int q = a; q < 5; ++q) at(q, a) = -at(q, a);
      }
      ok = ok && (at(a, a) > 0) && isfinite(at(a, a));
    }
    return ok;
  }
  __device__ __forceinline__ void solve(const double r[5], double x[5]) {  // (L L^T) x = r
    double t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double s = r[i];
#pragma unroll
      for (int q = 0; q < i; ++q) s -= at(i, q) * t[q];
      t[i] = s / at(i, i);
    }
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      double s = t[i];
#pragma unroll
      for (int q = i + 1; q < 5; ++q) s -= at(q, i) * x[q];
      x[i] = s / at(i, i);
    }
  }
};

struct BlkGeo {  // geometry of one (stage, obstacle) block at the current point
  double a1, a2, ct, st, tx, ty;
};

__device__ __forceinline__ void row_y(const Warp& w, const Glob& G, const BlkGeo& b, int r0, int E, int j, double yv[5]) {
  if (j < E) {
    int r = r0 + j;
    double A0 = w.A[2 * r], A1 = w.A[2 * r + 1];
    yv[0] = A0; yv[1] = A1;
    yv[2] = b.tx * A0 + b.ty * A1 - w.bk(r);
    yv[3] = b.ct * A0 + b.st * A1;
    yv[4] = -b.st * A0 + b.ct * A1;
  } else {
    int m = j - E;
    yv[0] = 0; yv[1] = 0; yv[2] = -G.g[m];
    yv[3] = (m == 0) ? 1.0 : (m == 2) ? -1.0 : 0.0;
    yv[4] = (m == 1) ? 1.0 : (m == 3) ? -1.0 : 0.0;
  }
}

// Cn^-1 v = (v - kn a (a.v)) / (2 Zn)
__device__ __forceinline__ void cn_inv(const BlkGeo& b, double ci0, double ci1, const double v[2], double o[2]) {
  double av = b.a1 * v[0] + b.a2 * v[1];
  o[0] = (v[0] - ci1 * b.a1 * av) * ci0;
  o[1] = (v[1] - ci1 * b.a2 * av) * ci0;
}

struct Err {  // warp-reduced quantities of one assemble pass
  double f, th, lgS, e1, e2, sumy, sumz, szmax, szmin, ctmax;
  bool ok;
};

}  // namespace obca

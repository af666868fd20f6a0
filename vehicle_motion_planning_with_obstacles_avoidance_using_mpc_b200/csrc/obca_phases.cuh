// obca_phases.cuh - the per-iteration phases of the warp-per-instance interior-point solver.
// See obca_kernel.cuh for the mapping; formulas follow SURVEY.md Appendix A (reference lines cited there).
#pragma once
#include "obca_kernel.cuh"

namespace obca {

// statistics of one inequality (slack S, multiplier Z, value d); returns sigma, and the mu-split of the
// step-form right-hand side  t - Z = mu * ta + tb
struct IneqAcc {
  double th = 0, lg = 0, cmax = 0, sumz = 0, szmax = 0, szmin = 1e300;
  __device__ __forceinline__ void add(double S, double Z, double d, double& sig, double& ta, double& tb) {
    double rd = d - S;
    sig = Z / S; ta = 1.0 / S; tb = -Z - sig * rd;
    th += fabs(rd); cmax = fmax(cmax, fabs(rd)); lg += log(S); sumz += Z;
    double sz = S * Z; szmax = fmax(szmax, sz); szmin = fmin(szmin, sz);
  }
};

// ------------------------------------------------------------------------------------------------------
// start point (oracle/obca_nlp.py start_point): init 0 = reference (all zero, T = 1: obca.py:856),
// 1 = poses from xref, 2 = A* warm start
// ------------------------------------------------------------------------------------------------------
__device__ void start_point(const Warp& w, Glob& G) {
  const obca_params& P = w.kp.P;
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  double z[3], pp[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    pp[j] = (k == 0) ? G.x0[j] : (w.act ? w.xref[3 * k + j] : 0.0);
    z[j] = (k == 0) ? G.x0[j] : (P.init >= OBCA_INIT_XREF ? pp[j] : 0.0);
    w.W(L.Z, j) = z[j];
    w.W(L.YD, j) = 0.0;
  }
  double pn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) pn[j] = sh_dn(pp[j]);
  double u[2] = {0, 0};
  G.T = 1.0;
  if (P.init == OBCA_INIT_WARM) {
    double seg = (w.act && k < N) ? sqrt((pn[0] - pp[0]) * (pn[0] - pp[0]) + (pn[1] - pp[1]) * (pn[1] - pp[1])) : 0.0;
    double len = wsum(seg);
    double h = P.Ts;
    if (w.free_) {
      double T0 = len / (N * P.uU[0] * P.Ts);
      T0 = fmin(fmax(T0, 1.0), fmax(G.Tmax, P.T_min));
      G.T = T0;
      h = T0 * P.Ts;
    }
    if (k < N) {
      double dth = pn[2] - pp[2] + M_PI;
      dth = dth - 2 * M_PI * floor(dth / (2 * M_PI)) - M_PI;
      double fwd = cos(pp[2]) * (pn[0] - pp[0]) + sin(pp[2]) * (pn[1] - pp[1]);
      u[0] = fmin(fmax(fwd / h, P.uL[0]), P.uU[0]);
      u[1] = fmin(fmax(dth / h, P.uL[1]), P.uU[1]);
    }
  }
  w.W(L.U, 0) = u[0]; w.W(L.U, 1) = u[1];
  for (int r = 0; r < w.R; ++r) w.W(L.LAM, r) = 0.0;
  for (int r = 0; r < 4 * w.no; ++r) w.W(L.MU, r) = 0.0;
  for (int r = 0; r < 2 * w.no; ++r) w.W(L.YE, r) = 0.0;
  if (P.init == OBCA_INIT_WARM) {
    double ct = cos(pp[2]), st = sin(pp[2]);
    double tx = pp[0] + G.off * ct, ty = pp[1] + G.off * st;
    for (int i = 0; i < w.no; ++i) {
      int jb = -1;
      double best = -1e300, nb = 1;
      for (int r = w.kp.eptr[i]; r < w.kp.eptr[i + 1]; ++r) {
        double A0 = w.A[2 * r], A1 = w.A[2 * r + 1];
        double nr = sqrt(A0 * A0 + A1 * A1);
        double sep = (A0 * tx + A1 * ty - w.bk(r)) / nr;
        if (sep > best) { best = sep; jb = r; nb = nr; }
      }
      if (jb < 0) continue;
      double l = 0.9 / nb;
      w.W(L.LAM, jb) = l;
      double a1 = w.A[2 * jb] * l, a2 = w.A[2 * jb + 1] * l;
      double r1 = -(ct * a1 + st * a2), r2 = -(-st * a1 + ct * a2);
      w.W(L.MU, 4 * i + 0) = fmax(r1, 0.0); w.W(L.MU, 4 * i + 1) = fmax(r2, 0.0);
      w.W(L.MU, 4 * i + 2) = fmax(-r1, 0.0); w.W(L.MU, 4 * i + 3) = fmax(-r2, 0.0);
    }
  }
  G.yt[0] = G.yt[1] = G.yt[2] = 0.0;
}

// values of the stage's own (non-obstacle) constraints at a point
struct StageVals {
  double f, cd[3], dxy[4], dub[8];
};
__device__ __forceinline__ void stage_vals(const Warp& w, const Glob& G, const double z[3], const double u[2],
                                           const double up[2], const double zn[3], double T, StageVals& o) {
  const obca_params& P = w.kp.P;
  const int k = w.k, N = w.N;
  double h = T * P.Ts;
  const double* M = (k < N) ? P.Q : P.P;
  double e[3], f = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) e[j] = z[j] - w.xref[3 * k + j];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) f += e[a] * M[3 * a + b] * e[b];
  if (k < N) {
    double uu[2] = {u[0], u[1]};
    if (w.uref) { uu[0] -= w.uref[2 * k]; uu[1] -= w.uref[2 * k + 1]; }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) f += uu[a] * P.R1[2 * a + b] * uu[b];
    if (k >= 1) {
      double du[2] = {u[0] - up[0], u[1] - up[1]}, s = 0;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) s += du[a] * P.R2[2 * a + b] * du[b];
      f += s / (h * h);
    }
    double st, ct;
    sincos(z[2], &st, &ct);
    o.cd[0] = z[0] + h * u[0] * ct - zn[0];
    o.cd[1] = z[1] + h * u[0] * st - zn[1];
    o.cd[2] = z[2] + h * u[1] - zn[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double ga = (up[j] - u[j]) / h;
      o.dub[j] = u[j] - P.uL[j];
      o.dub[2 + j] = P.uU[j] - u[j];
      o.dub[4 + j] = ga + P.acc_max[j];
      o.dub[6 + j] = P.acc_max[j] - ga;
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    o.dxy[j] = z[j] - P.xL[j];
    o.dxy[2 + j] = P.xU[j] - z[j];
  }
  if (k == 0 && w.free_) f += (N + 1) * (P.time_cost[0] * T + P.time_cost[1] * T * T);
  o.f = f;
  (void)G;
}

// ------------------------------------------------------------------------------------------------------
// assemble: everything that does not depend on the barrier parameter is computed once; the right-hand
// sides are kept split as  mu * (a part) + (b part)  so that mu can be chosen afterwards from this pass's
// own optimality error (IPOPT's barrier update) without a second pass.
// ------------------------------------------------------------------------------------------------------
__device__ void assemble(const Warp& w, const Glob& G, Err& E) {
  const obca_params& P = w.kp.P;
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  const bool act = w.act, free_ = w.free_;
  double z[3], u[2], yd[3], up[2], zn[3], ydm[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { z[j] = w.W(L.Z, j); yd[j] = w.W(L.YD, j); }
  u[0] = w.W(L.U, 0); u[1] = w.W(L.U, 1);
#pragma unroll
  for (int j = 0; j < 3; ++j) { zn[j] = sh_dn(z[j]); ydm[j] = sh_up(yd[j]); }
#pragma unroll
  for (int j = 0; j < 2; ++j) { up[j] = sh_up(u[j]); if (k == 0) up[j] = G.u0[j]; }
  const double T = free_ ? G.T : 1.0, h = T * P.Ts;
  double st, ct;
  sincos(z[2], &st, &ct);

  double H[36], ra[8], rb[8], gL[8], gf[8];
#pragma unroll
  for (int i = 0; i < 36; ++i) H[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { ra[i] = rb[i] = gL[i] = gf[i] = 0; }
#define HH(a, b) H[((a) >= (b)) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a))]

  StageVals sv;
  stage_vals(w, G, z, u, up, zn, T, sv);
  IneqAcc acc;
  double sumy = 0, e1 = 0, ceq_th = 0, ceq_max = 0;

  // (1) tracking cost
  {
    const double* M = (k < N) ? P.Q : P.P;
    double e[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = z[j] - w.xref[3 * k + j];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double s = 0;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double m = M[3 * a + b] + M[3 * b + a];
        s += m * e[b];
        if (b <= a) HH(a, b) += m;
      }
      gf[a] += s;
    }
  }
  if (k < N) {
    // (2) input cost
    double uu[2] = {u[0], u[1]};
    if (w.uref) { uu[0] -= w.uref[2 * k]; uu[1] -= w.uref[2 * k + 1]; }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      double s = 0;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        double m = P.R1[2 * a + b] + P.R1[2 * b + a];
        s += m * uu[b];
        if (b <= a) HH(6 + a, 6 + b) += m;
      }
      gf[6 + a] += s;
    }
    // (3) acceleration cost between u_{k-1} (state 3,4) and u_k, k >= 1 (the t == 0 term is identically 0)
    if (k >= 1) {
      double du[2] = {u[0] - up[0], u[1] - up[1]}, qv[2], Aacc = 0;
      const double ih2 = 1.0 / (h * h);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        qv[a] = 0;
#pragma unroll
        for (int b = 0; b < 2; ++b) qv[a] += 0.5 * (P.R2[2 * a + b] + P.R2[2 * b + a]) * du[b];
        Aacc += du[a] * qv[a];
      }
      Aacc *= ih2;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        gf[6 + a] += 2 * qv[a] * ih2;
        gf[3 + a] -= 2 * qv[a] * ih2;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          double m = (P.R2[2 * a + b] + P.R2[2 * b + a]) * ih2;
          if (b <= a) { HH(6 + a, 6 + b) += m; HH(3 + a, 3 + b) += m; }
          HH(6 + a, 3 + b) -= m;
        }
        if (free_) {
          double c = 4 * qv[a] * ih2 / T;
          HH(6 + a, 5) -= c; HH(5, 3 + a) += c;
        }
      }
      if (free_) { gf[5] -= 2 * Aacc / T; HH(5, 5) += 6 * Aacc / (T * T); }
    }
    // (5) dynamics: Hessian-of-Lagrangian terms and J^T y
    const double vv = u[0], ww = u[1];
    const double fth0 = -h * vv * st, fth1 = h * vv * ct;
    const double fT0 = free_ ? P.Ts * vv * ct : 0.0, fT1 = free_ ? P.Ts * vv * st : 0.0, fT2 = free_ ? P.Ts * ww : 0.0;
    HH(2, 2) += h * vv * (-yd[0] * ct - yd[1] * st);
    HH(6, 2) += h * (-yd[0] * st + yd[1] * ct);
    if (free_) {
      HH(5, 2) += P.Ts * vv * (-yd[0] * st + yd[1] * ct);
      HH(6, 5) += P.Ts * (yd[0] * ct + yd[1] * st);
      HH(7, 5) += P.Ts * yd[2];
    }
    gL[0] += yd[0]; gL[1] += yd[1]; gL[2] += yd[2] + fth0 * yd[0] + fth1 * yd[1];
    gL[6] += h * ct * yd[0] + h * st * yd[1]; gL[7] += h * yd[2];
    gL[5] += fT0 * yd[0] + fT1 * yd[1] + fT2 * yd[2];
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { sumy += fabs(yd[j]); ceq_th += fabs(sv.cd[j]); ceq_max = fmax(ceq_max, fabs(sv.cd[j])); }
    }
    // (7) input bounds and acceleration rows
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double sg, ta, tb;
      acc.add(w.W(L.SUB, j), w.W(L.ZUB, j), sv.dub[j], sg, ta, tb);
      HH(6 + j, 6 + j) += sg; ra[6 + j] += ta; rb[6 + j] += tb; gL[6 + j] -= w.W(L.ZUB, j);
      acc.add(w.W(L.SUB, 2 + j), w.W(L.ZUB, 2 + j), sv.dub[2 + j], sg, ta, tb);
      HH(6 + j, 6 + j) += sg; ra[6 + j] -= ta; rb[6 + j] -= tb; gL[6 + j] += w.W(L.ZUB, 2 + j);
      double ga = (up[j] - u[j]) / h;
      double s4, ta4, tb4, s6, ta6, tb6;
      const double Z4 = w.W(L.ZUB, 4 + j), Z6 = w.W(L.ZUB, 6 + j);
      acc.add(w.W(L.SUB, 4 + j), Z4, sv.dub[4 + j], s4, ta4, tb4);
      acc.add(w.W(L.SUB, 6 + j), Z6, sv.dub[6 + j], s6, ta6, tb6);
      const double jv[3] = {(k >= 1) ? 1.0 / h : 0.0, -1.0 / h, free_ ? -ga / T : 0.0};
      const int ix[3] = {3 + j, 6 + j, 5};
      const double ss = s4 + s6, tta = ta4 - ta6, ttb = tb4 - tb6, zz = Z4 - Z6;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ra[ix[a]] += tta * jv[a]; rb[ix[a]] += ttb * jv[a];
        gL[ix[a]] -= zz * jv[a];
#pragma unroll
        for (int b = 0; b < 3; ++b)
          if (ix[b] <= ix[a]) HH(ix[a], ix[b]) += ss * jv[a] * jv[b];
      }
      if (free_) {
        double yj = -zz, c = yj / (h * T);
        HH(6 + j, 5) += c;
        if (k >= 1) HH(5, 3 + j) -= c;
        HH(5, 5) += yj * 2 * ga / (T * T);
      }
    }
  }
  if (k >= 1) {
#pragma unroll
    for (int j = 0; j < 3; ++j) gL[j] -= ydm[j];
    // (6) state bounds
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double sg, ta, tb;
      acc.add(w.W(L.SXY, j), w.W(L.ZXY, j), sv.dxy[j], sg, ta, tb);
      HH(j, j) += sg; ra[j] += ta; rb[j] += tb; gL[j] -= w.W(L.ZXY, j);
      acc.add(w.W(L.SXY, 2 + j), w.W(L.ZXY, 2 + j), sv.dxy[2 + j], sg, ta, tb);
      HH(j, j) += sg; ra[j] -= ta; rb[j] -= tb; gL[j] += w.W(L.ZXY, 2 + j);
    }
  }
  double ctmax = 0;
  if (k == 0 && free_) {
    // (4) time cost and (8) T bounds live in stage 0
    gf[5] += (N + 1) * (P.time_cost[0] + 2 * P.time_cost[1] * T);
    HH(5, 5) += 2 * (N + 1) * P.time_cost[1];
    double sg, ta, tb;
    acc.add(G.STb[0], G.ZTb[0], T - P.T_min, sg, ta, tb);
    HH(5, 5) += sg; ra[5] += ta; rb[5] += tb; gL[5] -= G.ZTb[0];
    acc.add(G.STb[1], G.ZTb[1], G.Tmax - T, sg, ta, tb);
    HH(5, 5) += sg; ra[5] -= ta; rb[5] -= tb; gL[5] += G.ZTb[1];
  }
  if (k == N && w.has_term) {
    double sg, ta, tb;
    acc.add(G.Stm[0], G.Ztm[0], z[0] - G.term[0], sg, ta, tb);
    HH(0, 0) += sg; ra[0] += ta; rb[0] += tb; gL[0] -= G.Ztm[0];
    acc.add(G.Stm[1], G.Ztm[1], z[1] - G.term[1], sg, ta, tb);
    HH(1, 1) += sg; ra[1] += ta; rb[1] += tb; gL[1] -= G.Ztm[1];
    acc.add(G.Stm[2], G.Ztm[2], G.term[2] - z[1], sg, ta, tb);
    HH(1, 1) += sg; ra[1] -= ta; rb[1] -= tb; gL[1] += G.Ztm[2];
  }
  if (k == N && free_) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      gL[j] += G.yt[j];
      double c = z[j] - w.xref[3 * N + j];
      ceq_th += fabs(c); ceq_max = fmax(ceq_max, fabs(c)); ctmax = fmax(ctmax, fabs(c));
      sumy += fabs(G.yt[j]);
    }
  }

  // (11) obstacle blocks
  bool ok = true;
  BlkGeo b;
  b.ct = ct; b.st = st; b.tx = z[0] + G.off * ct; b.ty = z[1] + G.off * st;
  for (int i = 0; i < w.no; ++i) {
    const int r0 = w.kp.eptr[i], E = w.kp.eptr[i + 1] - r0;
    double a1 = 0, a2 = 0, bl = 0;
    for (int r = r0; r < r0 + E; ++r) {
      double l = w.W(L.LAM, r);
      a1 += w.A[2 * r] * l; a2 += w.A[2 * r + 1] * l; bl += w.bk(r) * l;
    }
    b.a1 = a1; b.a2 = a2;
    const double m0 = w.W(L.MU, 4 * i), m1 = w.W(L.MU, 4 * i + 1), m2 = w.W(L.MU, 4 * i + 2), m3 = w.W(L.MU, 4 * i + 3);
    const double ce1 = m0 - m2 + ct * a1 + st * a2, ce2 = m1 - m3 - st * a1 + ct * a2;
    const double dn = 1.0 - a1 * a1 - a2 * a2;
    const double dd = -(G.g[0] * m0 + G.g[1] * m1 + G.g[2] * m2 + G.g[3] * m3) + b.tx * a1 + b.ty * a2 - bl - P.dmin;
    const double y1 = w.W(L.YE, 2 * i), y2 = w.W(L.YE, 2 * i + 1);
    const double Sn = w.W(L.SN, i), Zn = w.W(L.ZN, i), Sd = w.W(L.SD, i), Zd = w.W(L.ZD, i);
    ceq_th += fabs(ce1) + fabs(ce2); ceq_max = fmax(ceq_max, fmax(fabs(ce1), fabs(ce2)));
    sumy += fabs(y1) + fabs(y2);
    double sn, tna, tnb, sd, tda, tdb;
    acc.add(Sn, Zn, dn, sn, tna, tnb);
    acc.add(Sd, Zd, dd, sd, tda, tdb);
    // rows
    Tri5 Lf;
    Lf.zero();
    double g0a[5] = {0, 0, 0, 0, 0}, g0b[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < E + 4; ++j) {
      double wv, S, Z, yv[5], yh[5];
      if (j < E) { wv = w.W(L.LAM, r0 + j); S = w.W(L.SL, r0 + j); Z = w.W(L.ZL, r0 + j); }
      else { wv = w.W(L.MU, 4 * i + j - E); S = w.W(L.SM, 4 * i + j - E); Z = w.W(L.ZM, 4 * i + j - E); }
      row_y(w, G, b, r0, E, j, yv);
      double sig, ta, tb;
      acc.add(S, Z, wv, sig, ta, tb);
      const double gl = y1 * yv[3] + y2 * yv[4] - Z + 2 * Zn * (a1 * yv[0] + a2 * yv[1]) - Zd * yv[2];
      e1 = fmax(e1, fabs(gl));
      tb -= gl;
      const double di = 1.0 / fmax(sig, SIG_MIN), sq = sqrt(di);
#pragma unroll
      for (int a = 0; a < 5; ++a) { g0a[a] += yv[a] * ta * di; g0b[a] += yv[a] * tb * di; yh[a] = yv[a] * sq; }
      Lf.insert(yh);
    }
    const double aa = a1 * a1 + a2 * a2, lam1 = 2 * Zn + 4 * sn * aa;
    const double ci0 = 1.0 / (2 * Zn), ci1 = 4 * sn / lam1;
    {
      double r1[5] = {0, 0, 0, 0, 0}, r2[5] = {0, 0, 0, 0, 0}, r3[5] = {0, 0, sqrt(1.0 / sd), 0, 0};
      if (aa > 0) {
        double na = sqrt(aa), e1_ = a1 / na, e2_ = a2 / na, s1 = sqrt(1.0 / lam1), s2 = sqrt(ci0);
        r1[0] = s1 * e1_; r1[1] = s1 * e2_; r2[0] = -s2 * e2_; r2[1] = s2 * e1_;
      } else {
        r1[0] = sqrt(ci0); r2[1] = r1[0];
      }
      Lf.insert(r1); Lf.insert(r2); Lf.insert(r3);
      ok = Lf.finish() && ok;
    }
    // column 0 (split in mu): h0 = -2 tn a,  rhs = (g0 - Cn^-1 h0, g0[2] - tds, g0[3] + e1, g0[4] + e2),
    // tds = mu / Zd - dd
    double h0a[2] = {-2 * tna * a1, -2 * tna * a2}, h0b[2] = {-2 * tnb * a1, -2 * tnb * a2};
    double cha[2], chb[2], rhs[5], eta0a[5], eta0b[5];
    cn_inv(b, ci0, ci1, h0a, cha);
    cn_inv(b, ci0, ci1, h0b, chb);
    rhs[0] = g0a[0] - cha[0]; rhs[1] = g0a[1] - cha[1]; rhs[2] = g0a[2] - 1.0 / Zd; rhs[3] = g0a[3]; rhs[4] = g0a[4];
    Lf.solve(rhs, eta0a);
    rhs[0] = g0b[0] - chb[0]; rhs[1] = g0b[1] - chb[1]; rhs[2] = g0b[2] + dd; rhs[3] = g0b[3] + ce1; rhs[4] = g0b[4] + ce2;
    Lf.solve(rhs, eta0b);
#pragma unroll
    for (int a = 0; a < 5; ++a) { w.W(L.ETA, 25 * i + a) = eta0a[a]; w.W(L.ETA, 25 * i + 5 + a) = eta0b[a]; }
    // Lagrangian gradient wrt the pose from this block
    const double offt = G.off * (-st * a1 + ct * a2);
    const double dpose[3] = {a1, a2, offt};
    const double jt0 = -st * a1 + ct * a2, jt1 = -ct * a1 - st * a2;
    gL[2] += y1 * jt0 + y2 * jt1;
#pragma unroll
    for (int a = 0; a < 3; ++a) gL[a] -= Zd * dpose[a];
    // pose columns and Schur complement (pose of stage 0 is fixed: columns unused but harmless)
    const double ydv = -Zd, c1 = y1 + ydv * G.off;
    const double hc[3][2] = {{ydv, 0.0}, {0.0, ydv}, {-c1 * st - y2 * ct, c1 * ct - y2 * st}};
    double bc[3][5], etc[3][5];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cn_inv(b, ci0, ci1, hc[c], bc[c]);
      bc[c][2] = dpose[c];
      bc[c][3] = (c == 2) ? jt0 : 0.0;
      bc[c][4] = (c == 2) ? jt1 : 0.0;
      Lf.solve(bc[c], etc[c]);
#pragma unroll
      for (int a = 0; a < 5; ++a) w.W(L.ETA, 25 * i + 10 + 5 * c + a) = etc[c][a];
    }
    HH(2, 2) += y1 * (-ct * a1 - st * a2) + y2 * (st * a1 - ct * a2) + ydv * G.off * (-ct * a1 - st * a2);
#pragma unroll
    for (int cp = 0; cp < 3; ++cp) {
#pragma unroll
      for (int c = 0; c <= cp; ++c) {
        double Gm = hc[cp][0] * bc[c][0] + hc[cp][1] * bc[c][1];
#pragma unroll
        for (int a = 0; a < 5; ++a) Gm -= bc[cp][a] * etc[c][a];
        HH(cp, c) -= Gm;
      }
      double Ga = -(bc[cp][0] * h0a[0] + bc[cp][1] * h0a[1]), Gb = -(bc[cp][0] * h0b[0] + bc[cp][1] * h0b[1]);
#pragma unroll
      for (int a = 0; a < 5; ++a) { Ga -= bc[cp][a] * eta0a[a]; Gb -= bc[cp][a] * eta0b[a]; }
      ra[cp] += Ga; rb[cp] += Gb;
    }
  }
  // step form: r = -grad L + Jd^T (t - Z)
#pragma unroll
  for (int a = 0; a < 8; ++a) { gL[a] += gf[a]; rb[a] -= gL[a]; }
  // optimality error pieces
  double gun[2] = {sh_dn(gL[3]), sh_dn(gL[4])};
  if (act) {
    if (k >= 1) e1 = fmax(e1, fmax(fabs(gL[0]), fmax(fabs(gL[1]), fabs(gL[2]))));
    if (k < N) {
#pragma unroll
      for (int j = 0; j < 2; ++j) e1 = fmax(e1, fabs(gL[6 + j] + ((k + 1 < N) ? gun[j] : 0.0)));
    }
  }
  double gT = wsum(act ? gL[5] : 0.0);
  E.e1 = wmax(act ? e1 : 0.0);
  if (free_) E.e1 = fmax(E.e1, fabs(gT));
  E.f = wsum(act ? sv.f : 0.0);
  E.th = wsum(act ? acc.th + ceq_th : 0.0);
  E.lgS = wsum(act ? acc.lg : 0.0);
  E.e2 = wmax(act ? fmax(acc.cmax, ceq_max) : 0.0);
  E.sumy = wsum(act ? sumy : 0.0);
  E.sumz = wsum(act ? acc.sumz : 0.0);
  E.szmax = wmax(act ? acc.szmax : 0.0);
  E.szmin = wmin(act ? acc.szmin : 1e300);
  E.ctmax = wmax(act ? ctmax : 0.0);
  E.ok = __all_sync(FULL, ok || !act);
  // store the stage QP
#pragma unroll
  for (int i = 0; i < 36; ++i) w.W(L.H, i) = H[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) { w.W(L.RA, i) = ra[i]; w.W(L.RB, i) = rb[i]; w.W(L.DSUB, i) = gf[i]; }
#pragma unroll
  for (int j = 0; j < 3; ++j) w.W(L.CD, j) = (k < N) ? sv.cd[j] : 0.0;
#undef HH
}

// ------------------------------------------------------------------------------------------------------
// Riccati recursion over (x, y, th, v_prev, w_prev, T | v, w); lane s works at step s, the cost-to-go is
// handed over through the workspace column of lane s+1.  Returns false if a pivot is not positive definite.
// r = mu * ra + rb is formed here.
// ------------------------------------------------------------------------------------------------------
__device__ bool riccati(const Warp& w, const Glob& G, double mu, double dw, double dc) {
  const obca_params& P = w.kp.P;
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  const bool free_ = w.free_;
  bool bad = false;
  for (int s = N; s >= 0; --s) {
    if (k == s) {
      double H[36], r[8];
#pragma unroll
      for (int i = 0; i < 36; ++i) H[i] = w.W(L.H, i);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = mu * w.W(L.RA, i) + w.W(L.RB, i);
#define HH(a, b) H[((a) >= (b)) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a))]
      if (s == N) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          HH(a, a) += dw;
          if (free_) { HH(a, a) += 1.0 / dc; r[a] -= (w.W(L.Z, a) - w.xref[3 * N + a]) / dc; }
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
          for (int b = 0; b <= a; ++b) w.W(L.PM, a * (a + 1) / 2 + b) = HH(a, b);
          w.W(L.PV, a) = r[a];
        }
      } else {
        double Pn[21], pn[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) Pn[i] = (&w.W(L.PM, i))[1];
#pragma unroll
        for (int i = 0; i < 6; ++i) pn[i] = (&w.W(L.PV, i))[1];
#define PP(a, b) Pn[((a) >= (b)) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a))]
        const double z2 = w.W(L.Z, 2), vv = w.W(L.U, 0), ww = w.W(L.U, 1);
        const double T = free_ ? G.T : 1.0, h = T * P.Ts;
        double st, ct;
        sincos(z2, &st, &ct);
        const double fth0 = -h * vv * st, fth1 = h * vv * ct, bv0 = h * ct, bv1 = h * st, bw = h;
        const double fT0 = free_ ? P.Ts * vv * ct : 0.0, fT1 = free_ ? P.Ts * vv * st : 0.0, fT2 = free_ ? P.Ts * ww : 0.0;
        // columns m_j = At e_j of At = [A B]; Pm_j = P m_j (6-vectors); columns 3,4 (v_prev, w_prev) are zero
        double Pm[8][6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          Pm[0][a] = PP(a, 0);
          Pm[1][a] = PP(a, 1);
          Pm[2][a] = fth0 * PP(a, 0) + fth1 * PP(a, 1) + PP(a, 2);
          Pm[3][a] = 0; Pm[4][a] = 0;
          Pm[5][a] = fT0 * PP(a, 0) + fT1 * PP(a, 1) + fT2 * PP(a, 2) + PP(a, 5);
          Pm[6][a] = bv0 * PP(a, 0) + bv1 * PP(a, 1) + PP(a, 3);
          Pm[7][a] = bw * PP(a, 2) + PP(a, 4);
        }
        double pc[6];
        const double c0 = w.W(L.CD, 0), c1 = w.W(L.CD, 1), c2 = w.W(L.CD, 2);
#pragma unroll
        for (int a = 0; a < 6; ++a) pc[a] = pn[a] - (PP(a, 0) * c0 + PP(a, 1) * c1 + PP(a, 2) * c2);
        // F = H + At^T P At,  f = r + At^T pc     (m_i^T v with the sparse m_i written out)
        auto mdot = [&](int i, const double* v) -> double {
          switch (i) {
            case 0: return v[0];
            case 1: return v[1];
            case 2: return fth0 * v[0] + fth1 * v[1] + v[2];
            case 5: return fT0 * v[0] + fT1 * v[1] + fT2 * v[2] + v[5];
            case 6: return bv0 * v[0] + bv1 * v[1] + v[3];
            case 7: return bw * v[2] + v[4];
            default: return 0.0;
          }
        };
        double F[36], f[8];
#define FF(a, b) F[((a) >= (b)) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a))]
#pragma unroll
        for (int a = 0; a < 8; ++a) {
#pragma unroll
          for (int b = 0; b <= a; ++b) FF(a, b) = HH(a, b) + mdot(a, Pm[b]);
          f[a] = r[a] + mdot(a, pc);
        }
        if (s >= 1) { FF(0, 0) += dw; FF(1, 1) += dw; FF(2, 2) += dw; }
        FF(6, 6) += dw; FF(7, 7) += dw;
        if (s == 0 && free_) FF(5, 5) += dw;
        const double q00 = FF(6, 6), q01 = FF(7, 6), q11 = FF(7, 7);
        const double det = q00 * q11 - q01 * q01;
        if (!(q00 > 0) || !(det > 0)) bad = true;
        const double i00 = q11 / det, i01 = -q01 / det, i11 = q00 / det;
        double K0[6], K1[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          K0[b] = -(i00 * FF(6, b) + i01 * FF(7, b));
          K1[b] = -(i01 * FF(6, b) + i11 * FF(7, b));
          w.W(L.K, b) = K0[b]; w.W(L.K, 6 + b) = K1[b];
        }
        const double kap0 = i00 * f[6] + i01 * f[7], kap1 = i01 * f[6] + i11 * f[7];
        w.W(L.KAP, 0) = kap0; w.W(L.KAP, 1) = kap1;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            double pab = FF(a, b) + FF(a, 6) * K0[b] + FF(a, 7) * K1[b];
            double pba = FF(b, a) + FF(b, 6) * K0[a] + FF(b, 7) * K1[a];
            w.W(L.PM, a * (a + 1) / 2 + b) = 0.5 * (pab + pba);
          }
          w.W(L.PV, a) = f[a] - FF(a, 6) * kap0 - FF(a, 7) * kap1;
        }
#undef FF
#undef PP
      }
#undef HH
    }
    __syncwarp();
  }
  if (k == 0 && free_ && !(w.W(L.PM, 20) > 0)) bad = true;
  return !__any_sync(FULL, bad);
}

// forward roll-out; the multiplier steps of the dynamics follow lane-parallel from the cost-to-go
__device__ void forward(const Warp& w, Glob& G, double dc) {
  const obca_params& P = w.kp.P;
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  const bool free_ = w.free_;
  double xi[6] = {0, 0, 0, 0, 0, 0};
  double dT = 0;
  if (free_) dT = bcast(w.W(L.PV, 5) / w.W(L.PM, 20), 0);
  G.dT = dT;
  xi[5] = dT;
  double mine[6] = {0, 0, 0, 0, 0, dT};
  const double T = free_ ? G.T : 1.0, h = T * P.Ts;
  const double z2 = w.W(L.Z, 2), vv = w.W(L.U, 0), ww = w.W(L.U, 1);
  double st, ct;
  sincos(z2, &st, &ct);
  const double fth0 = -h * vv * st, fth1 = h * vv * ct, bv0 = h * ct, bv1 = h * st, bw = h;
  const double fT0 = free_ ? P.Ts * vv * ct : 0.0, fT1 = free_ ? P.Ts * vv * st : 0.0, fT2 = free_ ? P.Ts * ww : 0.0;
  for (int s = 0; s < N; ++s) {
    double xn[6] = {0, 0, 0, 0, 0, 0};
    if (k == s) {
      double du0 = w.W(L.KAP, 0), du1 = w.W(L.KAP, 1);
#pragma unroll
      for (int b = 0; b < 6; ++b) { du0 += w.W(L.K, b) * xi[b]; du1 += w.W(L.K, 6 + b) * xi[b]; }
      w.W(L.DU, 0) = du0; w.W(L.DU, 1) = du1;
      xn[0] = xi[0] + fth0 * xi[2] + fT0 * xi[5] + bv0 * du0 + w.W(L.CD, 0);
      xn[1] = xi[1] + fth1 * xi[2] + fT1 * xi[5] + bv1 * du0 + w.W(L.CD, 1);
      xn[2] = xi[2] + fT2 * xi[5] + bw * du1 + w.W(L.CD, 2);
      xn[3] = du0; xn[4] = du1; xn[5] = xi[5];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) xi[a] = bcast(xn[a], s);
    if (k == s + 1) {
#pragma unroll
      for (int a = 0; a < 6; ++a) mine[a] = xi[a];
    }
  }
  // lane k holds d xi_k in `mine`; dy_{k-1} = (P_k d xi_k - p_k)[0:3] goes to column k-1
#pragma unroll
  for (int a = 0; a < 3; ++a) w.W(L.DZ, a) = mine[a];
  if (k >= 1 && w.act) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double sacc = -w.W(L.PV, a);
#pragma unroll
      for (int b = 0; b < 6; ++b) sacc += w.W(L.PM, symi(a, b)) * mine[b];
      (&w.W(L.DYD, a))[-1] = sacc;
    }
  }
  if (free_) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double v = (mine[a] + (w.W(L.Z, a) - w.xref[3 * N + a])) / dc;
      G.dyt[a] = bcast(v, N);
    }
  }
  __syncwarp();
}

struct StepInfo {
  double a_max, a_z, Dphi;
};

__device__ __forceinline__ void ftb(double S, double Z, double dS, double mu, double tau, double& amax, double& az, double& sls) {
  double dZ = mu / S - Z - (Z / S) * dS;
  if (dS < 0) amax = fmin(amax, -tau * S / dS);
  if (dZ < 0) az = fmin(az, -tau * Z / dZ);
  sls += dS / S;
}

// ------------------------------------------------------------------------------------------------------
// back-substitution of the dual blocks, slack steps, fraction to the boundary, directional derivative
// ------------------------------------------------------------------------------------------------------
__device__ void backsub(const Warp& w, Glob& G, double mu, double tau, StepInfo& si) {
  const obca_params& P = w.kp.P;
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  const bool act = w.act, free_ = w.free_;
  double z[3], u[2], up[2], zn[3], dz[3], du[2], dup[2];
#pragma unroll
  for (int j = 0; j < 3; ++j) { z[j] = w.W(L.Z, j); dz[j] = w.W(L.DZ, j); }
#pragma unroll
  for (int j = 0; j < 2; ++j) { u[j] = w.W(L.U, j); du[j] = w.W(L.DU, j); }
#pragma unroll
  for (int j = 0; j < 3; ++j) zn[j] = sh_dn(z[j]);
#pragma unroll
  for (int j = 0; j < 2; ++j) { up[j] = sh_up(u[j]); dup[j] = sh_up(du[j]); if (k == 0) { up[j] = G.u0[j]; dup[j] = 0.0; } }
  const double T = free_ ? G.T : 1.0, h = T * P.Ts, dT = G.dT;
  double st, ct;
  sincos(z[2], &st, &ct);
  StageVals sv;
  stage_vals(w, G, z, u, up, zn, T, sv);
  double amax = 1.0, az = 1.0, sls = 0.0, dphi = 0.0;
  // objective directional derivative: gf over (z_k, u_{k-1}, T, u_k) was parked in DSUB by assemble
  {
    double gf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gf[i] = w.W(L.DSUB, i);
    if (k >= 1) dphi += gf[0] * dz[0] + gf[1] * dz[1] + gf[2] * dz[2];
    if (k >= 1 && k < N) dphi += gf[3] * dup[0] + gf[4] * dup[1];
    if (free_) dphi += gf[5] * dT;
    if (k < N) dphi += gf[6] * du[0] + gf[7] * du[1];
  }
  if (k >= 1) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double S0 = w.W(L.SXY, j), S1 = w.W(L.SXY, 2 + j);
      double d0 = dz[j] + (sv.dxy[j] - S0), d1 = -dz[j] + (sv.dxy[2 + j] - S1);
      w.W(L.DSXY, j) = d0; w.W(L.DSXY, 2 + j) = d1;
      if (act) { ftb(S0, w.W(L.ZXY, j), d0, mu, tau, amax, az, sls); ftb(S1, w.W(L.ZXY, 2 + j), d1, mu, tau, amax, az, sls); }
    }
  }
  if (k < N) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double ga = (up[j] - u[j]) / h;
      double dga = (dup[j] - du[j]) / h - (free_ ? ga / T * dT : 0.0);
      double dS[4] = {du[j] + (sv.dub[j] - w.W(L.SUB, j)), -du[j] + (sv.dub[2 + j] - w.W(L.SUB, 2 + j)),
                      dga + (sv.dub[4 + j] - w.W(L.SUB, 4 + j)), -dga + (sv.dub[6 + j] - w.W(L.SUB, 6 + j))};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        w.W(L.DSUB, 2 * q + j) = dS[q];
        ftb(w.W(L.SUB, 2 * q + j), w.W(L.ZUB, 2 * q + j), dS[q], mu, tau, amax, az, sls);
      }
    }
  }
  if (free_) {
    G.dSTb[0] = dT + ((T - P.T_min) - G.STb[0]);
    G.dSTb[1] = -dT + ((G.Tmax - T) - G.STb[1]);
    if (k == 0) { ftb(G.STb[0], G.ZTb[0], G.dSTb[0], mu, tau, amax, az, sls); ftb(G.STb[1], G.ZTb[1], G.dSTb[1], mu, tau, amax, az, sls); }
  }
  if (w.has_term) {
    double zN0 = bcast(z[0], N), zN1 = bcast(z[1], N), dzN0 = bcast(dz[0], N), dzN1 = bcast(dz[1], N);
    G.dStm[0] = dzN0 + ((zN0 - G.term[0]) - G.Stm[0]);
    G.dStm[1] = dzN1 + ((zN1 - G.term[1]) - G.Stm[1]);
    G.dStm[2] = -dzN1 + ((G.term[2] - zN1) - G.Stm[2]);
    if (k == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) ftb(G.Stm[j], G.Ztm[j], G.dStm[j], mu, tau, amax, az, sls);
    }
  }
  // dual blocks
  BlkGeo b;
  b.ct = ct; b.st = st; b.tx = z[0] + G.off * ct; b.ty = z[1] + G.off * st;
  const double dp[3] = {(k >= 1) ? dz[0] : 0.0, (k >= 1) ? dz[1] : 0.0, (k >= 1) ? dz[2] : 0.0};
  for (int i = 0; i < w.no; ++i) {
    const int r0 = w.kp.eptr[i], E = w.kp.eptr[i + 1] - r0;
    double a1 = 0, a2 = 0, bl = 0;
    for (int r = r0; r < r0 + E; ++r) {
      double l = w.W(L.LAM, r);
      a1 += w.A[2 * r] * l; a2 += w.A[2 * r + 1] * l; bl += w.bk(r) * l;
    }
    b.a1 = a1; b.a2 = a2;
    const double m0 = w.W(L.MU, 4 * i), m1 = w.W(L.MU, 4 * i + 1), m2 = w.W(L.MU, 4 * i + 2), m3 = w.W(L.MU, 4 * i + 3);
    const double dn = 1.0 - a1 * a1 - a2 * a2;
    const double dd = -(G.g[0] * m0 + G.g[1] * m1 + G.g[2] * m2 + G.g[3] * m3) + b.tx * a1 + b.ty * a2 - bl - P.dmin;
    const double y1 = w.W(L.YE, 2 * i), y2 = w.W(L.YE, 2 * i + 1);
    const double Sn = w.W(L.SN, i), Zn = w.W(L.ZN, i), Sd = w.W(L.SD, i), Zd = w.W(L.ZD, i);
    const double sn = Zn / Sn, sd = Zd / Sd;
    const double tn = (mu - Sn * Zn) / Sn - sn * (dn - Sn);
    const double ydv = -Zd, c1 = y1 + ydv * G.off;
    const double hc[3][2] = {{ydv, 0.0}, {0.0, ydv}, {-c1 * st - y2 * ct, c1 * ct - y2 * st}};
    const double offt = G.off * (-st * a1 + ct * a2);
    const double dpose[3] = {a1, a2, offt};
    double et[5], ht[2] = {-2 * tn * a1, -2 * tn * a2};
#pragma unroll
    for (int a = 0; a < 5; ++a) et[a] = mu * w.W(L.ETA, 25 * i + a) + w.W(L.ETA, 25 * i + 5 + a);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int a = 0; a < 5; ++a) et[a] += w.W(L.ETA, 25 * i + 10 + 5 * c + a) * dp[c];
      ht[0] -= hc[c][0] * dp[c]; ht[1] -= hc[c][1] * dp[c];
    }
    double da1 = 0, da2 = 0, qdw = 0;
    for (int j = 0; j < E + 4; ++j) {
      double wv, S, Z, yv[5];
      if (j < E) { wv = w.W(L.LAM, r0 + j); S = w.W(L.SL, r0 + j); Z = w.W(L.ZL, r0 + j); }
      else { wv = w.W(L.MU, 4 * i + j - E); S = w.W(L.SM, 4 * i + j - E); Z = w.W(L.ZM, 4 * i + j - E); }
      row_y(w, G, b, r0, E, j, yv);
      const double sig = Z / S;
      const double gl = y1 * yv[3] + y2 * yv[4] - Z + 2 * Zn * (a1 * yv[0] + a2 * yv[1]) - Zd * yv[2];
      double s = (mu - S * Z) / S - sig * (wv - S) - gl;
#pragma unroll
      for (int a = 0; a < 5; ++a) s -= yv[a] * et[a];
      const double dwv = s / fmax(sig, SIG_MIN);
      if (j < E) { w.W(L.DLAM, r0 + j) = dwv; da1 += yv[0] * dwv; da2 += yv[1] * dwv; }
      else w.W(L.DMU, 4 * i + j - E) = dwv;
      qdw += yv[2] * dwv;
      if (act) ftb(S, Z, dwv + (wv - S), mu, tau, amax, az, sls);
    }
    w.W(L.DYE, 2 * i) = et[3]; w.W(L.DYE, 2 * i + 1) = et[4];
    double ada = a1 * da1 + a2 * da2;
    if (sn >= 1.0) ada = (a1 * (et[0] + ht[0]) + a2 * (et[1] + ht[1])) / (2 * Zn + 4 * sn * (a1 * a1 + a2 * a2));
    const double dSn = -2 * ada + (dn - Sn);
    double dSd;
    if (sd >= 1.0) dSd = (mu - Sd * Zd + Sd * et[2]) / Zd;
    else dSd = qdw + (dd - Sd) + dpose[0] * dp[0] + dpose[1] * dp[1] + dpose[2] * dp[2];
    w.W(L.DSN, i) = dSn; w.W(L.DSD, i) = dSd;
    if (act) { ftb(Sn, Zn, dSn, mu, tau, amax, az, sls); ftb(Sd, Zd, dSd, mu, tau, amax, az, sls); }
  }
  si.a_max = wmin(act ? amax : 1.0);
  si.a_z = wmin(act ? az : 1.0);
  si.Dphi = wsum(act ? dphi - mu * sls : 0.0);
}

// ------------------------------------------------------------------------------------------------------
// trial point X + a dX, S + a dS: constraint violation theta and barrier function phi
// ------------------------------------------------------------------------------------------------------
__device__ void trial(const Warp& w, const Glob& G, double a, double mu, double& th_out, double& ph_out) {
  const obca_params& P = w.kp.P;
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  const bool act = w.act, free_ = w.free_;
  double z[3], u[2], up[2], zn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) z[j] = w.W(L.Z, j) + ((k >= 1) ? a * w.W(L.DZ, j) : 0.0);
#pragma unroll
  for (int j = 0; j < 2; ++j) u[j] = w.W(L.U, j) + a * w.W(L.DU, j);
#pragma unroll
  for (int j = 0; j < 3; ++j) zn[j] = sh_dn(z[j]);
#pragma unroll
  for (int j = 0; j < 2; ++j) { up[j] = sh_up(u[j]); if (k == 0) up[j] = G.u0[j]; }
  const double T = free_ ? G.T + a * G.dT : 1.0;
  StageVals sv;
  stage_vals(w, G, z, u, up, zn, T, sv);
  double th = 0, lg = 0;
  auto ineq = [&](double d, double S) { th += fabs(d - S); lg += log(S); };
  if (k < N) {
    th += fabs(sv.cd[0]) + fabs(sv.cd[1]) + fabs(sv.cd[2]);
#pragma unroll
    for (int j = 0; j < 8; ++j) ineq(sv.dub[j], w.W(L.SUB, j) + a * w.W(L.DSUB, j));
  }
  if (k >= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) ineq(sv.dxy[j], w.W(L.SXY, j) + a * w.W(L.DSXY, j));
  }
  if (k == 0 && free_) {
    ineq(T - P.T_min, G.STb[0] + a * G.dSTb[0]);
    ineq(G.Tmax - T, G.STb[1] + a * G.dSTb[1]);
  }
  if (k == N && free_) {
#pragma unroll
    for (int j = 0; j < 3; ++j) th += fabs(z[j] - w.xref[3 * N + j]);
  }
  if (k == N && w.has_term) {
    ineq(z[0] - G.term[0], G.Stm[0] + a * G.dStm[0]);
    ineq(z[1] - G.term[1], G.Stm[1] + a * G.dStm[1]);
    ineq(G.term[2] - z[1], G.Stm[2] + a * G.dStm[2]);
  }
  double st, ct;
  sincos(z[2], &st, &ct);
  const double tx = z[0] + G.off * ct, ty = z[1] + G.off * st;
  for (int i = 0; i < w.no; ++i) {
    const int r0 = w.kp.eptr[i], E = w.kp.eptr[i + 1] - r0;
    double a1 = 0, a2 = 0, bl = 0;
    for (int r = r0; r < r0 + E; ++r) {
      const double l0 = w.W(L.LAM, r), dl = w.W(L.DLAM, r), S0 = w.W(L.SL, r);
      const double l = l0 + a * dl, S = S0 + a * (dl + (l0 - S0));
      a1 += w.A[2 * r] * l; a2 += w.A[2 * r + 1] * l; bl += w.bk(r) * l;
      ineq(l, S);
    }
    double m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double m0 = w.W(L.MU, 4 * i + q), dm = w.W(L.DMU, 4 * i + q), S0 = w.W(L.SM, 4 * i + q);
      m[q] = m0 + a * dm;
      ineq(m[q], S0 + a * (dm + (m0 - S0)));
    }
    th += fabs(m[0] - m[2] + ct * a1 + st * a2) + fabs(m[1] - m[3] - st * a1 + ct * a2);
    ineq(1.0 - a1 * a1 - a2 * a2, w.W(L.SN, i) + a * w.W(L.DSN, i));
    ineq(-(G.g[0] * m[0] + G.g[1] * m[1] + G.g[2] * m[2] + G.g[3] * m[3]) + tx * a1 + ty * a2 - bl - P.dmin,
         w.W(L.SD, i) + a * w.W(L.DSD, i));
  }
  const double f = wsum(act ? sv.f : 0.0);
  th_out = wsum(act ? th : 0.0);
  ph_out = f - mu * wsum(act ? lg : 0.0);
}

// ------------------------------------------------------------------------------------------------------
// accept the step: primal / slacks / equality multipliers with a, inequality multipliers with a_z (then
// clipped into [mu/(ks S), ks mu/S] as IPOPT does)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void upd(double& S, double& Z, double dS, double a, double az, double mu) {
  const double ks = 1e10;
  double dZ = mu / S - Z - (Z / S) * dS;
  double Sn = S + a * dS, Zn = Z + az * dZ;
  Zn = fmin(fmax(Zn, mu / (ks * Sn)), ks * mu / Sn);
  S = Sn; Z = Zn;
}

__device__ void update(const Warp& w, Glob& G, double a, double az, double mu) {
  const Lay& L = w.L;
  const int k = w.k, N = w.N;
  const bool free_ = w.free_;
  auto updw = [&](int oS, int oZ, int j, double dS) {
    double S = w.W(oS, j), Z = w.W(oZ, j);
    upd(S, Z, dS, a, az, mu);
    w.W(oS, j) = S; w.W(oZ, j) = Z;
  };
  if (k >= 1) {
#pragma unroll
    for (int j = 0; j < 3; ++j) w.W(L.Z, j) += a * w.W(L.DZ, j);
#pragma unroll
    for (int j = 0; j < 4; ++j) updw(L.SXY, L.ZXY, j, w.W(L.DSXY, j));
  }
  if (k < N) {
#pragma unroll
    for (int j = 0; j < 2; ++j) w.W(L.U, j) += a * w.W(L.DU, j);
#pragma unroll
    for (int j = 0; j < 8; ++j) updw(L.SUB, L.ZUB, j, w.W(L.DSUB, j));
#pragma unroll
    for (int j = 0; j < 3; ++j) w.W(L.YD, j) += a * w.W(L.DYD, j);
  }
  for (int r = 0; r < w.R; ++r) {
    const double l0 = w.W(L.LAM, r), dl = w.W(L.DLAM, r);
    updw(L.SL, L.ZL, r, dl + (l0 - w.W(L.SL, r)));
    w.W(L.LAM, r) = l0 + a * dl;
  }
  for (int r = 0; r < 4 * w.no; ++r) {
    const double m0 = w.W(L.MU, r), dm = w.W(L.DMU, r);
    updw(L.SM, L.ZM, r, dm + (m0 - w.W(L.SM, r)));
    w.W(L.MU, r) = m0 + a * dm;
  }
  for (int i = 0; i < w.no; ++i) {
    updw(L.SN, L.ZN, i, w.W(L.DSN, i));
    updw(L.SD, L.ZD, i, w.W(L.DSD, i));
    w.W(L.YE, 2 * i) += a * w.W(L.DYE, 2 * i);
    w.W(L.YE, 2 * i + 1) += a * w.W(L.DYE, 2 * i + 1);
  }
  if (free_) {
    G.T += a * G.dT;
    upd(G.STb[0], G.ZTb[0], G.dSTb[0], a, az, mu);
    upd(G.STb[1], G.ZTb[1], G.dSTb[1], a, az, mu);
#pragma unroll
    for (int j = 0; j < 3; ++j) G.yt[j] += a * G.dyt[j];
  }
  if (w.has_term) {
#pragma unroll
    for (int j = 0; j < 3; ++j) upd(G.Stm[j], G.Ztm[j], G.dStm[j], a, az, mu);
  }
  __syncwarp();
}

}  // namespace obca

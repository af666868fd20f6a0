/* ORACLE (test infrastructure, not product code): scalar C restatement of the OBCA-MPC NLP solve.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may build, load
 * or call this file.  The product path (csrc/ *.cu) never does.
 *
 * PARITY UNPINNED: the reference hands these NLPs to CasADi Opti + IPOPT (third party, unpinned, absent
 * here: SURVEY.md 8(c)) and ships no tests or golden outputs.  This file restates the PROBLEM from
 * /root/reference/src/obca.py line by line (same citations as oracle/obca_nlp.py):
 *   variables 842-856, cost 859-897 (fixed 1385-1414), dynamics 902-911, bounds 916-923, accel 928-939,
 *   init/terminal 944/951, duals >= 0 and T bounds 956-963, obstacle rows 968-1042 (mpc4: first time block
 *   only, 969; mpc6/8/obca2: advance through the time stack, 1482/1677/538), terminal set 1465-1466.
 * The SOLVER is the primal-dual interior-point method specified by oracle/ipm_dense.py (dense NumPy),
 * here with structure-exploiting linear algebra: per (stage, obstacle) dual block eliminated through a
 * 5x5 SPD system (diagonal + low rank), then a Riccati recursion over the augmented stage state
 * (x, y, theta, v_prev, w_prev, T).  tests/test_oracle.py checks it against ipm_dense.py.
 *
 * Plain sequential loops, one instance at a time (pthreads over instances for the CPU baseline).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/obca_b200.h"

#define NS OBCA_MAX_STAGES
#define RM OBCA_MAX_ROWS
#define OM OBCA_MAX_OBS
#define FILT_MAX 32

typedef struct {
  const obca_params* P;
  int N, nobs, R, free_, has_term, stacked, init;
  int eptr[OM + 1];
  double Ts, Tmax, dmin, off, g[4];
  double x0[3], u0[2], term[3];
  const double *xref, *uref, *A, *b0, *db;
  const double* guess;   /* OBCA_INIT_GUESS: the caller's poses [N+1,3] (read from the output array x before it is written) */
  /* feasibility restoration (resto != 0): the problem solved is
   *     min  rho sum(n) + rho sum(pt + nt) + zeta/2 |D_R ((z,u,T) - (z,u,T)_R)|^2
   *     s.t. dynamics, OBCA equalities, sign rows as they are;  d_i(X) + n_i - S_i = 0, n_i >= 0 for the row classes in
   *          rmask;  z_N - r_N - pt + nt = 0, pt, nt >= 0 (free modes)
   * (IPOPT's restoration problem, Waechter & Biegler 2006 sec. 3.3, with the rows that are always consistent kept hard) */
  int resto, rmask;
  double rho, zeta;
  double mu0;        /* > 0: barrier parameter this pass starts from (instead of params.mu_init)                    */
  double th_ref;     /* restoration: violation of the original problem at the point where it was called              */
  double zR[NS][3], uR[NS][2], TR;
} prob_t;

/* row classes that the restoration phase relaxes (bits of prob_t.rmask) */
enum { CLS_XY = 1, CLS_UB = 2, CLS_TB = 4, CLS_TM = 8, CLS_NORM = 16, CLS_DIST = 32, CLS_SIGN = 0 };
#define RELAXED(p, cls) ((p)->resto && ((p)->rmask & (cls)))

/* iterate: primal X, slacks S, multipliers y (eq) and Z (ineq) */
typedef struct {
  double z[NS][3], u[NS][2], T;
  double lam[NS][RM], mu[NS][4 * OM];
  double yd[NS][3], yt[3], ye[NS][2 * OM];
  /* inequality blocks: xy (k>=1) [x-xL,y-yL,xU-x,yU-y]; ub (k<N) [u-uL(2),uU-u(2),acc+amax(2),amax-acc(2)];
   * Tb [T-Tmin,Tmax-T]; tm [xN-ts0,yN-ts1,ts2-yN]; per (k,i): lam rows, mu rows, norm, dist */
  double Sxy[NS][4], Sub[NS][8], STb[2], Stm[3], Sl[NS][RM], Sm[NS][4 * OM], Sn[NS][OM], Sd[NS][OM];
  double Zxy[NS][4], Zub[NS][8], ZTb[2], Ztm[3], Zl[NS][RM], Zm[NS][4 * OM], Zn[NS][OM], Zd[NS][OM];
  /* restoration only: relaxation n >= 0 of a row and the multiplier V of that bound; terminal equality pt, nt */
  double nxy[NS][4], nub[NS][8], nTb[2], ntm[3], nn[NS][OM], nd[NS][OM];
  double Vxy[NS][4], Vub[NS][8], VTb[2], Vtm[3], Vn[NS][OM], Vd[NS][OM];
  double pt[3], nt[3], Vpt[3], Vnt[3];
} iter_t;

/* constraint values at a point */
typedef struct {
  double f;
  double cd[NS][3], ct[3], ce[NS][2 * OM];
  double dxy[NS][4], dub[NS][8], dTb[2], dtm[3], dn[NS][OM], dd[NS][OM]; /* d for lam/mu rows = the variable */
} vals_t;

typedef struct {
  double thmax, thmin;
  int n, wr, active;
  double th[FILT_MAX], ph[FILT_MAX];
} filt_t;

static double bk(const prob_t* p, int k, int r) { return p->b0[r] + ((p->stacked && p->db) ? k * p->db[r] : 0.0); }

/* ------------------------------------------------------------------------------------------------
 * values: objective f, equality residuals c, inequality values d at (z,u,T,lam,mu)
 * ---------------------------------------------------------------------------------------------- */
static void eval_values(const prob_t* p, const iter_t* it, vals_t* v) {
  const obca_params* P = p->P;
  int N = p->N;
  double T = p->free_ ? it->T : 1.0, h = T * p->Ts;
  double f = 0.0;
  for (int k = 0; k <= N; ++k) {
    const double* z = it->z[k];
    const double* M = (k < N) ? P->Q : P->P;
    double e[3];
    for (int j = 0; j < 3; ++j) e[j] = z[j] - p->xref[3 * k + j];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) f += e[a] * M[3 * a + b] * e[b];
    if (k < N) {
      double uu[2] = {it->u[k][0], it->u[k][1]};
      if (p->uref) { uu[0] -= p->uref[2 * k]; uu[1] -= p->uref[2 * k + 1]; }
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) f += uu[a] * P->R1[2 * a + b] * uu[b];
      if (k >= 1) { /* (u_k - u_{k-1})^T R2 (..)/h^2, k = 1..N-1; the t == 0 term is identically 0 (Q4) */
        double du[2] = {it->u[k][0] - it->u[k - 1][0], it->u[k][1] - it->u[k - 1][1]};
        double s = 0;
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) s += du[a] * P->R2[2 * a + b] * du[b];
        f += s / (h * h);
      }
      double ct = cos(z[2]), st = sin(z[2]);
      v->cd[k][0] = z[0] + h * it->u[k][0] * ct - it->z[k + 1][0];
      v->cd[k][1] = z[1] + h * it->u[k][0] * st - it->z[k + 1][1];
      v->cd[k][2] = z[2] + h * it->u[k][1] - it->z[k + 1][2];
      const double* up = (k == 0) ? p->u0 : it->u[k - 1];
      for (int j = 0; j < 2; ++j) {
        double ga = (up[j] - it->u[k][j]) / h;
        v->dub[k][j] = it->u[k][j] - P->uL[j];
        v->dub[k][2 + j] = P->uU[j] - it->u[k][j];
        v->dub[k][4 + j] = ga + P->acc_max[j];
        v->dub[k][6 + j] = P->acc_max[j] - ga;
      }
    }
    if (k >= 1)
      for (int j = 0; j < 2; ++j) {
        v->dxy[k][j] = z[j] - P->xL[j];
        v->dxy[k][2 + j] = P->xU[j] - z[j];
      }
    double ct = cos(z[2]), st = sin(z[2]);
    double tx = z[0] + p->off * ct, ty = z[1] + p->off * st;
    for (int i = 0; i < p->nobs; ++i) {
      double a1 = 0, a2 = 0, bl = 0;
      for (int r = p->eptr[i]; r < p->eptr[i + 1]; ++r) {
        a1 += p->A[2 * r] * it->lam[k][r];
        a2 += p->A[2 * r + 1] * it->lam[k][r];
        bl += bk(p, k, r) * it->lam[k][r];
      }
      const double* m = &it->mu[k][4 * i];
      v->ce[k][2 * i] = m[0] - m[2] + ct * a1 + st * a2;
      v->ce[k][2 * i + 1] = m[1] - m[3] - st * a1 + ct * a2;
      v->dn[k][i] = 1.0 - a1 * a1 - a2 * a2;
      v->dd[k][i] = -(p->g[0] * m[0] + p->g[1] * m[1] + p->g[2] * m[2] + p->g[3] * m[3]) + tx * a1 + ty * a2 - bl - p->dmin;
    }
  }
  if (p->free_) {
    f += (N + 1) * (P->time_cost[0] * T + P->time_cost[1] * T * T);
    for (int j = 0; j < 3; ++j) v->ct[j] = it->z[N][j] - p->xref[3 * N + j];
    v->dTb[0] = T - P->T_min;
    v->dTb[1] = p->Tmax - T;
  }
  if (p->has_term) {
    v->dtm[0] = it->z[N][0] - p->term[0];
    v->dtm[1] = it->z[N][1] - p->term[1];
    v->dtm[2] = p->term[2] - it->z[N][1];
  }
  v->f = f;
}

/* visit every inequality row: value d_, slack S_, multiplier Z_ and - in the restoration phase, for the relaxed row
 * classes - the relaxation variable n_ and its bound multiplier V_ (NULL otherwise) */
#define ROW_(cls, dv, Sp, Zp, np_, Vp, ...)                                                        \
  { double d_ = (dv); double* S_ = (Sp); double* Z_ = (Zp);                                          \
    double* n_ = RELAXED(p_, cls) ? (np_) : 0; double* V_ = n_ ? (Vp) : 0; (void)d_; (void)S_; (void)Z_; (void)n_; (void)V_; __VA_ARGS__ }
#define FOR_INEQ(p, it, v, ...)                                                                     \
  do {                                                                                             \
    const prob_t* p_ = (p);                                                                        \
    int N_ = p_->N;                                                                                \
    for (int k = 0; k <= N_; ++k) {                                                                \
      if (k >= 1) for (int j = 0; j < 4; ++j) ROW_(CLS_XY, (v)->dxy[k][j], &(it)->Sxy[k][j], &(it)->Zxy[k][j], &(it)->nxy[k][j], &(it)->Vxy[k][j], __VA_ARGS__) \
      if (k < N_) for (int j = 0; j < 8; ++j) ROW_(CLS_UB, (v)->dub[k][j], &(it)->Sub[k][j], &(it)->Zub[k][j], &(it)->nub[k][j], &(it)->Vub[k][j], __VA_ARGS__) \
      for (int r = 0; r < p_->R; ++r) ROW_(CLS_SIGN, (it)->lam[k][r], &(it)->Sl[k][r], &(it)->Zl[k][r], (double*)0, (double*)0, __VA_ARGS__) \
      for (int r = 0; r < 4 * p_->nobs; ++r) ROW_(CLS_SIGN, (it)->mu[k][r], &(it)->Sm[k][r], &(it)->Zm[k][r], (double*)0, (double*)0, __VA_ARGS__) \
      for (int i = 0; i < p_->nobs; ++i) ROW_(CLS_NORM, (v)->dn[k][i], &(it)->Sn[k][i], &(it)->Zn[k][i], &(it)->nn[k][i], &(it)->Vn[k][i], __VA_ARGS__) \
      for (int i = 0; i < p_->nobs; ++i) ROW_(CLS_DIST, (v)->dd[k][i], &(it)->Sd[k][i], &(it)->Zd[k][i], &(it)->nd[k][i], &(it)->Vd[k][i], __VA_ARGS__) \
    }                                                                                              \
    if (p_->free_) for (int j = 0; j < 2; ++j) ROW_(CLS_TB, (v)->dTb[j], &(it)->STb[j], &(it)->ZTb[j], &(it)->nTb[j], &(it)->VTb[j], __VA_ARGS__) \
    if (p_->has_term) for (int j = 0; j < 3; ++j) ROW_(CLS_TM, (v)->dtm[j], &(it)->Stm[j], &(it)->Ztm[j], &(it)->ntm[j], &(it)->Vtm[j], __VA_ARGS__) \
  } while (0)

/* restoration objective: rho (sum n + sum pt + sum nt) + zeta/2 |D_R ((z,u,T) - reference)|^2, D_R = min(1, 1/|reference|) */
static double dr2(double ref) { double a = fabs(ref); return a > 1.0 ? 1.0 / (a * a) : 1.0; }
static double resto_objective(const prob_t* p, const iter_t* it, const vals_t* v) {
  double f = 0, sn = 0;
  int N = p->N;
  for (int k = 0; k <= N; ++k) {
    if (k >= 1) for (int j = 0; j < 3; ++j) { double e = it->z[k][j] - p->zR[k][j]; f += dr2(p->zR[k][j]) * e * e; }
    if (k < N) for (int j = 0; j < 2; ++j) { double e = it->u[k][j] - p->uR[k][j]; f += dr2(p->uR[k][j]) * e * e; }
  }
  if (p->free_) { double e = it->T - p->TR; f += dr2(p->TR) * e * e; }
  iter_t* itm = (iter_t*)it;
  FOR_INEQ(p, itm, v, { if (n_) sn += *n_; });
  if (p->free_) for (int j = 0; j < 3; ++j) sn += it->pt[j] + it->nt[j];
  return 0.5 * p->zeta * f + p->rho * sn;
}

/* th: constraint violation (1-norm) of the problem being solved, ph: its barrier function, cmax: max-norm violation,
 * th_orig: violation of the ORIGINAL problem at (X, S) (what the restoration phase is there to reduce) */
static void theta_phi(const prob_t* p, const iter_t* it, const vals_t* v, double mu, double* th, double* ph,
                      double* cmax, double* th_orig) {
  double t = 0, lg = 0, cm = 0, to = 0;
  int N = p->N;
  for (int k = 0; k <= N; ++k) {
    if (k < N) for (int j = 0; j < 3; ++j) { t += fabs(v->cd[k][j]); cm = fmax(cm, fabs(v->cd[k][j])); }
    for (int j = 0; j < 2 * p->nobs; ++j) { t += fabs(v->ce[k][j]); cm = fmax(cm, fabs(v->ce[k][j])); }
  }
  to = t;
  if (p->free_) for (int j = 0; j < 3; ++j) {
    double c = v->ct[j];
    to += fabs(c);
    if (p->resto) { c += -it->pt[j] + it->nt[j]; lg += log(it->pt[j]) + log(it->nt[j]); }
    t += fabs(c); cm = fmax(cm, fabs(c));
  }
  iter_t* itm = (iter_t*)it;
  FOR_INEQ(p, itm, v, {
    double r_ = d_ - *S_;
    to += fabs(r_);
    if (n_) { r_ += *n_; lg += log(*n_); }
    t += fabs(r_); cm = fmax(cm, fabs(r_)); lg += log(*S_); });
  *th = t;
  *ph = (p->resto ? resto_objective(p, it, v) : v->f) - mu * lg;
  if (cmax) *cmax = cm;
  if (th_orig) *th_orig = to;
}

/* ------------------------------------------------------------------------------------------------
 * start point (oracle/obca_nlp.py start_point)
 * ---------------------------------------------------------------------------------------------- */
static void start_point(const prob_t* p, iter_t* it) {
  const obca_params* P = p->P;
  int N = p->N, init = p->init;
  if (init == OBCA_INIT_KEEP) { /* soft restart: primal point stays, equality multipliers start at 0 */
    memset(it->yd, 0, sizeof(it->yd)); memset(it->yt, 0, sizeof(it->yt)); memset(it->ye, 0, sizeof(it->ye));
    return;
  }
  memset(it, 0, sizeof(*it));
  for (int j = 0; j < 3; ++j) it->z[0][j] = p->x0[j];
  it->T = 1.0;
  const double* src = (init == OBCA_INIT_GUESS) ? p->guess : p->xref;   /* poses the start point is built from */
  if (init >= OBCA_INIT_XREF)
    for (int k = 1; k <= N; ++k)
      for (int j = 0; j < 3; ++j) it->z[k][j] = src[3 * k + j];
  if (init == OBCA_INIT_WARM || init == OBCA_INIT_GUESS) {
    double Pp[NS][3];
    for (int j = 0; j < 3; ++j) Pp[0][j] = p->x0[j];
    for (int k = 1; k <= N; ++k)
      for (int j = 0; j < 3; ++j) Pp[k][j] = src[3 * k + j];
    double len = 0;
    for (int k = 0; k < N; ++k) len += sqrt(pow(Pp[k + 1][0] - Pp[k][0], 2) + pow(Pp[k + 1][1] - Pp[k][1], 2));
    double h = p->Ts;
    if (p->free_) {
      double T0 = len / (N * P->uU[0] * p->Ts);
      T0 = fmin(fmax(T0, 1.0), fmax(p->Tmax, P->T_min));
      it->T = T0;
      h = T0 * p->Ts;
    }
    for (int k = 0; k < N; ++k) {
      double dth = Pp[k + 1][2] - Pp[k][2];
      dth = dth + M_PI;
      dth = dth - 2 * M_PI * floor(dth / (2 * M_PI)) - M_PI;
      double fwd = cos(Pp[k][2]) * (Pp[k + 1][0] - Pp[k][0]) + sin(Pp[k][2]) * (Pp[k + 1][1] - Pp[k][1]);
      it->u[k][0] = fmin(fmax(fwd / h, P->uL[0]), P->uU[0]);
      it->u[k][1] = fmin(fmax(dth / h, P->uL[1]), P->uU[1]);
    }
    for (int k = 0; k <= N; ++k) {
      double ct = cos(Pp[k][2]), st = sin(Pp[k][2]);
      double tx = Pp[k][0] + p->off * ct, ty = Pp[k][1] + p->off * st;
      for (int i = 0; i < p->nobs; ++i) {
        int jb = -1;
        double best = -1e300, nb = 1;
        for (int r = p->eptr[i]; r < p->eptr[i + 1]; ++r) {
          double nr = sqrt(p->A[2 * r] * p->A[2 * r] + p->A[2 * r + 1] * p->A[2 * r + 1]);
          double sep = (p->A[2 * r] * tx + p->A[2 * r + 1] * ty - bk(p, k, r)) / nr;
          if (sep > best) { best = sep; jb = r; nb = nr; }
        }
        if (jb < 0) continue;
        double l = 0.9 / nb;
        it->lam[k][jb] = l;
        double a1 = p->A[2 * jb] * l, a2 = p->A[2 * jb + 1] * l;
        double r1 = -(ct * a1 + st * a2), r2 = -(-st * a1 + ct * a2);
        double* m = &it->mu[k][4 * i];
        m[0] = fmax(r1, 0); m[1] = fmax(r2, 0); m[2] = fmax(-r1, 0); m[3] = fmax(-r2, 0);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * per (stage, obstacle) dual block, solved in STEP form (all unknowns are Newton steps, every right-hand
 * side is a residual, so rounding errors scale with the step and vanish at convergence).
 * Unknowns: dw (rows j = lambda rows then the 4 mu rows; each row has a sign constraint with slack, so its
 * curvature is sigma_j = Z_j/S_j) and dzeta = (tau_a (2), delta_d, dye (2)):
 *     tau_a   = Cn (alpha^T dw)        Cn = 2 Zn I + 4 sn a a^T   (Hessian + Sigma term of the norm row)
 *     delta_d = sd (grad dist . dv) - (td - Zd)  = -(step of the dist-row multiplier), kept in augmented form
 *     dye     = step of the multipliers of the two equalities
 * Block KKT with Y = [alpha0 alpha1 q je1 je2] (one 5-vector per row), C = blockdiag(Cn^-1, 1/sd, 0, 0):
 *     sigma_j dw_j + y_j . dzeta = t'_j + y_j . h       (h: right-hand sides living in span(alpha0, alpha1))
 *     Y^T dw - C dzeta           = -kappa               (kappa: coupling to the pose / constraint residuals)
 * With eta = dzeta - h:   M eta = g0 + kappa - C h,   M = Y^T D^-1 Y + C,   dw_j = (t'_j - y_j . eta)/sigma_j.
 * M = R^T R is factorised in square-root form (Givens row insertion), never as a Gram matrix: sigma_j
 * spans 1e-9 .. 1e+16 near convergence and cond(M) must not be squared.
 * ---------------------------------------------------------------------------------------------- */
#define SIG_MIN 1e-8 /* primal regularisation of the OBCA duals: curvature of a sign row is max(Z/S, SIG_MIN) */

/* sigma = Z/S and the step-form right-hand side  t - Z = (mu - S Z)/S - sigma (d - S)  of one inequality */
static void sig_t(double S, double Z, double d, double mu, double* sig, double* t) {
  *sig = Z / S;
  *t = (mu - S * Z) / S - (*sig) * (d - S);
}
/* the same for a row relaxed by the restoration phase:  d + n - S = 0, n >= 0 with cost rho n and bound multiplier V
 * (stationarity in n: rho - Z - V = 0).  Eliminating dS, dn, dV from the Newton equations leaves
 *     dZ = t - sigma (grad d . dX),   sigma = 1 / (S/Z + n/V),   t = -sigma [(d + n - S) + (mu - n (rho - Z))/V - (mu - S Z)/Z]
 * which tends to the ordinary row as n -> 0 */
static void sig_t_relaxed(double S, double Z, double n, double V, double d, double mu, double rho, double* sig, double* t) {
  *sig = 1.0 / (S / Z + n / V);
  *t = -(*sig) * ((d + n - S) + (mu - n * (rho - Z)) / V - (mu - S * Z) / Z);
}
/* steps of slack and relaxation of such a row from the step dZ of its multiplier */
static void relaxed_steps(double S, double Z, double n, double V, double mu, double rho, double dZ, double* dS, double* dn) {
  *dS = (mu - S * Z - S * dZ) / Z;
  double dV = (rho - Z - V) - dZ;
  *dn = (mu - n * V - n * dV) / V;
}

typedef struct {
  double L[5][5];                  /* R^T */
  double g0[5];                    /* Y^T D^-1 t' */
  double a1, a2, ct, st, tx, ty;
  double Ci[2];                    /* Cn^-1 v = (v - Ci1 a (a.v)) Ci0 */
  double h0[2], hc[3][2], jt[2];   /* rhs / pose columns (x,y,theta) in the (alpha0, alpha1) basis; Je_p[:,theta] */
  double eta[4][5];                /* M^-1 b for columns (0, x, y, theta) */
  double dpose[3];                 /* grad of dist wrt pose */
} blk_t;

static void row_y(const prob_t* p, const blk_t* b, int k, int i, int j, int E, double yv[5]) {
  if (j < E) {
    int r = p->eptr[i] + j;
    double A0 = p->A[2 * r], A1 = p->A[2 * r + 1];
    yv[0] = A0; yv[1] = A1;
    yv[2] = b->tx * A0 + b->ty * A1 - bk(p, k, r);
    yv[3] = b->ct * A0 + b->st * A1;
    yv[4] = -b->st * A0 + b->ct * A1;
  } else {
    int m = j - E;
    yv[0] = 0; yv[1] = 0; yv[2] = -p->g[m];
    yv[3] = (m == 0) ? 1.0 : (m == 2) ? -1.0 : 0.0;
    yv[4] = (m == 1) ? 1.0 : (m == 3) ? -1.0 : 0.0;
  }
}

static void tri5_solve(const double L[5][5], const double* r, double* x) { /* (L L^T) x = r */
  double t[5];
  for (int i = 0; i < 5; ++i) {
    double s = r[i];
    for (int q = 0; q < i; ++q) s -= L[i][q] * t[q];
    t[i] = s / L[i][i];
  }
  for (int i = 4; i >= 0; --i) {
    double s = t[i];
    for (int q = i + 1; q < 5; ++q) s -= L[q][i] * x[q];
    x[i] = s / L[i][i];
  }
}

static void qr5_insert(double L[5][5], double row[5]) {
  for (int c = 0; c < 5; ++c) {
    double a = L[c][c], b = row[c];
    if (b == 0.0) continue;
    double r = sqrt(a * a + b * b), cs = a / r, sn = b / r;
    L[c][c] = r;
    for (int q = c + 1; q < 5; ++q) {
      double u = L[q][c], w = row[q];
      L[q][c] = cs * u + sn * w;
      row[q] = -sn * u + cs * w;
    }
  }
}

/* step-form right-hand side of row j:  (t_j - Z_j) - (d L / d w_j)  with the CURRENT multipliers */
static double row_rhs(const iter_t* it, int k, int i, const blk_t* b, double mu, double w, double S, double Z,
                      const double yv[5]) {
  double sig = Z / S;
  double gl = it->ye[k][2 * i] * yv[3] + it->ye[k][2 * i + 1] * yv[4] - Z + 2 * it->Zn[k][i] * (b->a1 * yv[0] + b->a2 * yv[1]) -
              it->Zd[k][i] * yv[2];
  return (mu - S * Z) / S - sig * (w - S) - gl;
}

static void row_szw(const iter_t* it, int k, int i, int r0, int E, int j, double* w, double* S, double* Z) {
  if (j < E) { *w = it->lam[k][r0 + j]; *S = it->Sl[k][r0 + j]; *Z = it->Zl[k][r0 + j]; }
  else { *w = it->mu[k][4 * i + j - E]; *S = it->Sm[k][4 * i + j - E]; *Z = it->Zm[k][4 * i + j - E]; }
}

static void cn_inv(const blk_t* b, const double v[2], double o[2]) {
  double av = b->a1 * v[0] + b->a2 * v[1];
  o[0] = (v[0] - b->Ci[1] * b->a1 * av) * b->Ci[0];
  o[1] = (v[1] - b->Ci[1] * b->a2 * av) * b->Ci[0];
}

/* norm and dist rows of block (k, i): sigma and step-form right-hand side (dist: divided by sigma) */
static void rows_nd(const prob_t* p, const iter_t* it, const vals_t* v, double mu, int k, int i, double* sn, double* tn,
                    double* sd, double* tds) {
  double Sn = it->Sn[k][i], Zn = it->Zn[k][i], Sd = it->Sd[k][i], Zd = it->Zd[k][i];
  if (RELAXED(p, CLS_NORM))
    sig_t_relaxed(Sn, Zn, it->nn[k][i], it->Vn[k][i], v->dn[k][i], mu, p->rho, sn, tn);
  else {
    *sn = Zn / Sn;
    *tn = (mu - Sn * Zn) / Sn - (*sn) * (v->dn[k][i] - Sn); /* step form: t - Z */
  }
  if (RELAXED(p, CLS_DIST)) {
    double t;
    sig_t_relaxed(Sd, Zd, it->nd[k][i], it->Vd[k][i], v->dd[k][i], mu, p->rho, sd, &t);
    *tds = t / (*sd);
  } else {
    *sd = Zd / Sd;
    *tds = (mu - Sd * Zd) / Zd - (v->dd[k][i] - Sd);        /* (td - Zd) / sd */
  }
}

/* builds the block factorisation; if Hp/rp != NULL adds this block's Schur complement to the pose
 * Hessian (3x3) and reduced gradient (3) */
static int block_setup(const prob_t* p, const iter_t* it, const vals_t* v, double mu, int k, int i, blk_t* b,
                       double Hp[3][3], double rp[3]) {
  int E = p->eptr[i + 1] - p->eptr[i], r0 = p->eptr[i];
  const double* z = it->z[k];
  b->ct = cos(z[2]); b->st = sin(z[2]);
  b->tx = z[0] + p->off * b->ct; b->ty = z[1] + p->off * b->st;
  double a1 = 0, a2 = 0;
  for (int r = r0; r < r0 + E; ++r) { a1 += p->A[2 * r] * it->lam[k][r]; a2 += p->A[2 * r + 1] * it->lam[k][r]; }
  b->a1 = a1; b->a2 = a2;
  memset(b->g0, 0, sizeof(b->g0)); memset(b->L, 0, sizeof(b->L));
  for (int j = 0; j < E + 4; ++j) {
    double w, S, Z, yv[5], yh[5];
    row_szw(it, k, i, r0, E, j, &w, &S, &Z);
    row_y(p, b, k, i, j, E, yv);
    double sig = Z / S, t = row_rhs(it, k, i, b, mu, w, S, Z, yv);
    double di = 1.0 / fmax(sig, SIG_MIN), sq = sqrt(di);
    for (int a = 0; a < 5; ++a) { b->g0[a] += yv[a] * t * di; yh[a] = yv[a] * sq; }
    qr5_insert(b->L, yh);
  }
  double Zn = it->Zn[k][i], Zd = it->Zd[k][i];
  double sn, tn, sd, tds;
  rows_nd(p, it, v, mu, k, i, &sn, &tn, &sd, &tds);
  {
    /* Cn = 2 Zn I + 4 sn a a^T has eigenpairs (2 Zn + 4 sn |a|^2, a/|a|) and (2 Zn, a_perp); when the norm
     * row is active sn ~ 1e10 and Cn^-1 is numerically singular, so it is only ever used in this spectral
     * form:  Cn^-1 v = (v - kn a (a.v)) / (2 Zn),  kn = 4 sn / (2 Zn + 4 sn |a|^2). */
    double aa = a1 * a1 + a2 * a2, lam1 = 2 * Zn + 4 * sn * aa;
    b->Ci[0] = 1.0 / (2 * Zn); b->Ci[1] = 4 * sn / lam1;
    double r1[5] = {0, 0, 0, 0, 0}, r2[5] = {0, 0, 0, 0, 0}, r3[5] = {0, 0, sqrt(1.0 / sd), 0, 0};
    if (aa > 0) {
      double na = sqrt(aa), e1 = a1 / na, e2 = a2 / na, s1 = sqrt(1.0 / lam1), s2 = sqrt(b->Ci[0]);
      r1[0] = s1 * e1; r1[1] = s1 * e2; r2[0] = -s2 * e2; r2[1] = s2 * e1;
    } else {
      r1[0] = sqrt(b->Ci[0]); r2[1] = r1[0];
    }
    qr5_insert(b->L, r1); qr5_insert(b->L, r2); qr5_insert(b->L, r3);
    for (int a = 0; a < 5; ++a) {
      if (b->L[a][a] < 0) for (int q = a; q < 5; ++q) b->L[q][a] = -b->L[q][a];
      if (!(b->L[a][a] > 0) || !isfinite(b->L[a][a])) return -1;
    }
  }
  /* column 0: eta0 = M^-1 (g0 + kappa0 - C h0),  kappa0 = (0, 0, -(td - Zd)/sd, e1, e2),  h0 = -2 (tn - Zn) a */
  b->h0[0] = -2 * tn * a1; b->h0[1] = -2 * tn * a2;
  double rhs[5], ch[2];
  cn_inv(b, b->h0, ch);
  rhs[0] = b->g0[0] - ch[0];
  rhs[1] = b->g0[1] - ch[1];
  rhs[2] = b->g0[2] - tds;
  rhs[3] = b->g0[3] + v->ce[k][2 * i];
  rhs[4] = b->g0[4] + v->ce[k][2 * i + 1];
  tri5_solve(b->L, rhs, b->eta[0]);
  double offt = p->off * (-b->st * a1 + b->ct * a2);
  b->dpose[0] = a1; b->dpose[1] = a2; b->dpose[2] = offt;
  if (k == 0) return 0; /* pose fixed */
  double y1 = it->ye[k][2 * i], y2 = it->ye[k][2 * i + 1], yd = -Zd;
  double c1 = y1 + yd * p->off;
  /* -H_wp columns are -U h^c (W cross terms only); kappa^c = (0, 0, dpose_c, Je_p[:,c]) */
  b->hc[0][0] = yd; b->hc[0][1] = 0;
  b->hc[1][0] = 0; b->hc[1][1] = yd;
  b->hc[2][0] = -c1 * b->st - y2 * b->ct; b->hc[2][1] = c1 * b->ct - y2 * b->st;
  b->jt[0] = -b->st * a1 + b->ct * a2; b->jt[1] = -b->ct * a1 - b->st * a2;
  double bc[3][5];
  for (int c = 0; c < 3; ++c) {
    cn_inv(b, b->hc[c], bc[c]);
    bc[c][2] = b->dpose[c];
    bc[c][3] = (c == 2) ? b->jt[0] : 0.0;
    bc[c][4] = (c == 2) ? b->jt[1] : 0.0;
    tri5_solve(b->L, bc[c], b->eta[1 + c]); /* dzeta^c = eta^c - h^c */
  }
  if (Hp) {
    /* own pose term of this block: W (theta,theta) */
    Hp[2][2] += y1 * (-b->ct * a1 - b->st * a2) + y2 * (b->st * a1 - b->ct * a2) + yd * p->off * (-b->ct * a1 - b->st * a2);
    /* Schur complement  Gamma(c',c) = h^c'^T Cn^-1 h^c - b^c'^T M^-1 b^c ;  Gamma(c',0) = -b^c' . eta0 - (Cn^-1 h^c') . h0 */
    for (int cp = 0; cp < 3; ++cp) {
      for (int c = 0; c < 3; ++c) {
        double G = b->hc[cp][0] * bc[c][0] + b->hc[cp][1] * bc[c][1];
        for (int a = 0; a < 5; ++a) G -= bc[cp][a] * b->eta[1 + c][a];
        Hp[cp][c] -= G;
      }
      double G0 = -(bc[cp][0] * b->h0[0] + bc[cp][1] * b->h0[1]);
      for (int a = 0; a < 5; ++a) G0 -= bc[cp][a] * b->eta[0][a];
      rp[cp] += G0;
    }
  }
  return 0;
}

/* search direction */
typedef struct {
  double z[NS][3], u[NS][2], T;
  double lam[NS][RM], mu[NS][4 * OM];
  double yd[NS][3], yt[3], ye[NS][2 * OM]; /* NEW multipliers y + dy (the step is this minus y) */
  double Sxy[NS][4], Sub[NS][8], STb[2], Stm[3], Sl[NS][RM], Sm[NS][4 * OM], Sn[NS][OM], Sd[NS][OM];
  double nxy[NS][4], nub[NS][8], nTb[2], ntm[3], nn[NS][OM], nd[NS][OM], pt[3], nt[3]; /* restoration only */
} dir_t;

static void block_backsub(const prob_t* p, const iter_t* it, const vals_t* v, double mu, int k, int i, const blk_t* b,
                          const double dp[3], dir_t* d) {
  int E = p->eptr[i + 1] - p->eptr[i], r0 = p->eptr[i];
  double et[5], ht[2] = {b->h0[0], b->h0[1]};
  for (int a = 0; a < 5; ++a) et[a] = b->eta[0][a];
  if (k >= 1)
    for (int c = 0; c < 3; ++c) {
      for (int a = 0; a < 5; ++a) et[a] += b->eta[1 + c][a] * dp[c];
      ht[0] -= b->hc[c][0] * dp[c]; ht[1] -= b->hc[c][1] * dp[c];
    }
  double da1 = 0, da2 = 0, qdw = 0;
  for (int j = 0; j < E + 4; ++j) {
    double w, S, Z, yv[5];
    row_szw(it, k, i, r0, E, j, &w, &S, &Z);
    row_y(p, b, k, i, j, E, yv);
    double sig = Z / S, s = row_rhs(it, k, i, b, mu, w, S, Z, yv);
    for (int a = 0; a < 5; ++a) s -= yv[a] * et[a];
    double dw = s / fmax(sig, SIG_MIN);
    if (j < E) {
      d->lam[k][r0 + j] = dw; d->Sl[k][r0 + j] = dw + (w - S);
      da1 += yv[0] * dw; da2 += yv[1] * dw;
    } else {
      d->mu[k][4 * i + j - E] = dw; d->Sm[k][4 * i + j - E] = dw + (w - S);
    }
    qdw += yv[2] * dw;
  }
  d->ye[k][2 * i] = it->ye[k][2 * i] + et[3]; d->ye[k][2 * i + 1] = it->ye[k][2 * i + 1] + et[4];
  /* Rows kept in augmented form (norm, dist): when the row is active (sigma = Z/S >= 1) the step of its
   * multiplier comes from the solve and the slack step from the linearised complementarity, so that the
   * stationarity rows stay consistent without multiplying a rounding error by sigma; when inactive the
   * slack step comes from the primal direction.  (dZ = mu/S - Z - sigma dS is applied by the caller.) */
  double Sn = it->Sn[k][i], Zn = it->Zn[k][i], Sd = it->Sd[k][i], Zd = it->Zd[k][i];
  double sn, tn, sd, tds;
  rows_nd(p, it, v, mu, k, i, &sn, &tn, &sd, &tds);
  double ada = b->a1 * da1 + b->a2 * da2;
  if (sn >= 1.0) /* tau_a = eta_a + h_a = Cn (alpha^T dw)  =>  a . (alpha^T dw) = a . tau_a / (2 Zn + 4 sn |a|^2) */
    ada = (b->a1 * (et[0] + ht[0]) + b->a2 * (et[1] + ht[1])) / (2 * Zn + 4 * sn * (b->a1 * b->a1 + b->a2 * b->a2));
  if (RELAXED(p, CLS_NORM))
    relaxed_steps(Sn, Zn, it->nn[k][i], it->Vn[k][i], mu, p->rho, tn - sn * (-2 * ada), &d->Sn[k][i], &d->nn[k][i]);
  else
    d->Sn[k][i] = -2 * ada + (v->dn[k][i] - Sn);
  if (RELAXED(p, CLS_DIST)) {
    double gd = qdw;                                  /* grad dist . (dw, dpose) */
    if (k >= 1) for (int c = 0; c < 3; ++c) gd += b->dpose[c] * dp[c];
    double dZ = (sd >= 1.0) ? -et[2] : sd * (tds - gd);
    relaxed_steps(Sd, Zd, it->nd[k][i], it->Vd[k][i], mu, p->rho, dZ, &d->Sd[k][i], &d->nd[k][i]);
  } else if (sd >= 1.0) {
    double dZ = -et[2];
    d->Sd[k][i] = (mu - Sd * Zd - Sd * dZ) / Zd;
  } else {
    double s = qdw + (v->dd[k][i] - Sd);
    if (k >= 1) for (int c = 0; c < 3; ++c) s += b->dpose[c] * dp[c];
    d->Sd[k][i] = s;
  }
}

/* ------------------------------------------------------------------------------------------------
 * stage QP data over (xi, u) = (x, y, th, v_prev, w_prev, T | v, w):  H (8x8), reduced gradient r
 * (without J^T y: the linear solve returns the new multipliers), gradient of the Lagrangian gL (for the
 * optimality error) and objective gradient gf (for the directional derivative of the barrier function)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double H[NS][8][8], r[NS][8], gL[NS][8], gf[NS][8];
  double A[NS][3][2];  /* hF_z theta column (2) and Ts*F (3) are kept explicit below */
  double Fth[NS][2], FT[NS][3], Bv[NS][2], Bw[NS]; /* d z+/d theta (x,y), d z+/d T, d z+/d v (x,y), d th+/d w */
} stageqp_t;

#define SIGT(cls, S_, Z_, n_, V_, d_, sg_, t_)                                                          \
  do { if (RELAXED(p, cls)) sig_t_relaxed(S_, Z_, n_, V_, d_, mu, p->rho, sg_, t_); else sig_t(S_, Z_, d_, mu, sg_, t_); } while (0)
static int assemble(const prob_t* p, const iter_t* it, const vals_t* v, double mu, stageqp_t* q, blk_t (*blk)[OM]) {
  const obca_params* P = p->P;
  int N = p->N;
  double T = p->free_ ? it->T : 1.0, h = T * p->Ts;
  memset(q->H, 0, sizeof(q->H)); memset(q->r, 0, sizeof(q->r)); memset(q->gL, 0, sizeof(q->gL));
  memset(q->gf, 0, sizeof(q->gf));
  for (int k = 0; k <= N; ++k) {
    double (*H)[8] = q->H[k];
    double* r = q->r[k];
    double* gL = q->gL[k];
    double* gf = q->gf[k];
    const double* z = it->z[k];
    double ct = cos(z[2]), st = sin(z[2]);
    /* (1) tracking; restoration: proximity to the reference point instead of the objective */
    const double* M = (k < N) ? P->Q : P->P;
    double e[3];
    for (int j = 0; j < 3; ++j) e[j] = z[j] - p->xref[3 * k + j];
    if (p->resto) {
      if (k >= 1) for (int a = 0; a < 3; ++a) { double w_ = p->zeta * dr2(p->zR[k][a]); gf[a] += w_ * (z[a] - p->zR[k][a]); H[a][a] += w_; }
    } else
    for (int a = 0; a < 3; ++a) {
      double s = 0;
      for (int b = 0; b < 3; ++b) {
        s += (M[3 * a + b] + M[3 * b + a]) * e[b];
        H[a][b] += M[3 * a + b] + M[3 * b + a];
      }
      gf[a] += s;
    }
    if (k < N) {
      const double* u = it->u[k];
      /* (2) input cost */
      double uu[2] = {u[0], u[1]};
      if (p->uref) { uu[0] -= p->uref[2 * k]; uu[1] -= p->uref[2 * k + 1]; }
      if (p->resto) {
        for (int a = 0; a < 2; ++a) { double w_ = p->zeta * dr2(p->uR[k][a]); gf[6 + a] += w_ * (u[a] - p->uR[k][a]); H[6 + a][6 + a] += w_; }
      } else
      for (int a = 0; a < 2; ++a) {
        double s = 0;
        for (int b = 0; b < 2; ++b) {
          s += (P->R1[2 * a + b] + P->R1[2 * b + a]) * uu[b];
          H[6 + a][6 + b] += P->R1[2 * a + b] + P->R1[2 * b + a];
        }
        gf[6 + a] += s;
      }
      /* (3) acceleration cost between u_{k-1} (state 3,4) and u_k, k >= 1 */
      if (k >= 1 && !p->resto) {
        double du[2] = {u[0] - it->u[k - 1][0], u[1] - it->u[k - 1][1]}, qv[2], Aacc = 0;
        for (int a = 0; a < 2; ++a) {
          qv[a] = 0;
          for (int b = 0; b < 2; ++b) qv[a] += 0.5 * (P->R2[2 * a + b] + P->R2[2 * b + a]) * du[b];
          Aacc += du[a] * qv[a];
        }
        Aacc /= h * h;
        for (int a = 0; a < 2; ++a) {
          gf[6 + a] += 2 * qv[a] / (h * h);
          gf[3 + a] -= 2 * qv[a] / (h * h);
          for (int b = 0; b < 2; ++b) {
            double m = (P->R2[2 * a + b] + P->R2[2 * b + a]) / (h * h);
            H[6 + a][6 + b] += m; H[3 + a][3 + b] += m; H[6 + a][3 + b] -= m; H[3 + a][6 + b] -= m;
          }
          if (p->free_) {
            double c = 4 * qv[a] / (h * h * T);
            H[5][6 + a] -= c; H[6 + a][5] -= c; H[5][3 + a] += c; H[3 + a][5] += c;
          }
        }
        if (p->free_) { gf[5] -= 2 * Aacc / T; H[5][5] += 6 * Aacc / (T * T); }
      }
      /* (5) dynamics: linearisation and Hessian-of-Lagrangian terms */
      double vv = u[0], ww = u[1];
      q->Fth[k][0] = -h * vv * st; q->Fth[k][1] = h * vv * ct;
      q->Bv[k][0] = h * ct; q->Bv[k][1] = h * st; q->Bw[k] = h;
      q->FT[k][0] = p->free_ ? p->Ts * vv * ct : 0; q->FT[k][1] = p->free_ ? p->Ts * vv * st : 0;
      q->FT[k][2] = p->free_ ? p->Ts * ww : 0;
      const double* y = it->yd[k];
      H[2][2] += h * vv * (-y[0] * ct - y[1] * st);
      H[2][6] += h * (-y[0] * st + y[1] * ct); H[6][2] = H[2][6];
      if (p->free_) {
        double a = p->Ts * vv * (-y[0] * st + y[1] * ct);
        H[5][2] += a; H[2][5] += a;
        a = p->Ts * (y[0] * ct + y[1] * st);
        H[5][6] += a; H[6][5] += a;
        a = p->Ts * y[2];
        H[5][7] += a; H[7][5] += a;
      }
      /* J^T y for the Lagrangian gradient: c_k depends on z_k, u_k, T */
      gL[0] += y[0]; gL[1] += y[1]; gL[2] += y[2] + q->Fth[k][0] * y[0] + q->Fth[k][1] * y[1];
      gL[6] += q->Bv[k][0] * y[0] + q->Bv[k][1] * y[1]; gL[7] += h * y[2];
      gL[5] += q->FT[k][0] * y[0] + q->FT[k][1] * y[1] + q->FT[k][2] * y[2];
      /* (7) input bounds and acceleration rows */
      const double* up = (k == 0) ? p->u0 : it->u[k - 1];
      for (int j = 0; j < 2; ++j) {
        double sg, t;
        SIGT(CLS_UB, it->Sub[k][j], it->Zub[k][j], it->nub[k][j], it->Vub[k][j], v->dub[k][j], &sg, &t);
        H[6 + j][6 + j] += sg; r[6 + j] += t; gL[6 + j] -= it->Zub[k][j];
        SIGT(CLS_UB, it->Sub[k][2 + j], it->Zub[k][2 + j], it->nub[k][2 + j], it->Vub[k][2 + j], v->dub[k][2 + j], &sg, &t);
        H[6 + j][6 + j] += sg; r[6 + j] -= t; gL[6 + j] += it->Zub[k][2 + j];
        double ga = (up[j] - u[j]) / h;
        double s4, t4, s6, t6;
        SIGT(CLS_UB, it->Sub[k][4 + j], it->Zub[k][4 + j], it->nub[k][4 + j], it->Vub[k][4 + j], v->dub[k][4 + j], &s4, &t4);
        SIGT(CLS_UB, it->Sub[k][6 + j], it->Zub[k][6 + j], it->nub[k][6 + j], it->Vub[k][6 + j], v->dub[k][6 + j], &s6, &t6);
        /* Jacobian of (ga) wrt (up_j, u_j, T) */
        double jv[3] = {(k >= 1) ? 1.0 / h : 0.0, -1.0 / h, p->free_ ? -ga / T : 0.0};
        int ix[3] = {3 + j, 6 + j, 5};
        double ss = s4 + s6, tt = t4 - t6, zz = it->Zub[k][4 + j] - it->Zub[k][6 + j];
        for (int a = 0; a < 3; ++a) {
          r[ix[a]] += tt * jv[a];
          gL[ix[a]] -= zz * jv[a];
          for (int b = 0; b < 3; ++b) H[ix[a]][ix[b]] += ss * jv[a] * jv[b];
        }
        if (p->free_) {
          double yj = -zz; /* W -= Z * d2(+-ga) */
          double c = yj / (h * T);
          H[5][6 + j] += c; H[6 + j][5] += c;
          if (k >= 1) { H[5][3 + j] -= c; H[3 + j][5] -= c; }
          H[5][5] += yj * 2 * ga / (T * T);
        }
      }
    }
    /* -y_{k-1} on z_k */
    if (k >= 1) for (int j = 0; j < 3; ++j) gL[j] -= it->yd[k - 1][j];
    /* (6) state bounds, k >= 1 */
    if (k >= 1)
      for (int j = 0; j < 2; ++j) {
        double sg, t;
        SIGT(CLS_XY, it->Sxy[k][j], it->Zxy[k][j], it->nxy[k][j], it->Vxy[k][j], v->dxy[k][j], &sg, &t);
        H[j][j] += sg; r[j] += t; gL[j] -= it->Zxy[k][j];
        SIGT(CLS_XY, it->Sxy[k][2 + j], it->Zxy[k][2 + j], it->nxy[k][2 + j], it->Vxy[k][2 + j], v->dxy[k][2 + j], &sg, &t);
        H[j][j] += sg; r[j] -= t; gL[j] += it->Zxy[k][2 + j];
      }
    if (k == 0 && p->free_) {
      /* (4) time cost and (8) T bounds live in stage 0 */
      if (p->resto) {
        double w_ = p->zeta * dr2(p->TR);
        gf[5] += w_ * (T - p->TR); H[5][5] += w_;
      } else {
        gf[5] += (N + 1) * (P->time_cost[0] + 2 * P->time_cost[1] * T);
        H[5][5] += 2 * (N + 1) * P->time_cost[1];
      }
      double sg, t;
      SIGT(CLS_TB, it->STb[0], it->ZTb[0], it->nTb[0], it->VTb[0], v->dTb[0], &sg, &t);
      H[5][5] += sg; r[5] += t; gL[5] -= it->ZTb[0];
      SIGT(CLS_TB, it->STb[1], it->ZTb[1], it->nTb[1], it->VTb[1], v->dTb[1], &sg, &t);
      H[5][5] += sg; r[5] -= t; gL[5] += it->ZTb[1];
    }
    if (k == N && p->has_term) {
      double sg, t;
      SIGT(CLS_TM, it->Stm[0], it->Ztm[0], it->ntm[0], it->Vtm[0], v->dtm[0], &sg, &t);
      H[0][0] += sg; r[0] += t; gL[0] -= it->Ztm[0];
      SIGT(CLS_TM, it->Stm[1], it->Ztm[1], it->ntm[1], it->Vtm[1], v->dtm[1], &sg, &t);
      H[1][1] += sg; r[1] += t; gL[1] -= it->Ztm[1];
      SIGT(CLS_TM, it->Stm[2], it->Ztm[2], it->ntm[2], it->Vtm[2], v->dtm[2], &sg, &t);
      H[1][1] += sg; r[1] -= t; gL[1] += it->Ztm[2];
    }
    if (k == N && p->free_) for (int j = 0; j < 3; ++j) gL[j] += it->yt[j];
    /* (11) obstacle blocks: Schur complement onto the pose + Lagrangian gradient wrt pose */
    double Hp[3][3] = {{0}}, rp[3] = {0};
    for (int i = 0; i < p->nobs; ++i) {
      if (block_setup(p, it, v, mu, k, i, &blk[k][i], Hp, rp)) return -1;
      const blk_t* b = &blk[k][i];
      double y1 = it->ye[k][2 * i], y2 = it->ye[k][2 * i + 1];
      gL[2] += y1 * (-st * b->a1 + ct * b->a2) + y2 * (-ct * b->a1 - st * b->a2);
      for (int a = 0; a < 3; ++a) gL[a] -= it->Zd[k][i] * b->dpose[a];
    }
    for (int a = 0; a < 3; ++a) {
      r[a] += rp[a];
      for (int b = 0; b < 3; ++b) H[a][b] += Hp[a][b];
    }
    /* step form: r = -grad L + Jd^T (t - Z)  (the dist rows are carried by their own unknown in the blocks) */
    for (int a = 0; a < 8; ++a) { gL[a] += gf[a]; r[a] -= gL[a]; }
  }
  return 0;
}

/* Riccati recursion; returns 0 if every pivot is positive definite (inertia (n, m, 0)) */
typedef struct {
  double P[NS][6][6], p[NS][6], K[NS][2][6], kap[NS][2];
} ricc_t;

/* terminal equality in regularised form  dz_N - dc_a dy_a = -ct_a:  the Levenberg-Marquardt folding of the ordinary
 * iteration (dc_a = dc, ct_a = c_a) or - restoration - the elimination of pt, nt (dc_a = pt/Vpt + nt/Vnt) */
typedef struct { double dc[3], ct[3]; } termreg_t;

static int riccati(const prob_t* p, const stageqp_t* q, const vals_t* v, double dw, const termreg_t* tr, ricc_t* R) {
  int N = p->N;
  /* terminal */
  for (int a = 0; a < 6; ++a) {
    for (int b = 0; b < 6; ++b) R->P[N][a][b] = q->H[N][a][b];
    R->p[N][a] = q->r[N][a];
  }
  for (int a = 0; a < 3; ++a) {
    R->P[N][a][a] += dw;
    /* regularised terminal equality  dz_N - dc dy = -c  =>  dy = (dz_N + c)/dc */
    if (p->free_) { R->P[N][a][a] += 1.0 / tr->dc[a]; R->p[N][a] -= tr->ct[a] / tr->dc[a]; }
  }
  for (int k = N - 1; k >= 0; --k) {
    /* At = [A B] (6x8): next xi = At (xi,u) + c */
    double At[6][8];
    memset(At, 0, sizeof(At));
    At[0][0] = 1; At[1][1] = 1; At[2][2] = 1;
    At[0][2] = q->Fth[k][0]; At[1][2] = q->Fth[k][1];
    At[0][5] = q->FT[k][0]; At[1][5] = q->FT[k][1]; At[2][5] = q->FT[k][2];
    At[5][5] = 1;
    At[0][6] = q->Bv[k][0]; At[1][6] = q->Bv[k][1]; At[2][7] = q->Bw[k];
    At[3][6] = 1; At[4][7] = 1;
    double PA[6][8], F[8][8], f[8], pc[6];
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 8; ++b) {
        double s = 0;
        for (int c = 0; c < 6; ++c) s += R->P[k + 1][a][c] * At[c][b];
        PA[a][b] = s;
      }
    for (int a = 0; a < 6; ++a) {
      double s = R->p[k + 1][a];
      for (int c = 0; c < 3; ++c) s -= R->P[k + 1][a][c] * v->cd[k][c];
      pc[a] = s;
    }
    for (int a = 0; a < 8; ++a) {
      for (int b = 0; b < 8; ++b) {
        double s = q->H[k][a][b];
        for (int c = 0; c < 6; ++c) s += At[c][a] * PA[c][b];
        F[a][b] = s;
      }
      double s = q->r[k][a];
      for (int c = 0; c < 6; ++c) s += At[c][a] * pc[c];
      f[a] = s;
    }
    /* regularisation on the real variables of this stage: z_k (k >= 1), u_k, T (once, stage 0) */
    if (k >= 1) for (int a = 0; a < 3; ++a) F[a][a] += dw;
    F[6][6] += dw; F[7][7] += dw;
    if (k == 0 && p->free_) F[5][5] += dw;
    double q00 = F[6][6], q01 = 0.5 * (F[6][7] + F[7][6]), q11 = F[7][7];
    double det = q00 * q11 - q01 * q01;
    if (!(q00 > 0) || !(det > 0)) return -1;
    double i00 = q11 / det, i01 = -q01 / det, i11 = q00 / det;
    for (int b = 0; b < 6; ++b) {
      R->K[k][0][b] = -(i00 * F[6][b] + i01 * F[7][b]);
      R->K[k][1][b] = -(i01 * F[6][b] + i11 * F[7][b]);
    }
    R->kap[k][0] = i00 * f[6] + i01 * f[7];
    R->kap[k][1] = i01 * f[6] + i11 * f[7];
    for (int a = 0; a < 6; ++a) {
      for (int b = 0; b < 6; ++b) R->P[k][a][b] = F[a][b] + F[a][6] * R->K[k][0][b] + F[a][7] * R->K[k][1][b];
      R->p[k][a] = f[a] - F[a][6] * R->kap[k][0] - F[a][7] * R->kap[k][1];
    }
    for (int a = 0; a < 6; ++a)
      for (int b = a + 1; b < 6; ++b) R->P[k][a][b] = R->P[k][b][a] = 0.5 * (R->P[k][a][b] + R->P[k][b][a]);
  }
  if (p->free_ && !(R->P[0][5][5] > 0)) return -1;
  return 0;
}

static void forward(const prob_t* p, const iter_t* it, const stageqp_t* q, const vals_t* v, const ricc_t* R,
                    const termreg_t* tr, dir_t* d) {
  int N = p->N;
  double xi[6] = {0, 0, 0, 0, 0, 0};
  if (p->free_) xi[5] = R->p[0][5] / R->P[0][5][5];
  d->T = xi[5];
  for (int j = 0; j < 3; ++j) d->z[0][j] = 0;
  for (int k = 0; k < N; ++k) {
    double du[2];
    for (int a = 0; a < 2; ++a) {
      double s = R->kap[k][a];
      for (int b = 0; b < 6; ++b) s += R->K[k][a][b] * xi[b];
      du[a] = s;
    }
    d->u[k][0] = du[0]; d->u[k][1] = du[1];
    double xn[6];
    xn[0] = xi[0] + q->Fth[k][0] * xi[2] + q->FT[k][0] * xi[5] + q->Bv[k][0] * du[0] + v->cd[k][0];
    xn[1] = xi[1] + q->Fth[k][1] * xi[2] + q->FT[k][1] * xi[5] + q->Bv[k][1] * du[0] + v->cd[k][1];
    xn[2] = xi[2] + q->FT[k][2] * xi[5] + q->Bw[k] * du[1] + v->cd[k][2];
    xn[3] = du[0]; xn[4] = du[1]; xn[5] = xi[5];
    for (int a = 0; a < 3; ++a) {
      double s = -R->p[k + 1][a];
      for (int b = 0; b < 6; ++b) s += R->P[k + 1][a][b] * xn[b];
      d->yd[k][a] = it->yd[k][a] + s; /* y_k + dy_k */
    }
    /* the dw / 1/dc terms were folded into P[N] only; for k+1 < N the regularisation dw on z_{k+1} was
     * added to F at stage k+1 (inside P[k+1] already) */
    memcpy(xi, xn, sizeof(xi));
    for (int j = 0; j < 3; ++j) d->z[k + 1][j] = xn[j];
  }
  if (p->free_) for (int j = 0; j < 3; ++j) d->yt[j] = it->yt[j] + (d->z[N][j] + tr->ct[j]) / tr->dc[j];
}

/* ------------------------------------------------------------------------------------------------
 * the interior-point loop (oracle/ipm_dense.py solve(), soc = False)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  iter_t it, tr, best, wd, fail; /* best: stored acceptable point; wd: watchdog reference iterate; fail: point of failure */
  vals_t v, vt;
  stageqp_t q;
  blk_t blk[NS][OM];
  ricc_t R;
  dir_t d;
} work_t;

typedef void (*trace_fn)(int it, double f, double th, double E0, double mu, double dw, double alpha);
static trace_fn g_trace = 0;
/* restoration phase: required reduction of the violation per call (IPOPT's kappa_resto is 0.9 together with its filter;
 * here the filter restarts, so the reduction asked for is larger), penalty rho (IPOPT: 1000), relaxed row classes
 * (state box, terminal set, OBCA distance rows; relaxing the input / acceleration / norm rows as well made no
 * difference on the BASELINE batches and the T bounds must stay hard - measured), rounds per attempt.
 * The setter is a developer hook for exactly those measurements. */
static double g_kappa_resto = 0.1, g_resto_tol = 1e-8, g_feas_tol = 1e-6, g_rho = 1000.0;
static int g_rmask = CLS_XY | CLS_TM | CLS_DIST, g_max_resto = 2, g_post_mode = 4;
void obca_oracle_set_resto(double kappa, double rtol, double rho, int rmask, int max_resto, int post_mode) {
  g_kappa_resto = kappa; g_resto_tol = rtol; g_rho = rho; g_rmask = rmask; g_max_resto = max_resto; g_post_mode = post_mode;
}
static int g_stall_iters = 10;
void obca_oracle_set_stall(int n) { g_stall_iters = n; }
static int g_budget = OBCA_RECOVERY_BUDGET;
void obca_oracle_set_budget(int n) { g_budget = n; }
static int g_verbose = 0;
void obca_oracle_set_verbose(int v) { g_verbose = v; }
static int g_acc_stall = 10;                  /* iterations without halving the error at the acceptable level */
void obca_oracle_set_acc_stall(int n) { g_acc_stall = n; }
static int g_wd_trigger = 10, g_wd_max = 3;   /* watchdog: shortened steps before it starts / full steps on trust */
void obca_oracle_set_watchdog(int trigger, int max_trust) { g_wd_trigger = trigger; g_wd_max = max_trust; }
void obca_oracle_set_trace(trace_fn fn) { g_trace = fn; }

static void apply_step(const prob_t* p, const iter_t* it, const dir_t* d, double a, iter_t* o) {
  int N = p->N;
  if (o != it) memcpy(o, it, sizeof(*o));
  for (int k = 0; k <= N; ++k) {
    if (k >= 1) for (int j = 0; j < 3; ++j) o->z[k][j] = it->z[k][j] + a * d->z[k][j];
    if (k < N) for (int j = 0; j < 2; ++j) o->u[k][j] = it->u[k][j] + a * d->u[k][j];
    for (int r = 0; r < p->R; ++r) { o->lam[k][r] = it->lam[k][r] + a * d->lam[k][r]; o->Sl[k][r] = it->Sl[k][r] + a * d->Sl[k][r]; }
    for (int r = 0; r < 4 * p->nobs; ++r) { o->mu[k][r] = it->mu[k][r] + a * d->mu[k][r]; o->Sm[k][r] = it->Sm[k][r] + a * d->Sm[k][r]; }
    for (int i = 0; i < p->nobs; ++i) { o->Sn[k][i] = it->Sn[k][i] + a * d->Sn[k][i]; o->Sd[k][i] = it->Sd[k][i] + a * d->Sd[k][i]; }
    if (k >= 1) for (int j = 0; j < 4; ++j) o->Sxy[k][j] = it->Sxy[k][j] + a * d->Sxy[k][j];
    if (k < N) for (int j = 0; j < 8; ++j) o->Sub[k][j] = it->Sub[k][j] + a * d->Sub[k][j];
  }
  if (p->free_) {
    o->T = it->T + a * d->T;
    for (int j = 0; j < 2; ++j) o->STb[j] = it->STb[j] + a * d->STb[j];
  }
  if (p->has_term) for (int j = 0; j < 3; ++j) o->Stm[j] = it->Stm[j] + a * d->Stm[j];
  if (p->resto) {
    for (int k = 0; k <= N; ++k) {
      if (k >= 1) for (int j = 0; j < 4; ++j) o->nxy[k][j] = it->nxy[k][j] + a * d->nxy[k][j];
      if (k < N) for (int j = 0; j < 8; ++j) o->nub[k][j] = it->nub[k][j] + a * d->nub[k][j];
      for (int i = 0; i < p->nobs; ++i) { o->nn[k][i] = it->nn[k][i] + a * d->nn[k][i]; o->nd[k][i] = it->nd[k][i] + a * d->nd[k][i]; }
    }
    for (int j = 0; j < 2; ++j) o->nTb[j] = it->nTb[j] + a * d->nTb[j];
    for (int j = 0; j < 3; ++j) {
      o->ntm[j] = it->ntm[j] + a * d->ntm[j];
      o->pt[j] = it->pt[j] + a * d->pt[j]; o->nt[j] = it->nt[j] + a * d->nt[j];
    }
  }
}

static double objective_of(const prob_t* p, const iter_t* it) {
  vals_t* v = (vals_t*)malloc(sizeof(vals_t));
  eval_values(p, it, v);
  double f = v->f;
  free(v);
  return f;
}

static int solve_one(const prob_t* p, work_t* w, int* iters_out, double* err_out, double* mu_out) {
  const obca_params* P = p->P;
  int N = p->N;
  iter_t* it = &w->it;
  vals_t* v = &w->v;
  dir_t* d = &w->d;
  const double s_max = 100.0, kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99, kappa_sigma = 1e10;
  const double dw_first = 1e-4, dw_min = 1e-20, dw_max = 1e20, kw_plus_first = 100.0, kw_plus = 8.0, kw_minus = 1.0 / 3.0;
  const double dc_min = 1e-8, lm_cap = 1e4, stall_alpha = 1e-3;
  const int stall_iters = g_stall_iters;
  const double g_th = 1e-5, g_ph = 1e-8, s_th = 1.1, s_ph = 2.3, eta_ph = 1e-8;
  double tol = P->tol;

  double mu = (p->mu0 > 0) ? p->mu0 : P->mu_init;
  if (g_verbose) fprintf(stderr, "  -- pass: resto %d init %d mu %.2e\n", p->resto, p->init, mu);
  if (p->resto) {
    /* restoration starts from the iterate of the failed pass (IPOPT: x, s kept; n, p on the central path of the
     * residual they absorb so that the relaxed rows start satisfied; bound multipliers min(rho, z); y = 0) */
    if (g_post_mode & 4) {
      /* project the point onto the rows the restoration problem keeps hard, so that it starts feasible for its own
       * constraints: poses by rolling the inputs out through the dynamics, OBCA duals pushed inside their sign
       * bounds, mu_1..4 so that the two OBCA equalities hold */
      double T = p->free_ ? it->T : 1.0, h = T * p->Ts;
      for (int k = 0; k < N; ++k) {
        const double* z = it->z[k];
        it->z[k + 1][0] = z[0] + h * it->u[k][0] * cos(z[2]);
        it->z[k + 1][1] = z[1] + h * it->u[k][0] * sin(z[2]);
        it->z[k + 1][2] = z[2] + h * it->u[k][1];
      }
      for (int k = 0; k <= N; ++k) {
        double ct = cos(it->z[k][2]), st = sin(it->z[k][2]);
        for (int i = 0; i < p->nobs; ++i) {
          double a1 = 0, a2 = 0;
          for (int r = p->eptr[i]; r < p->eptr[i + 1]; ++r) {
            it->lam[k][r] = fmax(it->lam[k][r], P->bound_push);
            a1 += p->A[2 * r] * it->lam[k][r]; a2 += p->A[2 * r + 1] * it->lam[k][r];
          }
          double c1 = ct * a1 + st * a2, c2 = -st * a1 + ct * a2;
          double* m = &it->mu[k][4 * i];
          double b1 = fmax(fmin(m[0], m[2]), P->bound_push), b2 = fmax(fmin(m[1], m[3]), P->bound_push);
          m[0] = b1 + fmax(-c1, 0); m[2] = b1 + fmax(c1, 0);
          m[1] = b2 + fmax(-c2, 0); m[3] = b2 + fmax(c2, 0);
        }
      }
    }
    eval_values(p, it, v);
    memset(it->yd, 0, sizeof(it->yd)); memset(it->yt, 0, sizeof(it->yt)); memset(it->ye, 0, sizeof(it->ye));
    const double rho = p->rho;
    FOR_INEQ(p, it, v, {
      if (g_post_mode & 2) *Z_ = fmin(*Z_, rho); else { *Z_ = 1.0; *S_ = fmax(d_, P->bound_push); }
      if (n_) {
        double c_ = d_ - *S_, h_ = (mu - rho * c_) / (2 * rho);
        *n_ = h_ + sqrt(h_ * h_ + mu * c_ / (2 * rho));      /* c - p + n = 0 with p n on the central path   */
        *S_ += c_ + *n_;                                     /* the slack absorbs p (it carries no cost)       */
        *V_ = mu / *n_;
      } });
    if (p->free_) for (int j = 0; j < 3; ++j) {
      double c_ = v->ct[j], h_ = (mu - rho * c_) / (2 * rho);
      it->nt[j] = h_ + sqrt(h_ * h_ + mu * c_ / (2 * rho));
      it->pt[j] = c_ + it->nt[j];
      it->Vpt[j] = mu / it->pt[j]; it->Vnt[j] = mu / it->nt[j];
    }
  } else {
    start_point(p, it);
    eval_values(p, it, v);
    FOR_INEQ(p, it, v, { *S_ = fmax(d_, P->bound_push); *Z_ = 1.0; });
  }
  filt_t F;
  memset(&F, 0, sizeof(F));
  int nstall = 0, acc_count = 0, iter = 0, status = OBCA_ST_MAXITER;
  double dw_last = 0.0, E0 = 0, best_E0 = 1e300, e_min = 1e300;
  int e_min_iter = 0, best_lvl = 0;
  /* watchdog (Chamberlain et al.; IPOPT's watchdog, triggered earlier): after WD_TRIGGER consecutive shortened steps a
   * rejected full step is taken anyway from a saved reference iterate; if within WD_MAX further full steps no point
   * acceptable to the reference is reached, the reference is restored and ordinary backtracking resumes there */
  const int WD_TRIGGER = g_wd_trigger, WD_MAX = p->resto ? 0 : g_wd_max;   /* no watchdog in the restoration pass */
  int in_wd = 0, wd_count = 0, wd_block = 0, n_short = 0;
  double wd_th = 0, wd_ph = 0, wd_dphi = 0, wd_alpha = 1;
  int m_eq = 3 * N + (p->free_ ? 3 : 0) + 2 * p->nobs * (N + 1);
  int q_in = 4 * N + 8 * N + (p->free_ ? 2 : 0) + (p->has_term ? 3 : 0) + (p->R + 6 * p->nobs) * (N + 1);
  if (p->resto) { /* the bounds n >= 0 (and pt, nt >= 0) count as inequalities of the restoration problem */
    iter_t* itc = it;
    FOR_INEQ(p, itc, v, { if (n_) q_in++; });
    if (p->free_) q_in += 6;
  }
  double th_orig = 0, rs_best = 1e300;
  int rs_iter = 0;
  enum { RESTO_STALL = 15, RESTO_MAXITER = 200 };

  for (;;) {
    eval_values(p, it, v);
    /* assemble at the CURRENT mu first for the optimality error (gL, c, d do not depend on mu) */
    if (assemble(p, it, v, mu, &w->q, w->blk)) { status = OBCA_ST_REGFAIL; break; }
    /* optimality error pieces */
    double e1 = 0, e2 = 0, sumy = 0, sumz = 0, szmax = 0, szmin = 1e300, th, ph0, cmax;
    {
      double gT = 0;
      for (int k = 0; k <= N; ++k) {
        if (k >= 1) for (int j = 0; j < 3; ++j) e1 = fmax(e1, fabs(w->q.gL[k][j]));
        if (k < N) for (int j = 0; j < 2; ++j) e1 = fmax(e1, fabs(w->q.gL[k][6 + j] + ((k + 1 < N) ? w->q.gL[k + 1][3 + j] : 0.0)));
        gT += w->q.gL[k][5];
        if (k < N) for (int j = 0; j < 3; ++j) sumy += fabs(it->yd[k][j]);
        for (int j = 0; j < 2 * p->nobs; ++j) sumy += fabs(it->ye[k][j]);
        /* Lagrangian gradient wrt the block variables */
        const double* z = it->z[k];
        double ct = cos(z[2]), st = sin(z[2]), tx = z[0] + p->off * ct, ty = z[1] + p->off * st;
        for (int i = 0; i < p->nobs; ++i) {
          const blk_t* b = &w->blk[k][i];
          double y1 = it->ye[k][2 * i], y2 = it->ye[k][2 * i + 1];
          for (int r = p->eptr[i]; r < p->eptr[i + 1]; ++r) {
            double A0 = p->A[2 * r], A1 = p->A[2 * r + 1];
            double gl = y1 * (ct * A0 + st * A1) + y2 * (-st * A0 + ct * A1) - it->Zl[k][r] + it->Zn[k][i] * 2 * (b->a1 * A0 + b->a2 * A1) -
                        it->Zd[k][i] * (tx * A0 + ty * A1 - bk(p, k, r));
            e1 = fmax(e1, fabs(gl));
          }
          for (int m = 0; m < 4; ++m) {
            double je1 = (m == 0) ? 1.0 : (m == 2) ? -1.0 : 0.0, je2 = (m == 1) ? 1.0 : (m == 3) ? -1.0 : 0.0;
            double gl = y1 * je1 + y2 * je2 - it->Zm[k][4 * i + m] + it->Zd[k][i] * p->g[m];
            e1 = fmax(e1, fabs(gl));
          }
        }
      }
      if (p->free_) { e1 = fmax(e1, fabs(gT)); for (int j = 0; j < 3; ++j) sumy += fabs(it->yt[j]); }
      FOR_INEQ(p, it, v, { (void)d_; sumz += *Z_; double sz = (*S_) * (*Z_); szmax = fmax(szmax, sz); szmin = fmin(szmin, sz);
        if (n_) { sumz += *V_; sz = (*n_) * (*V_); szmax = fmax(szmax, sz); szmin = fmin(szmin, sz);
                  e1 = fmax(e1, fabs(p->rho - *Z_ - *V_)); } });
      if (p->resto && p->free_) for (int j = 0; j < 3; ++j) {
        double sp = it->pt[j] * it->Vpt[j], sn_ = it->nt[j] * it->Vnt[j];
        sumz += it->Vpt[j] + it->Vnt[j];
        szmax = fmax(szmax, fmax(sp, sn_)); szmin = fmin(szmin, fmin(sp, sn_));
        e1 = fmax(e1, fmax(fabs(p->rho - it->yt[j] - it->Vpt[j]), fabs(p->rho + it->yt[j] - it->Vnt[j])));
      }
      theta_phi(p, it, v, mu, &th, &ph0, &cmax, &th_orig);
      e2 = cmax;
    }
    double sd = fmax(s_max, (sumy + sumz) / (m_eq + q_in)) / s_max, sc = fmax(s_max, sumz / q_in) / s_max;
    E0 = fmax(fmax(e1 / sd, e2), szmax / sc);
    if (p->resto) {
      /* the restoration phase ends as soon as the violation of the original problem has dropped to kappa_resto times
       * what it was (IPOPT: 0.9, plus acceptance by the filter of the original problem - here that filter starts
       * afresh); if instead its own problem converges the point is a local minimiser of the violation */
      if (th_orig <= fmax(g_kappa_resto * p->th_ref, g_feas_tol)) { status = OBCA_ST_OK; break; }
      if (E0 <= fmax(tol, g_resto_tol)) { status = OBCA_ST_INFEASIBLE; break; }
      /* the violation has stopped decreasing (1 % in RESTO_STALL iterations) although the restoration problem's own
       * constraints hold: a local minimiser of the violation, reported as such (its dual error sits on the noise floor
       * of the regularised steps and would never reach tol) */
      if (th_orig < 0.99 * rs_best) { rs_best = th_orig; rs_iter = iter; }
      if (iter - rs_iter >= RESTO_STALL && th <= 1e-6 * fmax(1.0, th_orig)) { status = OBCA_ST_INFEASIBLE; break; }
      if (iter >= RESTO_MAXITER) { status = OBCA_ST_RESTOFAIL; break; }
    } else {
    if (E0 <= tol) { status = OBCA_ST_OK; break; }
    if (E0 <= P->acceptable_tol) {
      if (++acc_count >= P->acceptable_iter) { status = OBCA_ST_ACCEPTABLE; break; }
    } else
      acc_count = 0;
    }
    /* stall at the acceptable level: an acceptable point is stored, the barrier parameter is final and the error has
     * not halved for ACC_STALL iterations - the iterate is wandering on the noise floor (objective constant to 10
     * digits).  End like IPOPT does when it cannot progress from an acceptable point: with the stored point. */
    if (best_E0 < 1e300 && mu <= tol / 10 * (1 + 1e-12) && iter - e_min_iter >= g_acc_stall) { status = OBCA_ST_LSFAIL; break; }
    if (iter >= P->max_iter) { status = OBCA_ST_MAXITER; break; }
    /* barrier update */
    int changed = 0;
    for (;;) {
      double e3 = fmax(szmax - mu, mu - szmin) / sc;
      double Emu = fmax(fmax(e1 / sd, e2), e3);
      if (Emu <= kappa_eps * mu && mu > tol / 10) {
        mu = fmax(tol / 10, fmin(kappa_mu * mu, pow(mu, theta_mu)));
        changed = 1;
      } else
        break;
    }
    if (changed) {
      in_wd = 0; /* a new barrier problem: the current point becomes an ordinary iterate */
      if (F.active) { F.n = 0; F.wr = 0; }
      if (assemble(p, it, v, mu, &w->q, w->blk)) { status = OBCA_ST_REGFAIL; break; }
      theta_phi(p, it, v, mu, &th, &ph0, &cmax, 0);
    }
    /* acceptable level: IPOPT's acceptable tolerance, or - at the final barrier parameter - primal feasible to 1e-6,
     * complementary, with only the dual infeasibility above tol (the rounding-noise floor of a degenerate vertex of
     * the OBCA dual polytope; same condition as at_floor below).  Judged after the barrier update: the
     * iteration that lowers mu to its final value already counts */
    const int acc_lvl = (E0 <= P->acceptable_tol) ? OBCA_ST_ACCEPTABLE : (mu <= 1e-6 && th <= 1e-6 && fmax(E0, e1) <= 1e3 * P->acceptable_tol) ? OBCA_ST_FLOOR : 0;
    if (acc_lvl && !p->resto) {
      /* IPOPT stores the best acceptable iterate and falls back to it when the run ends in a failure
       * ("Solved To Acceptable Level"); a point that only meets the noise-floor level is reported as such */
      if (E0 < 0.1 * best_E0) { best_E0 = E0; best_lvl = acc_lvl; w->best = *it; }   /* a new copy per decade of improvement */
      if (E0 < 0.5 * e_min) { e_min = E0; e_min_iter = iter; }
    }
    double tau = fmax(tau_min, 1 - mu);
    termreg_t trg;
    memset(&trg, 0, sizeof(trg));
    if (p->free_ && p->resto) {
      for (int j = 0; j < 3; ++j) {
        double pp = it->pt[j], nn_ = it->nt[j], Vp = it->Vpt[j], Vn = it->Vnt[j], y = it->yt[j];
        trg.dc[j] = pp / Vp + nn_ / Vn;
        trg.ct[j] = (v->ct[j] - pp + nn_) - (mu - pp * (p->rho - y)) / Vp + (mu - nn_ * (p->rho + y)) / Vn;
      }
    } else if (p->free_) {
      double cm = fmax(fabs(v->ct[0]), fmax(fabs(v->ct[1]), fabs(v->ct[2])));
      double dc = fmax(dc_min, cm / lm_cap);
      for (int j = 0; j < 3; ++j) { trg.dc[j] = dc; trg.ct[j] = v->ct[j]; }
    }
    /* inertia correction */
    double dw = 0.0;
    int regfail = 0;
    for (;;) {
      if (riccati(p, &w->q, v, dw, &trg, &w->R) == 0) break;
      if (dw == 0.0)
        dw = (dw_last == 0.0) ? dw_first : fmax(dw_min, kw_minus * dw_last);
      else
        dw = dw * ((dw_last == 0.0) ? kw_plus_first : kw_plus);
      if (dw > dw_max) { regfail = 1; break; }
    }
    if (regfail) { status = OBCA_ST_REGFAIL; break; }
    if (dw > 0) dw_last = dw;
    forward(p, it, &w->q, v, &w->R, &trg, d);
    /* back-substitute the blocks; slack directions */
    double Dphi = 0;
    for (int k = 0; k <= N; ++k) {
      for (int i = 0; i < p->nobs; ++i) block_backsub(p, it, v, mu, k, i, &w->blk[k][i], d->z[k], d);
      if (k >= 1)
        for (int j = 0; j < 2; ++j) {
          d->Sxy[k][j] = d->z[k][j] + (v->dxy[k][j] - it->Sxy[k][j]);
          d->Sxy[k][2 + j] = -d->z[k][j] + (v->dxy[k][2 + j] - it->Sxy[k][2 + j]);
        }
      if (k < N) {
        double T = p->free_ ? it->T : 1.0, h = T * p->Ts;
        const double* up = (k == 0) ? p->u0 : it->u[k - 1];
        for (int j = 0; j < 2; ++j) {
          d->Sub[k][j] = d->u[k][j] + (v->dub[k][j] - it->Sub[k][j]);
          d->Sub[k][2 + j] = -d->u[k][j] + (v->dub[k][2 + j] - it->Sub[k][2 + j]);
          double ga = (up[j] - it->u[k][j]) / h;
          double dga = (((k >= 1) ? d->u[k - 1][j] : 0.0) - d->u[k][j]) / h - (p->free_ ? ga / T * d->T : 0.0);
          d->Sub[k][4 + j] = dga + (v->dub[k][4 + j] - it->Sub[k][4 + j]);
          d->Sub[k][6 + j] = -dga + (v->dub[k][6 + j] - it->Sub[k][6 + j]);
        }
      }
      /* objective directional derivative: gf over (z_k, u_{k-1}, T, u_k) */
      const double* gf = w->q.gf[k];
      if (k >= 1) for (int j = 0; j < 3; ++j) Dphi += gf[j] * d->z[k][j];
      if (k >= 1 && k < N) for (int j = 0; j < 2; ++j) Dphi += gf[3 + j] * d->u[k - 1][j];
      if (p->free_) Dphi += gf[5] * d->T;
      if (k < N) for (int j = 0; j < 2; ++j) Dphi += gf[6 + j] * d->u[k][j];
    }
    if (p->free_) {
      d->STb[0] = d->T + (v->dTb[0] - it->STb[0]);
      d->STb[1] = -d->T + (v->dTb[1] - it->STb[1]);
    }
    if (p->has_term) {
      d->Stm[0] = d->z[N][0] + (v->dtm[0] - it->Stm[0]);
      d->Stm[1] = d->z[N][1] + (v->dtm[1] - it->Stm[1]);
      d->Stm[2] = -d->z[N][1] + (v->dtm[2] - it->Stm[2]);
    }
    if (p->resto) {
      /* relaxed rows: the ordinary slack step above is g + (d - S) with g = grad d . dX; the row's multiplier step is
       * dZ = t - sigma g, and slack and relaxation follow from their complementarity conditions */
#define RELAX_FIX(cls, Sx, Zx, nx, Vx, dx, dSx, dnx)                                                  \
      do { if (RELAXED(p, cls)) { double sg_, t_, g_ = (dSx) - ((dx) - (Sx));                          \
        sig_t_relaxed(Sx, Zx, nx, Vx, dx, mu, p->rho, &sg_, &t_);                                      \
        relaxed_steps(Sx, Zx, nx, Vx, mu, p->rho, t_ - sg_ * g_, &(dSx), &(dnx)); } } while (0)
      for (int k = 0; k <= N; ++k) {
        if (k >= 1) for (int j = 0; j < 4; ++j) RELAX_FIX(CLS_XY, it->Sxy[k][j], it->Zxy[k][j], it->nxy[k][j], it->Vxy[k][j], v->dxy[k][j], d->Sxy[k][j], d->nxy[k][j]);
        if (k < N) for (int j = 0; j < 8; ++j) RELAX_FIX(CLS_UB, it->Sub[k][j], it->Zub[k][j], it->nub[k][j], it->Vub[k][j], v->dub[k][j], d->Sub[k][j], d->nub[k][j]);
      }
      if (p->free_) for (int j = 0; j < 2; ++j) RELAX_FIX(CLS_TB, it->STb[j], it->ZTb[j], it->nTb[j], it->VTb[j], v->dTb[j], d->STb[j], d->nTb[j]);
      if (p->has_term) for (int j = 0; j < 3; ++j) RELAX_FIX(CLS_TM, it->Stm[j], it->Ztm[j], it->ntm[j], it->Vtm[j], v->dtm[j], d->Stm[j], d->ntm[j]);
#undef RELAX_FIX
      if (p->free_) for (int j = 0; j < 3; ++j) {
        double dy = d->yt[j] - it->yt[j], y = it->yt[j];
        double dVp = (p->rho - y - it->Vpt[j]) - dy, dVn = (p->rho + y - it->Vnt[j]) + dy;
        d->pt[j] = (mu - it->pt[j] * it->Vpt[j] - it->pt[j] * dVp) / it->Vpt[j];
        d->nt[j] = (mu - it->nt[j] * it->Vnt[j] - it->nt[j] * dVn) / it->Vnt[j];
      }
    }
    /* fraction to the boundary on S and Z;  dZ = mu/S - Z - Sigma dS  (not stored) */
    double a_max = 1.0, a_z = 1.0, sls = 0;
    {
      iter_t* dS = (iter_t*)0; (void)dS;
      /* walk S/Z and the matching dS entries in the same order */
#define STEP_INEQ(Sarr, Zarr, dSarr)                                           \
      do { double S_ = (Sarr), Z_ = (Zarr), dS_ = (dSarr);                     \
        double dZ_ = mu / S_ - Z_ - (Z_ / S_) * dS_;                            \
        if (dS_ < 0) a_max = fmin(a_max, -tau * S_ / dS_);                      \
        if (dZ_ < 0) a_z = fmin(a_z, -tau * Z_ / dZ_);                          \
        sls += dS_ / S_; } while (0)
      for (int k = 0; k <= N; ++k) {
        if (k >= 1) for (int j = 0; j < 4; ++j) STEP_INEQ(it->Sxy[k][j], it->Zxy[k][j], d->Sxy[k][j]);
        if (k < N) for (int j = 0; j < 8; ++j) STEP_INEQ(it->Sub[k][j], it->Zub[k][j], d->Sub[k][j]);
        for (int r = 0; r < p->R; ++r) STEP_INEQ(it->Sl[k][r], it->Zl[k][r], d->Sl[k][r]);
        for (int r = 0; r < 4 * p->nobs; ++r) STEP_INEQ(it->Sm[k][r], it->Zm[k][r], d->Sm[k][r]);
        for (int i = 0; i < p->nobs; ++i) STEP_INEQ(it->Sn[k][i], it->Zn[k][i], d->Sn[k][i]);
        for (int i = 0; i < p->nobs; ++i) STEP_INEQ(it->Sd[k][i], it->Zd[k][i], d->Sd[k][i]);
      }
      if (p->free_) for (int j = 0; j < 2; ++j) STEP_INEQ(it->STb[j], it->ZTb[j], d->STb[j]);
      if (p->has_term) for (int j = 0; j < 3; ++j) STEP_INEQ(it->Stm[j], it->Ztm[j], d->Stm[j]);
    }
    if (p->resto) {
      /* the bounds n >= 0 and their multipliers V (dV = (rho - Z - V) - dZ); cost rho n enters the directional derivative */
      double rdn = 0;
#define STEP_RELAX(cls, Sx, Zx, dSx, nx, Vx, dnx)                                                      \
      do { if (RELAXED(p, cls)) { double dZ_ = mu / (Sx) - (Zx) - ((Zx) / (Sx)) * (dSx);               \
        double dV_ = (p->rho - (Zx) - (Vx)) - dZ_;                                                     \
        if ((dnx) < 0) a_max = fmin(a_max, -tau * (nx) / (dnx));                                       \
        if (dV_ < 0) a_z = fmin(a_z, -tau * (Vx) / dV_);                                               \
        sls += (dnx) / (nx); rdn += (dnx); } } while (0)
      for (int k = 0; k <= N; ++k) {
        if (k >= 1) for (int j = 0; j < 4; ++j) STEP_RELAX(CLS_XY, it->Sxy[k][j], it->Zxy[k][j], d->Sxy[k][j], it->nxy[k][j], it->Vxy[k][j], d->nxy[k][j]);
        if (k < N) for (int j = 0; j < 8; ++j) STEP_RELAX(CLS_UB, it->Sub[k][j], it->Zub[k][j], d->Sub[k][j], it->nub[k][j], it->Vub[k][j], d->nub[k][j]);
        for (int i = 0; i < p->nobs; ++i) STEP_RELAX(CLS_NORM, it->Sn[k][i], it->Zn[k][i], d->Sn[k][i], it->nn[k][i], it->Vn[k][i], d->nn[k][i]);
        for (int i = 0; i < p->nobs; ++i) STEP_RELAX(CLS_DIST, it->Sd[k][i], it->Zd[k][i], d->Sd[k][i], it->nd[k][i], it->Vd[k][i], d->nd[k][i]);
      }
      if (p->free_) for (int j = 0; j < 2; ++j) STEP_RELAX(CLS_TB, it->STb[j], it->ZTb[j], d->STb[j], it->nTb[j], it->VTb[j], d->nTb[j]);
      if (p->has_term) for (int j = 0; j < 3; ++j) STEP_RELAX(CLS_TM, it->Stm[j], it->Ztm[j], d->Stm[j], it->ntm[j], it->Vtm[j], d->ntm[j]);
#undef STEP_RELAX
      if (p->free_) for (int j = 0; j < 3; ++j) {
        double dy = d->yt[j] - it->yt[j], y = it->yt[j];
        double dVp = (p->rho - y - it->Vpt[j]) - dy, dVn = (p->rho + y - it->Vnt[j]) + dy;
        if (d->pt[j] < 0) a_max = fmin(a_max, -tau * it->pt[j] / d->pt[j]);
        if (d->nt[j] < 0) a_max = fmin(a_max, -tau * it->nt[j] / d->nt[j]);
        if (dVp < 0) a_z = fmin(a_z, -tau * it->Vpt[j] / dVp);
        if (dVn < 0) a_z = fmin(a_z, -tau * it->Vnt[j] / dVn);
        sls += d->pt[j] / it->pt[j] + d->nt[j] / it->nt[j]; rdn += d->pt[j] + d->nt[j];
      }
      Dphi += p->rho * rdn;
    }
    Dphi -= mu * sls;
    if (g_verbose > 1) { /* DEBUG: directional consistency of the Newton step by finite differences */
      double eps = 1e-7;
      apply_step(p, it, d, eps, &w->tr);
      eval_values(p, &w->tr, &w->vt);
      double worst[8] = {0}; const char* nm[8] = {"dyn", "e", "term", "xy", "ub", "norm", "dist", "Tb/tm"};
      for (int k = 0; k <= N; ++k) {
        if (k < N) for (int j = 0; j < 3; ++j) worst[0] = fmax(worst[0], fabs((w->vt.cd[k][j] - (1 - eps) * v->cd[k][j]) / eps));
        for (int j = 0; j < 2 * p->nobs; ++j) worst[1] = fmax(worst[1], fabs((w->vt.ce[k][j] - (1 - eps) * v->ce[k][j]) / eps));
#define RES(dv, S, n, rel) ((dv) - (S) + ((rel) ? (n) : 0.0))
        if (k >= 1) for (int j = 0; j < 4; ++j) worst[3] = fmax(worst[3], fabs((RES(w->vt.dxy[k][j], w->tr.Sxy[k][j], w->tr.nxy[k][j], RELAXED(p, CLS_XY)) - (1 - eps) * RES(v->dxy[k][j], it->Sxy[k][j], it->nxy[k][j], RELAXED(p, CLS_XY))) / eps));
        if (k < N) for (int j = 0; j < 8; ++j) worst[4] = fmax(worst[4], fabs((RES(w->vt.dub[k][j], w->tr.Sub[k][j], w->tr.nub[k][j], RELAXED(p, CLS_UB)) - (1 - eps) * RES(v->dub[k][j], it->Sub[k][j], it->nub[k][j], RELAXED(p, CLS_UB))) / eps));
        for (int i = 0; i < p->nobs; ++i) {
          worst[5] = fmax(worst[5], fabs((RES(w->vt.dn[k][i], w->tr.Sn[k][i], w->tr.nn[k][i], RELAXED(p, CLS_NORM)) - (1 - eps) * RES(v->dn[k][i], it->Sn[k][i], it->nn[k][i], RELAXED(p, CLS_NORM))) / eps));
          worst[6] = fmax(worst[6], fabs((RES(w->vt.dd[k][i], w->tr.Sd[k][i], w->tr.nd[k][i], RELAXED(p, CLS_DIST)) - (1 - eps) * RES(v->dd[k][i], it->Sd[k][i], it->nd[k][i], RELAXED(p, CLS_DIST))) / eps));
        }
      }
      if (p->free_) for (int j = 0; j < 3; ++j) {
        double r1 = w->vt.ct[j] + (p->resto ? -w->tr.pt[j] + w->tr.nt[j] : 0), r0 = v->ct[j] + (p->resto ? -it->pt[j] + it->nt[j] : 0);
        worst[2] = fmax(worst[2], fabs((r1 - (1 - eps) * r0) / eps));
      }
      fprintf(stderr, "      lin-check:");
      for (int c = 0; c < 7; ++c) fprintf(stderr, " %s %.1e", nm[c], worst[c]);
      double dmaxn = 0; for (int k = 0; k <= N; ++k) { for (int j = 0; j < 3; ++j) dmaxn = fmax(dmaxn, fabs(d->z[k][j])); }
      fprintf(stderr, " |dz| %.1e dT %.1e a_max %.1e a_z %.1e Dphi %.2e\n", dmaxn, d->T, a_max, a_z, Dphi);
    }
    if (!F.active) {
      F.thmax = 1e4 * fmax(1.0, th); F.thmin = 1e-4 * fmax(1.0, th);
      F.active = 1; F.n = 0; F.wr = 0;
    }
    double a_min;
    if (Dphi < 0 && th <= F.thmin)
      a_min = fmin(g_th, fmin(g_ph * th / (-Dphi), (th > 0) ? pow(th, s_th) / pow(-Dphi, s_ph) : g_th));
    else if (Dphi < 0)
      a_min = fmin(g_th, g_ph * th / (-Dphi));
    else
      a_min = g_th;
    a_min *= 0.05;
    double a = a_max;
    int accepted = 0, restored = 0;
    /* acceptance of a trial (tht, pht) against a reference point (th_r, ph_r, Dphi_r) reached with step a_r:
     * 0 rejected, 1 sufficient decrease (h-type: the reference goes into the filter), 2 Armijo (f-type) */
#define ACCEPT_TEST(res, tht, pht, th_r, ph_r, dphi_r, a_r)                                                   \
    do { (res) = 0;                                                                                           \
      if (isfinite(pht) && (tht) < F.thmax) {                                                                 \
        int dominated_ = 0;                                                                                   \
        for (int i_ = 0; i_ < F.n; ++i_)                                                                      \
          if ((tht) >= F.th[i_] && (pht) >= F.ph[i_]) { dominated_ = 1; break; }                              \
        if (!dominated_) {                                                                                    \
          int sw_ = ((dphi_r) < 0) && ((a_r) * pow(-(dphi_r), s_ph) > pow((th_r), s_th));                     \
          if ((th_r) <= F.thmin && sw_) {                                                                     \
            if ((pht) <= (ph_r) + eta_ph * (a_r) * (dphi_r) + 10 * 2.220446049250313e-16 * fabs(ph_r)) (res) = 2; \
          } else if ((tht) <= (1 - g_th) * (th_r) || (pht) <= (ph_r) - g_ph * (th_r))                         \
            (res) = 1;                                                                                        \
        }                                                                                                     \
      } } while (0)
    int first = 1;
    while (a >= a_min * (1 - 1e-12)) {
      apply_step(p, it, d, a, &w->tr);
      eval_values(p, &w->tr, &w->vt);
      double tht, pht;
      theta_phi(p, &w->tr, &w->vt, mu, &tht, &pht, 0, 0);
      if (in_wd) {
        /* watchdog: only full steps, judged against the reference iterate */
        ACCEPT_TEST(accepted, tht, pht, wd_th, wd_ph, wd_dphi, wd_alpha);
        if (accepted) {
          in_wd = 0;
          if (accepted == 1) { /* the reference point enters the filter */
            int slot = (F.n < FILT_MAX) ? F.n++ : (F.wr % FILT_MAX);
            F.th[slot] = (1 - g_th) * wd_th; F.ph[slot] = wd_ph - g_ph * wd_th;
            F.wr++;
          }
          accepted = 3;
        } else if (++wd_count >= WD_MAX) {
          *it = w->wd; in_wd = 0; wd_block = 1; restored = 1;   /* give up: back to the reference iterate */
        } else
          accepted = 3;                                         /* one more full step on trust */
        break;
      }
      ACCEPT_TEST(accepted, tht, pht, th, ph0, Dphi, a);
      if (accepted) break;
      if (first && !wd_block && n_short >= WD_TRIGGER && WD_MAX > 0 && isfinite(pht)) {
        w->wd = *it; wd_th = th; wd_ph = ph0; wd_dphi = Dphi; wd_alpha = a;
        in_wd = 1; wd_count = 0; accepted = 3;
        break;
      }
      first = 0;
      a *= 0.5;
    }
#undef ACCEPT_TEST
    if (restored) { iter++; continue; }
    if (g_trace) g_trace(iter, p->resto ? th_orig : v->f, th, E0, mu, dw, accepted ? a : -1.0);
    /* IPOPT returns Solved_To_Acceptable_Level when it cannot make progress from a point that meets the
     * acceptable tolerance; near a degenerate vertex of the OBCA dual polytope the step noise floor is
     * above tol, so this is how such instances end
     * (second clause: barrier parameter at most 1e-6, primal feasible to 1e-6 and complementary, with only
     * the dual infeasibility sitting on the rounding-noise floor of the degenerate-vertex linear algebra) */
    int at_floor = (E0 <= P->acceptable_tol) ? OBCA_ST_ACCEPTABLE : (mu <= 1e-6 && th <= 1e-6 && fmax(E0, e1) <= 1e3 * P->acceptable_tol) ? OBCA_ST_FLOOR : 0;
    if (p->resto) at_floor = 0;
    if (!accepted) { status = at_floor ? at_floor : OBCA_ST_LSFAIL; break; }
    nstall = (a < stall_alpha) ? nstall + 1 : 0;
    if (nstall >= stall_iters) { status = at_floor ? at_floor : OBCA_ST_STALL; break; }
    if (accepted != 3) { wd_block = 0; n_short = (a < a_max) ? n_short + 1 : 0; }
    else if (!in_wd) n_short = 0;   /* watchdog succeeded */
    if (accepted == 1) {
      int slot = (F.n < FILT_MAX) ? F.n++ : (F.wr % FILT_MAX);
      F.th[slot] = (1 - g_th) * th; F.ph[slot] = ph0 - g_ph * th;
      F.wr++;
    }
    /* update: primal + slacks with a, equality multipliers towards y+ with a, Z with a_z */
    {
#define UPD_Z(Sarr, Zarr, dSarr, Snew)                                         \
      do { double S_ = (Sarr), Z_ = (Zarr), dS_ = (dSarr);                     \
        double dZ_ = mu / S_ - Z_ - (Z_ / S_) * dS_;                            \
        double Zn_ = Z_ + a_z * dZ_, Sn_ = (Snew);                              \
        Zn_ = fmin(fmax(Zn_, mu / (kappa_sigma * Sn_)), kappa_sigma * mu / Sn_); \
        (Zarr) = Zn_; } while (0)
      iter_t* tr = &w->tr; /* holds X + a dX, S + a dS of the accepted trial */
      if (p->resto) { /* multipliers of n >= 0 first: their step needs the row's current Z */
#define UPD_V(cls, Sx, Zx, dSx, Vx, nnew)                                                              \
        do { if (RELAXED(p, cls)) { double dZ_ = mu / (Sx) - (Zx) - ((Zx) / (Sx)) * (dSx);             \
          double Vn_ = (Vx) + a_z * ((p->rho - (Zx) - (Vx)) - dZ_);                                    \
          (Vx) = fmin(fmax(Vn_, mu / (kappa_sigma * (nnew))), kappa_sigma * mu / (nnew)); } } while (0)
        for (int k = 0; k <= N; ++k) {
          if (k >= 1) for (int j = 0; j < 4; ++j) UPD_V(CLS_XY, it->Sxy[k][j], it->Zxy[k][j], d->Sxy[k][j], it->Vxy[k][j], tr->nxy[k][j]);
          if (k < N) for (int j = 0; j < 8; ++j) UPD_V(CLS_UB, it->Sub[k][j], it->Zub[k][j], d->Sub[k][j], it->Vub[k][j], tr->nub[k][j]);
          for (int i = 0; i < p->nobs; ++i) UPD_V(CLS_NORM, it->Sn[k][i], it->Zn[k][i], d->Sn[k][i], it->Vn[k][i], tr->nn[k][i]);
          for (int i = 0; i < p->nobs; ++i) UPD_V(CLS_DIST, it->Sd[k][i], it->Zd[k][i], d->Sd[k][i], it->Vd[k][i], tr->nd[k][i]);
        }
        if (p->free_) for (int j = 0; j < 2; ++j) UPD_V(CLS_TB, it->STb[j], it->ZTb[j], d->STb[j], it->VTb[j], tr->nTb[j]);
        if (p->has_term) for (int j = 0; j < 3; ++j) UPD_V(CLS_TM, it->Stm[j], it->Ztm[j], d->Stm[j], it->Vtm[j], tr->ntm[j]);
#undef UPD_V
        if (p->free_) for (int j = 0; j < 3; ++j) {
          double dy = d->yt[j] - it->yt[j], y = it->yt[j];
          double Vp = it->Vpt[j] + a_z * ((p->rho - y - it->Vpt[j]) - dy), Vn = it->Vnt[j] + a_z * ((p->rho + y - it->Vnt[j]) + dy);
          it->Vpt[j] = fmin(fmax(Vp, mu / (kappa_sigma * tr->pt[j])), kappa_sigma * mu / tr->pt[j]);
          it->Vnt[j] = fmin(fmax(Vn, mu / (kappa_sigma * tr->nt[j])), kappa_sigma * mu / tr->nt[j]);
        }
      }
      for (int k = 0; k <= N; ++k) {
        if (k >= 1) for (int j = 0; j < 4; ++j) UPD_Z(it->Sxy[k][j], it->Zxy[k][j], d->Sxy[k][j], tr->Sxy[k][j]);
        if (k < N) for (int j = 0; j < 8; ++j) UPD_Z(it->Sub[k][j], it->Zub[k][j], d->Sub[k][j], tr->Sub[k][j]);
        for (int r = 0; r < p->R; ++r) UPD_Z(it->Sl[k][r], it->Zl[k][r], d->Sl[k][r], tr->Sl[k][r]);
        for (int r = 0; r < 4 * p->nobs; ++r) UPD_Z(it->Sm[k][r], it->Zm[k][r], d->Sm[k][r], tr->Sm[k][r]);
        for (int i = 0; i < p->nobs; ++i) UPD_Z(it->Sn[k][i], it->Zn[k][i], d->Sn[k][i], tr->Sn[k][i]);
        for (int i = 0; i < p->nobs; ++i) UPD_Z(it->Sd[k][i], it->Zd[k][i], d->Sd[k][i], tr->Sd[k][i]);
        if (k < N) for (int j = 0; j < 3; ++j) it->yd[k][j] += a * (d->yd[k][j] - it->yd[k][j]);
        for (int j = 0; j < 2 * p->nobs; ++j) it->ye[k][j] += a * (d->ye[k][j] - it->ye[k][j]);
      }
      if (p->free_) {
        for (int j = 0; j < 2; ++j) UPD_Z(it->STb[j], it->ZTb[j], d->STb[j], tr->STb[j]);
        for (int j = 0; j < 3; ++j) it->yt[j] += a * (d->yt[j] - it->yt[j]);
      }
      if (p->has_term) for (int j = 0; j < 3; ++j) UPD_Z(it->Stm[j], it->Ztm[j], d->Stm[j], tr->Stm[j]);
      apply_step(p, it, d, a, it);
    }
    iter++;
  }
  if (status < 0 && best_E0 < 1e300) { *it = w->best; status = best_lvl; E0 = best_E0; }
  else if (status < 0 && in_wd) *it = w->wd;   /* failed on a step taken on trust: the point reached is the watchdog's reference */
  *iters_out = iter;
  if (err_out) *err_out = E0;
  if (mu_out) *mu_out = mu;
  return status;
}

static int failed_search(int st) { return st == OBCA_ST_LSFAIL || st == OBCA_ST_REGFAIL || st == OBCA_ST_STALL; }
/* outcomes after which another start point may still succeed (the problems are non-convex) */
static int failed_attempt(int st) { return failed_search(st) || st == OBCA_ST_INFEASIBLE || st == OBCA_ST_RESTOFAIL; }

/* One attempt = the interior-point pass and, where its line search / regularisation / progress fails, IPOPT's remedy:
 * the feasibility-restoration phase from the point reached (same algorithm on the restoration problem, see prob_t),
 * then the original problem again from the restored point with multipliers, slacks, barrier parameter and filter
 * afresh.  Where the restoration phase cannot reduce the violation (a local minimiser of the violation - e.g. a
 * predicted pose inside an obstacle, where the OBCA distance has no gradient - or its own line search fails) the fresh
 * start is taken from the point of failure instead.  At most g_max_resto rounds, within the iteration budget; every
 * call of the restoration phase has to get below kappa_resto times the lowest violation seen so far. */
static int solve_attempt(const prob_t* p, work_t* w, int* iters_io, int use_resto, int budget) {
  int it_a = 0;
  double mu_end = 0;
  int st = solve_one(p, w, &it_a, 0, &mu_end);
  *iters_io += it_a;
  if (g_verbose) fprintf(stderr, "  pass init=%d: status %d iters %d mu %.1e\n", p->init, st, it_a, mu_end);
  double th_goal = 1e300;
  int infeasible = 0;
  for (int nres = 0; use_resto && failed_search(st) && nres < g_max_resto && *iters_io < budget; ++nres) {
    prob_t pr = *p;
    eval_values(p, &w->it, &w->v);
    double th, ph, cmax, th_orig;
    theta_phi(p, &w->it, &w->v, mu_end, &th, &ph, &cmax, &th_orig);
    int st_r = OBCA_ST_OK;
    it_a = 0;
    /* called at an almost feasible point (IPOPT aborts there): nothing to restore, only the fresh start below */
    if (th_orig > g_feas_tol) {
      w->fail = w->it;
      pr.resto = 1; pr.rmask = g_rmask; pr.rho = g_rho;
      for (int k = 0; k <= p->N; ++k) {
        for (int j = 0; j < 3; ++j) pr.zR[k][j] = w->it.z[k][j];
        for (int j = 0; j < 2; ++j) pr.uR[k][j] = w->it.u[k][j];
      }
      pr.TR = w->it.T;
      th_goal = fmin(th_goal, th_orig);
      pr.th_ref = th_goal;
      pr.mu0 = fmax(mu_end, cmax);
      pr.zeta = sqrt(pr.mu0);
      st_r = solve_one(&pr, w, &it_a, 0, 0);
      *iters_io += it_a;
      if (g_verbose) {
        eval_values(p, &w->it, &w->v);
        double th2, to2;
        theta_phi(p, &w->it, &w->v, mu_end, &th2, &ph, &cmax, &to2);
        fprintf(stderr, "  restoration: th_ref %.3e mu0 %.1e -> status %d iters %d th_orig %.3e\n", th_orig, pr.mu0, st_r, it_a, to2);
      }
      if (st_r < 0) w->it = w->fail; else th_goal *= g_kappa_resto;
      infeasible = (st_r == OBCA_ST_INFEASIBLE);
    }
    prob_t pk = *p;
    pk.init = OBCA_INIT_KEEP;
    st = solve_one(&pk, w, &it_a, 0, &mu_end);
    *iters_io += it_a;
    if (g_verbose) fprintf(stderr, "  pass keep: status %d iters %d mu %.1e\n", st, it_a, mu_end);
  }
  if (failed_search(st) && infeasible) st = OBCA_ST_INFEASIBLE;
  return st;
}

/* ------------------------------------------------------------------------------------------------
 * batch driver with the same argument list as obca_b200_solve (host pointers)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const obca_params* P;
  int batch, t0, t1;
  const double *x0, *u0, *xref, *uref, *T_max, *term, *Ts_inst, *A, *b0, *db;
  const int32_t* edge_ptr;
  int shared;
  double *x, *u, *lam, *mu, *T, *obj;
  int32_t *status, *iters;
} job_t;

static void* worker(void* arg) {
  job_t* J = (job_t*)arg;
  const obca_params* P = J->P;
  int N = P->N, R = P->rows, no = P->n_obs;
  work_t* w = (work_t*)malloc(sizeof(work_t));
  for (int b = J->t0; b < J->t1; ++b) {
    prob_t p;
    memset(&p, 0, sizeof(p));
    p.P = P; p.N = N; p.nobs = no; p.R = R;
    p.free_ = (P->mode == OBCA_MODE_FREE || P->mode == OBCA_MODE_FREE_STACKED);
    p.has_term = (P->mode == OBCA_MODE_FIXED_SET) || (P->mode == OBCA_MODE_FIXED_OBCA2 && P->has_term);
    p.stacked = (P->mode != OBCA_MODE_FREE);
    for (int i = 0; i <= no; ++i) p.eptr[i] = J->edge_ptr[i];
    p.Ts = J->Ts_inst ? J->Ts_inst[b] : P->Ts; p.dmin = P->dmin;
    double L = P->ego[0] + P->ego[2], W = P->ego[1] + P->ego[3];
    p.g[0] = L / 2; p.g[1] = W / 2; p.g[2] = L / 2; p.g[3] = W / 2;
    p.off = L / 2 - P->ego[2];
    for (int j = 0; j < 3; ++j) p.x0[j] = J->x0[3 * b + j];
    for (int j = 0; j < 2; ++j) p.u0[j] = J->u0[2 * b + j];
    p.xref = J->xref + (size_t)b * 3 * (N + 1);
    p.uref = J->uref ? J->uref + (size_t)b * 2 * N : 0;
    p.Tmax = (p.free_ && J->T_max) ? J->T_max[b] : 1.0;
    if (p.has_term) for (int j = 0; j < 3; ++j) p.term[j] = J->term[3 * b + j];
    size_t ob = J->shared ? 0 : (size_t)b;
    p.A = J->A + ob * 2 * R; p.b0 = J->b0 + ob * R; p.db = J->db ? J->db + ob * R : 0;
    /* recovery sequence (include/obca_b200.h): per start point up to n soft restarts from the point reached
     * (OBCA_INIT_SOFT), then the next start point (OBCA_INIT_RETRY) */
    static const int order[4][4] = {{0, 2, 1, -1}, {1, 2, 0, -1}, {2, 1, 0, -1}, {OBCA_INIT_GUESS, 2, 1, 0}};
    const int has_guess = (P->init & 15) == OBCA_INIT_GUESS;
    const int base = has_guess ? 3 : (P->init & 15) % 3, retry = (P->init & OBCA_INIT_RETRY) != 0, nsoft = OBCA_SOFT_RESTARTS(P->init);
    double guess[NS * 3];
    if (has_guess) { memcpy(guess, J->x + (size_t)b * 3 * (N + 1), sizeof(double) * 3 * (N + 1)); p.guess = guess; }
    int iters = 0, st = OBCA_ST_MAXITER;
    const int use_resto = (P->init & OBCA_INIT_NORESTO) == 0;
    int budget = g_budget;
    for (int a = 0; a < 4 && order[base][a] >= 0; ++a) {
      if (P->init & OBCA_INIT_PATIENT) budget = iters + g_budget;   /* the iteration budget counts per start point */
      for (int s_ = 0; s_ <= nsoft; ++s_) {
        p.init = s_ == 0 ? order[base][a] : OBCA_INIT_KEEP;
        st = solve_attempt(&p, w, &iters, use_resto, budget);
        if (!failed_attempt(st) || iters >= budget) break;
      }
      if (!retry || !failed_attempt(st) || iters >= budget) break;
    }
    const iter_t* it = &w->it;
    for (int k = 0; k <= N; ++k) {
      for (int j = 0; j < 3; ++j) J->x[((size_t)b * (N + 1) + k) * 3 + j] = it->z[k][j];
      if (k < N) for (int j = 0; j < 2; ++j) J->u[((size_t)b * N + k) * 2 + j] = it->u[k][j];
      for (int r = 0; r < R; ++r) J->lam[((size_t)b * (N + 1) + k) * R + r] = it->lam[k][r];
      for (int r = 0; r < 4 * no; ++r) J->mu[((size_t)b * (N + 1) + k) * 4 * no + r] = it->mu[k][r];
    }
    J->T[b] = p.free_ ? it->T : 1.0;
    J->obj[b] = objective_of(&p, it);
    J->status[b] = st;
    J->iters[b] = iters;
  }
  free(w);
  return 0;
}

int obca_oracle_solve(const obca_params* P, int batch, const double* x0, const double* u0, const double* xref,
                      const double* uref, const double* T_max, const double* term, const double* Ts_inst,
                      const int32_t* edge_ptr,
                      const double* A, const double* b0, const double* db, int obstacles_shared, double* x, double* u,
                      double* lam, double* mu, double* T, double* obj, int32_t* status, int32_t* iters, int nthreads) {
  if (!P || P->N + 1 > NS || P->N < 1 || P->rows > RM || P->n_obs > OM) return OBCA_E_SIZE;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > batch) nthreads = batch > 0 ? batch : 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    job_t* J = &jobs[t];
    J->P = P; J->batch = batch;
    J->t0 = (int)((long long)batch * t / nthreads); J->t1 = (int)((long long)batch * (t + 1) / nthreads);
    J->x0 = x0; J->u0 = u0; J->xref = xref; J->uref = uref; J->T_max = T_max; J->term = term; J->Ts_inst = Ts_inst;
    J->A = A; J->b0 = b0; J->db = db; J->edge_ptr = edge_ptr; J->shared = obstacles_shared;
    J->x = x; J->u = u; J->lam = lam; J->mu = mu; J->T = T; J->obj = obj; J->status = status; J->iters = iters;
    if (nthreads == 1) worker(J);
    else pthread_create(&th[t], 0, worker, J);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
  return OBCA_OK;
}

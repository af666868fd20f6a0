"""ORACLE (test infrastructure): dense primal-dual interior-point solver for the compact OBCA NLP.

This is the *specification* of the solver algorithm that the C oracle (``oracle/obca_oracle.c``) and the
CUDA kernel implement with structured linear algebra.  Here the KKT system is assembled densely and
factorised with LAPACK's Bunch-Kaufman LDL^T (``scipy.linalg.ldl``), the inertia is read from D - no
structure is exploited, so it is an independent check of the Riccati/Schur elimination used elsewhere.

Formulation (slack form, every inequality gets a slack as ``Opti`` hands it to IPOPT):
    min f(X)  s.t.  c(X) = 0,  d(X) - S = 0,  S >= 0          duals: y (eq), Z (ineq)
Algorithm (IPOPT-flavoured, Waechter & Biegler 2006):
  Newton system on the barrier problem, slacks and Z eliminated (Sigma = Z/S):
      [W + Jd^T Sigma Jd + dw*Mw   J^T    ] [dX]   = - [grad f + J^T y - Jd^T (mu/S - Sigma (d - S))]
      [J                           -dc*Mc ] [dy]       [c]
      dS = Jd dX + (d - S),   dZ = mu/S - Z - Sigma dS
  inertia correction: dw in {0, 1e-4, ...x8 (x100 first time)} on the trajectory variables (z,u,T) until the
  matrix has inertia (n, m, 0); dc on the terminal-equality rows only (the only rows that can lose rank),
  chosen Levenberg-Marquardt style so that the terminal multiplier step stays bounded;
  fraction-to-boundary tau = max(0.99, 1-mu) on S and Z; filter line search with second-order correction;
  watchdog (IPOPT: watchdog_shortened_iter_trigger 10, watchdog_trial_iter_max 3): after 10 consecutive shortened
  steps a rejected full step is taken on trust from a saved reference iterate; if 3 more full steps reach no point
  acceptable to the reference (filter + Armijo with the reference's values), the reference is restored and ordinary
  backtracking resumes there.  (The optional second-order correction, soc=True, is not part of the specification the
  C oracle and the CUDA kernel implement.)
  monotone barrier update mu <- max(tol/10, min(0.2 mu, mu^1.5)) when E_mu <= 10 mu;
  stop when IPOPT's scaled optimality error E_0 <= tol (1e-8); the best iterate at the acceptable level (E_0 <=
  acceptable_tol, or - at mu <= 1e-6 - theta <= 1e-6 and scaled and unscaled error <= 1e3 acceptable_tol) is stored and becomes the result ("Solved To
  Acceptable Level") if the run later ends in a failure or stalls there (error not halved for 10 iterations).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from . import obca_nlp as nlp

DEFAULT_OPTS = dict(tol=1e-8, max_iter=3000, mu_init=10.0, kappa_eps=10.0, kappa_mu=0.2, theta_mu=1.5,
                    tau_min=0.99, bound_push=0.1, s_max=100.0, kappa_sigma=1e10,
                    dw_first=1e-4, dw_min=1e-20, dw_max=1e20, kw_plus_first=100.0, kw_plus=8.0, kw_minus=1.0 / 3.0,
                    dc_min=1e-8, lm_cap=1e4, acceptable_tol=1e-6, acceptable_iter=15, filt_max=32,
                    stall_alpha=1e-3, stall_iters=10, sig_min=1e-8,
                    wd_trigger=10, wd_max=3, acc_stall=10,
                    soft_restarts=0, retry=False, X0=None,
                    init="warm", verbose=False, soc=False, dbg=False)

# status codes (shared with oracle/obca_oracle.c and the CUDA kernel)
ST_OK, ST_ACCEPTABLE, ST_MAXITER, ST_REGFAIL, ST_EMPTYBOX, ST_LSFAIL, ST_STALL = 0, 1, -1, -2, -3, -4, -5


def _inertia(K):
    lu, d, perm = sla.ldl(K, lower=True)
    n = K.shape[0]
    pos = neg = zero = 0
    i = 0
    while i < n:
        if i + 1 < n and d[i + 1, i] != 0.0:
            ev = np.linalg.eigvalsh(d[i:i + 2, i:i + 2])
            for e in ev:
                if e > 0: pos += 1
                elif e < 0: neg += 1
                else: zero += 1
            i += 2
        else:
            e = d[i, i]
            if e > 0: pos += 1
            elif e < 0: neg += 1
            else: zero += 1
            i += 1
    return pos, neg, zero


def gname(lay, p, r):
    if r < lay.g_u: return "xy%d.%d" % ((r - lay.g_xy) // 4 + 1, (r - lay.g_xy) % 4)
    if r < lay.g_u + 8 * p.N: return "u%d.%d" % ((r - lay.g_u) // 8, (r - lay.g_u) % 8)
    if p.free and r < lay.g_T + 2: return "T.%d" % (r - lay.g_T)
    if p.term is not None and r < lay.g_term + 3: return "term.%d" % (r - lay.g_term)
    for k in range(p.N + 1):
        for i in range(p.nobs):
            E = int(p.edges[i])
            if lay.g_w[k, i] <= r < lay.g_w[k, i] + E + 6:
                j = r - lay.g_w[k, i]
                return "w%d,%d.%s" % (k, i, ("lam%d" % j) if j < E else ("mu%d" % (j - E)) if j < E + 4 else ("norm" if j == E + 4 else "dist"))
    return "?"


RECOVERY_BUDGET = 300      # OBCA_RECOVERY_BUDGET: no further pass once the passes add up to this many iterations
_RETRY_ORDER = {"zero": ("zero", "warm", "xref"), "xref": ("xref", "warm", "zero"), "warm": ("warm", "xref", "zero")}


def solve(p: nlp.Problem, opts=None):
    """Recovery sequence (OBCA_INIT_SOFT / OBCA_INIT_RETRY of include/obca_b200.h): an attempt that ends with a failed
    line search, regularisation or stall is followed by up to ``soft_restarts`` restarts from the point it reached
    (multipliers, slacks, barrier parameter and filter afresh), then - with ``retry`` - by the other start points;
    ``iters`` is the total."""
    o = dict(DEFAULT_OPTS)
    if opts:
        o.update(opts)
    total = 0
    res = None
    for init in (_RETRY_ORDER[o["init"]] if o["retry"] else (o["init"],)):
        for s_ in range(o["soft_restarts"] + 1):
            res = _solve_once(p, dict(o, init=init) if s_ == 0 else dict(o, init="keep", X0=res["X"]))
            total += res["iters"]
            if res["status"] not in (ST_LSFAIL, ST_REGFAIL, ST_STALL) or total >= RECOVERY_BUDGET:
                break
        if res["status"] not in (ST_LSFAIL, ST_REGFAIL, ST_STALL) or total >= RECOVERY_BUDGET:
            break
    res["iters"] = total
    return res


class Model:
    """What the interior-point loop needs from an NLP: layout, evaluation, start point."""

    def __init__(self, p: nlp.Problem):
        self.p = p
        self.lay = nlp.Layout(p)

    def evaluate(self, X, y=None, zi=None, want=("f", "g", "c", "J", "d", "Jd", "W")):
        return nlp.evaluate(self.p, self.lay, X, y, zi, want=want)

    def start(self, init):
        return nlp.start_point(self.p, self.lay, init)


def _solve_once(p: nlp.Problem, opts=None, model=None):
    o = dict(DEFAULT_OPTS)
    if opts:
        o.update(opts)
    model = Model(p) if model is None else model
    lay = model.lay
    n, m, q = lay.n, lay.m, lay.q
    res = dict(status=-1, iters=0, lay=lay)
    Mw = np.zeros(n); Mw[:lay.ntraj] = 1.0
    sign_rows, sign_cols = nlp.sign_rows(p, lay)

    X = np.array(o["X0"], float) if o["init"] == "keep" else model.start(o["init"])
    ev = model.evaluate(X, want=("c", "d"))
    S = np.maximum(ev["d"], o["bound_push"])
    Z = np.ones(q)
    y = np.zeros(m)
    mu = o["mu_init"]
    filt = None
    nfilt_wr = 0
    nstall = 0
    dw_last = 0.0
    tol = o["tol"]
    acc_count = 0
    best = None
    e_min, e_min_iter = 1e300, 0
    status = ST_MAXITER
    in_wd = False; wd_count = 0; wd_block = False; n_short = 0
    wd_ref = None; wd_state = None

    def err(gr, Jm, Jd, c, d, S, y, Z, mu_t):
        sd = max(o["s_max"], (np.abs(y).sum() + Z.sum()) / max(1, m + q)) / o["s_max"]
        sc = max(o["s_max"], Z.sum() / max(1, q)) / o["s_max"]
        e1 = np.abs(gr + Jm.T @ y - Jd.T @ Z).max() / sd
        e2 = max(np.abs(c).max() if m else 0.0, np.abs(d - S).max())
        e3 = np.abs(S * Z - mu_t).max() / sc
        return max(e1, e2, e3), (e1, e2, e3, e1 * sd)

    def phi_theta(X, S, mu):
        e = model.evaluate(X, want=("f", "c", "d"))
        return e["f"] - mu * np.log(S).sum(), np.abs(e["c"]).sum() + np.abs(e["d"] - S).sum(), e["c"], e["d"]

    it = 0
    hist = []
    while True:
        ev = model.evaluate(X, y, Z)
        f, gr, c, Jm, d, Jd, W = ev["f"], ev["g"], ev["c"], ev["J"], ev["d"], ev["Jd"], ev["W"]
        E0, parts = err(gr, Jm, Jd, c, d, S, y, Z, 0.0)
        th = np.abs(c).sum() + np.abs(d - S).sum()
        if o["verbose"]:
            print("it %3d f %.8e th %.2e E0 %.2e (%.1e %.1e %.1e) mu %.1e dw %.1e" % (it, f, th, E0, *parts[:3], mu, dw_last))
        hist.append((f, th, E0, mu))
        if E0 <= tol:
            status = 0
            break
        if E0 <= o["acceptable_tol"]:
            acc_count += 1
            if acc_count >= o["acceptable_iter"]:
                status = 1
                break
        else:
            acc_count = 0
        # stall at the acceptable level (final barrier parameter, error not halved for acc_stall iterations): end with
        # the stored point, as IPOPT does when it cannot progress from an acceptable point
        if best is not None and mu <= tol / 10 * (1 + 1e-12) and it - e_min_iter >= o["acc_stall"]:
            status = ST_LSFAIL
            break
        if it >= o["max_iter"]:
            status = ST_MAXITER
            break
        # barrier update
        changed = False
        while True:
            Emu, _ = err(gr, Jm, Jd, c, d, S, y, Z, mu)
            if Emu <= o["kappa_eps"] * mu and mu > tol / 10:
                mu = max(tol / 10, min(o["kappa_mu"] * mu, mu ** o["theta_mu"]))
                changed = True
            else:
                break
        # acceptable level, judged after the barrier update: the iteration that lowers mu to its final value counts
        # (second clause, the rounding-noise floor: reported under its own status by the C oracle and the kernel; the
        # dual infeasibility is taken UNSCALED as well - diverging multipliers make IPOPT's scaled error small at points
        # that are not stationary - and the level follows the caller's acceptable_tol: 1e-3 for mpc4, 1e-5 for mpc6/8)
        floor_lvl = mu <= 1e-6 and th <= 1e-6 and max(E0, parts[3]) <= 1e3 * o["acceptable_tol"]
        acc_lvl = E0 <= o["acceptable_tol"] or floor_lvl
        if acc_lvl:
            if best is None or E0 < 0.1 * best[0]:  # IPOPT stores the acceptable point (here: a new copy per decade) ...
                best = (E0, X.copy(), S.copy(), y.copy(), Z.copy())
            if E0 < 0.5 * e_min:
                e_min, e_min_iter = E0, it
        if changed and filt is not None:
            filt = []
            nfilt_wr = 0
        if changed:
            in_wd = False
        tau = max(o["tau_min"], 1 - mu)

        Sig = Z / S
        rd = d - S
        rhs_x = -(gr + Jm.T @ y - Jd.T @ (mu / S - Sig * rd))
        H0 = W + Jd.T @ (Sig[:, None] * Jd)
        # primal regularisation of the OBCA duals: curvature of a sign row is max(Z/S, sig_min)
        H0[sign_cols, sign_cols] += np.maximum(o["sig_min"] - Sig[sign_rows], 0.0)
        # terminal-equality regularisation (Levenberg-Marquardt: keeps the terminal multiplier step bounded
        # when the linearised dynamics cannot reach the terminal pose, e.g. theta = v = 0)
        Mc = np.zeros(m)
        dc = 0.0
        if p.free:
            dc = max(o["dc_min"], np.abs(c[lay.c_term:lay.c_term + 3]).max() / o["lm_cap"])
            Mc[lay.c_term:lay.c_term + 3] = dc
        # inertia correction (IPOPT alg. IC)
        dw = 0.0
        ntry = 0
        while True:
            K = np.zeros((n + m, n + m))
            K[:n, :n] = H0 + np.diag(dw * Mw)
            K[:n, n:] = Jm.T
            K[n:, :n] = Jm
            K[n:, n:] = -np.diag(Mc)
            pos, neg, zero = _inertia(K)
            if pos == n and neg == m:
                break
            ntry += 1
            if dw == 0.0:
                dw = o["dw_first"] if dw_last == 0.0 else max(o["dw_min"], o["kw_minus"] * dw_last)
            else:
                dw = dw * (o["kw_plus_first"] if dw_last == 0.0 else o["kw_plus"])
            if dw > o["dw_max"]:
                status = ST_REGFAIL
                break
        if status == ST_REGFAIL:
            break
        if dw > 0:
            dw_last = dw
        lu = sla.lu_factor(K)
        lm = dict(dc=dc)

        def solve_kkt(rx, rc):
            r0 = np.concatenate([rx, rc])
            s0 = sla.lu_solve(lu, r0); s0 += sla.lu_solve(lu, r0 - K @ s0)
            return s0[:n], s0[n:]

        dX, dy = solve_kkt(rhs_x, -c)
        dS = Jd @ dX + rd
        dZ = mu / S - Z - Sig * dS

        def ftb(v, dv):
            r = np.where(dv < 0, -tau * v / np.where(dv < 0, dv, -1.0), np.inf)
            return min(1.0, r.min()) if r.size else 1.0
        a_max = ftb(S, dS)
        a_z = ftb(Z, dZ)

        if o["dbg"]:
            r = np.where(dS < 0, -S / np.where(dS < 0, dS, -1.0), np.inf)
            il = np.argsort(r)[:5]
            print("      limit rows:", [(gname(lay, p, int(i)), "S=%.2e dS=%.2e d=%.2e" % (S[i], dS[i], d[i])) for i in il])
        # filter line search (IPOPT Alg. A) with second-order correction
        Dphi = gr @ dX - mu * (dS / S).sum()
        ph0 = f - mu * np.log(S).sum()
        if filt is None:
            th_max = 1e4 * max(1.0, th); th_min = 1e-4 * max(1.0, th)
            filt = []
        g_th, g_ph, s_th, s_ph, eta_ph = 1e-5, 1e-8, 1.1, 2.3, 1e-8
        if Dphi < 0 and th <= th_min:
            a_min = min(g_th, g_ph * th / (-Dphi), th ** s_th / (-Dphi) ** s_ph if th > 0 else g_th)
        elif Dphi < 0:
            a_min = min(g_th, g_ph * th / (-Dphi))
        else:
            a_min = g_th
        a_min *= 0.05

        def acceptable(tht, pht, a, th_r=None, ph_r=None, dphi_r=None):
            """0 rejected, 1 sufficient decrease w.r.t. the reference (current point by default), 2 Armijo"""
            th_r = th if th_r is None else th_r
            ph_r = ph0 if ph_r is None else ph_r
            dphi_r = Dphi if dphi_r is None else dphi_r
            if not np.isfinite(pht) or tht >= th_max:
                return 0
            for (tf, pf) in filt:
                if tht >= tf and pht >= pf:
                    return 0
            sw = dphi_r < 0 and a * (-dphi_r) ** s_ph > th_r ** s_th
            if th_r <= th_min and sw:
                return 2 if pht <= ph_r + eta_ph * a * dphi_r + 10 * np.finfo(float).eps * abs(ph_r) else 0
            if tht <= (1 - g_th) * th_r or pht <= ph_r - g_ph * th_r:
                return 1
            return 0

        def filt_add(th_e, ph_e):
            nonlocal nfilt_wr
            if len(filt) < o["filt_max"]:
                filt.append(((1 - g_th) * th_e, ph_e - g_ph * th_e))
            else:
                filt[nfilt_wr % o["filt_max"]] = ((1 - g_th) * th_e, ph_e - g_ph * th_e)
            nfilt_wr += 1
        a = a_max
        accepted = 0
        nbt = 0
        dXa, dSa = dX, dS
        nsoc = 0
        restored = False
        while a >= a_min * (1 - 1e-12):
            pht, tht, ct, dt = phi_theta(X + a * dX, S + a * dS, mu)
            if in_wd:
                accepted = acceptable(tht, pht, wd_ref[3], wd_ref[0], wd_ref[1], wd_ref[2])
                if accepted:
                    in_wd = False
                    if accepted == 1:
                        filt_add(wd_ref[0], wd_ref[1])
                    accepted = 3
                else:
                    wd_count += 1
                    if wd_count >= o["wd_max"]:
                        X, S, y, Z = [v.copy() for v in wd_state]
                        in_wd = False; wd_block = True; restored = True
                    else:
                        accepted = 3
                break
            accepted = acceptable(tht, pht, a)
            if accepted:
                break
            if nbt == 0 and not wd_block and n_short >= o["wd_trigger"] and o["wd_max"] > 0 and np.isfinite(pht):
                wd_state = (X.copy(), S.copy(), y.copy(), Z.copy())
                wd_ref = (th, ph0, Dphi, a)
                in_wd = True; wd_count = 0; accepted = 3
                break
            if nbt == 0 and tht >= th and o["soc"]:
                csoc = a * c + ct
                dsoc = a * rd + (dt - (S + a * dS))
                th_old = th
                for ps in range(4):
                    nsoc += 1
                    rx2 = -(gr + Jm.T @ y - Jd.T @ (mu / S - Sig * dsoc))
                    dXs, _ = solve_kkt(rx2, -csoc)
                    dSs = Jd @ dXs + dsoc
                    a_s = ftb(S, dSs)
                    phs, ths, cs, ds = phi_theta(X + a_s * dXs, S + a_s * dSs, mu)
                    accepted = acceptable(ths, phs, a_s)
                    if accepted:
                        dXa, dSa = dXs, dSs; a = a_s
                        break
                    if ths > 0.99 * th_old:
                        break
                    th_old = ths
                    csoc = a_s * csoc + cs
                    dsoc = a_s * dsoc + (ds - (S + a_s * dSs))
                if accepted:
                    break
            a *= 0.5
            nbt += 1
        if restored:
            it += 1
            continue
        at_floor = E0 <= o["acceptable_tol"] or (mu <= 1e-6 and th <= 1e-6 and max(E0, parts[3]) <= 1e3 * o["acceptable_tol"])
        if not accepted:
            status = ST_ACCEPTABLE if at_floor else ST_LSFAIL
            break
        nstall = nstall + 1 if a < o["stall_alpha"] else 0
        if nstall >= o["stall_iters"]:
            status = ST_ACCEPTABLE if at_floor else ST_STALL
            break
        if accepted != 3:
            wd_block = False
            n_short = n_short + 1 if a < a_max else 0
        elif not in_wd:
            n_short = 0
        if accepted == 1:
            filt_add(th, ph0)
        if o["verbose"]:
            print("      a_max %.2e a %.2e a_z %.2e |dX| %.2e |dy| %.2e acc %d bt %d soc %d ntry %d nf %d dc %.1e" % (
                a_max, a, a_z, np.abs(dX).max(), np.abs(dy).max(), accepted, nbt, nsoc, ntry, len(filt), lm.get("dc", 0)))
        X = X + a * dXa
        S = S + a * dSa
        y = y + a * dy
        Z = Z + a_z * dZ
        ks = o["kappa_sigma"]
        Z = np.clip(Z, mu / (ks * S), ks * mu / S)
        it += 1

    if status < 0 and best is not None:            # ... and ends there when the run fails later on
        E0, X, S, y, Z = best
        status = ST_ACCEPTABLE
    x, u, T, lam, mu_d = nlp.unpack(p, lay, X)
    res.update(status=status, iters=it, X=X, S=S, y=y, Z=Z, x=x, u=u, T=T, lam=lam, mu_dual=mu_d,
               obj=nlp.objective_of(p, x, u, T), err=E0, mu=mu, hist=hist)
    return res

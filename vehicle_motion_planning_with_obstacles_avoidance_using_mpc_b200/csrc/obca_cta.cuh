// obca_cta.cuh - batched OBCA-MPC interior-point solver for sm_100a, one thread block per NLP instance.
//
// The NLP is the reference's (src/obca.py: obca_mpc4 828-1071, obca_mpc6 1361-1562, obca_mpc8 1564-1758,
// obca2 338-629) in the compact variable set of SURVEY.md Appendix A; the algorithm is the primal-dual
// interior-point method specified by oracle/ipm_dense.py.  Nothing here is shared with oracle/.
//
// Mapping (why: ncu on the first, warp-per-instance kernel showed 80 % of the issue slots lost to instruction
// fetch and the rest to L2 latency on a 73 KB per-warp workspace; see profiles/ and DESIGN.md):
//   * thread t < nb = n_obs*(N+1) owns the OBCA dual block of (obstacle i = t/(N+1), stage k = t%(N+1)):
//     lambda, mu, their slacks and multipliers live in REGISTERS for the whole solve
//   * the last warp is the "stage warp": lane k owns stage k's pose/input/bound state in SHARED memory, and
//     the 32 lanes share the entries of the 8x8 stage matrices during the (sequential) Riccati sweep
//   * everything an iteration touches is on-chip; HBM is read once (inputs) and written once (results)
// Per iteration:  block assemble (square-root 5x5 factorisation per thread, Schur complement onto the 3x3 pose
// block) || stage assemble -> combine -> Riccati (lanes = matrix entries) -> roll-out -> block/stage steps,
// fraction to the boundary -> filter line search -> update.  Control flow is uniform across the block; all
// decisions are taken on block-reduced scalars.
//
// The file compiles for the device (nvcc) and for the host (g++): tools/emu runs the very same phase code
// with the threads of a block executed one after the other, so the kernel logic is testable without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/obca_b200.h"

#if defined(__CUDACC__)
#define OB_HD __host__ __device__ __forceinline__
#else
#define OB_HD inline
#endif

namespace obca {

constexpr int FILT_MAX = 32;
constexpr double SIG_MIN = 1e-8;  // primal regularisation of the OBCA duals (curvature floor of a sign row)
constexpr int NPART = 12;            // per-thread partial results handed to the block reductions
constexpr int NPART_X = NPART + 1;   // + one slot that only the restoration pass uses (reduced separately)
constexpr int OBCA_ST_STORED = 100;  // internal: the result arrays already hold the (acceptable) answer
// feasibility-restoration phase (IPOPT's remedy for a failed line search; Waechter & Biegler 2006, sec. 3.3): see
// solve_with_recovery.  kappa: reduction of the violation a call has to reach; rho: l1 penalty (IPOPT: 1000)
constexpr double RESTO_KAPPA = 0.1, RESTO_RHO = 1000.0, RESTO_FEAS_TOL = 1e-6, RESTO_TOL = 1e-8;
constexpr int RESTO_ROUNDS = 2, RESTO_STALL = 15, RESTO_MAXITER = 200;

struct KParams {
  obca_params P;
  int32_t eptr[OBCA_MAX_OBS + 1];
  int32_t batch, shared_obs, free_, has_term, stacked;
  const double *x0, *u0, *xref, *uref, *Tmax, *term, *Ts_inst, *A, *b0, *db;
  double *x, *u, *lam, *mu, *T, *obj;
  int32_t *status, *iters;
  unsigned int* counter;  // persistent-block work queue
  const int32_t* index;     // work item w solves instance index[w] (NULL: w itself)
  const int32_t* count_dev; // number of work items read on the device (NULL: batch)
  double* wd_buf;         // checkpoints in HBM: two slots of wd_stride doubles per resident block
  int64_t wd_stride;
  unsigned long long* prof;   // phase cycle counters (48 words) of the -DOBCA_PROFILE build, else NULL
  // first-pass kernel: instances whose pass fails and that the flags allow to recover are appended here (the recovery
  // kernel then solves exactly this work list); NULL: report the failure
  int32_t* fail_list;
  unsigned int* fail_count;
  unsigned int* bulk_timeouts;   // diagnostics: bulk-copy prefetches that did not complete in time (plain loads took over)
  int64_t smem_stride;           // doubles of shared memory per instance (a block hosts several side by side)
  const unsigned int* poll;      // recovery kernel running beside the first pass: != NULL -> poll the list until *poll != 0
  const unsigned int* heartbeat; // ... and the producer's work counter (no movement for 4 ms: the producer is not running)
  int32_t wd_block0;             // first checkpoint slot of this launch (the concurrent recovery block has its own)
};

// block-uniform scalar state of one instance (shared memory)
struct Glob {
  double T, STb[2], ZTb[2], Stm[3], Ztm[3], yt[3];
  double dT, dSTb[2], dStm[3], dyt[3];
  double Tmax, x0[3], u0[2], term[3];
  double Ts, off, g[4];
  double fth[FILT_MAX], fph[FILT_MAX];
  // loop-carried control state that is read rarely: kept here instead of in every thread's registers (the iteration
  // body is register-bound).  Written by one thread, read by all after a block barrier.
  double c_best_E0, c_best_f, c_thmax, c_thmin, c_dw_last;
  double c_wd_th, c_wd_ph, c_wd_dphi, c_wd_alpha, c_wd_pw_th, c_wd_pw_dphi, c_wd_cmax;
  int bad;
  int init;   // start point of the current attempt (OBCA_INIT_*; set by load, changed by the retry rule)
  // what a pass leaves for the recovery sequence: violation (1-norm, max-norm) and barrier parameter at its last point
  double c_th_end, c_cmax_end, c_mu_end;
  // restoration pass only.  Relaxed rows  d + n - S = 0, n >= 0 (cost rho n, bound multiplier V): state box NXY/VXY
  // and OBCA distance ND/VD live in the shared arrays, the terminal set here; terminal equality z_N - r_N - pt + nt = 0
  double ntm[3], Vtm[3], dntm[3];
  double pt[3], nt[3], Vpt[3], Vnt[3], dpt[3], dnt[3];
  double tdc[3], tcta[3], tctb[3];   // eliminated terminal equality: dz_N - tdc dy = -(mu tcta + tctb)
  double TR, zeta, th_ref, mu0;      // reference time scale, proximity weight, violation to get below, first mu
};

// shared-memory map of one instance; stage arrays are [element][stage], block arrays [element][block]
struct Sm {
  int N, S1, no, R, nb, T, nwarps, has_uref;
  double *Z, *U, *YD, *SXY, *ZXY, *SUB, *ZUB;
  double *DZ, *DU, *DYD, *DSXY, *DSUB;
  double *H, *RA, *RB, *CD, *GL, *DYN;
  double *K, *KAP, *PM, *PV, *XI;
  double *ETA, *DLAM, *DMU, *DYE, *DSN, *DSD, *EX;
  double *A, *B0, *DB, *XREF, *UREF;
  double *NXY, *VXY, *DNXY, *ZR, *UR, *ND, *VD, *DND;   // restoration pass (ZR also carries the caller's guess, OBCA_INIT_GUESS)
  double *RIC, *RED, *SCR_D, *SCR_H;
  double* PF;   // staging buffer of the next instance's inputs (bulk-copy prefetch, obca_kernel.cuh); starts 16-byte aligned
  uint32_t* TAB;
  Glob* G;
  OB_HD double& st(double* p, int e, int k) const { return p[e * S1 + k]; }
  OB_HD double& bl(double* p, int e, int t) const { return p[e * nb + t]; }
};

// row stride of the block-reduction scratch: one value per thread, one pad word per 16 (bank spread of stage A)
OB_HD int red_stride(int T) { return T; }

constexpr int EX_N = 16;   // per block: G(6) Ga(3) Gb(3) gLz(3) h22(1)
constexpr int RIC_N = 96;   // F8(36) f(8) W(36) pc(6) of the Riccati step in flight
// task tables of the cooperative Riccati sweep (uint32 words, filled per instance by fill_tables).  One task = one
// output entry = a dot product of <= 4 terms.  Operands are addressed by 16-bit references into the block's shared
// memory: bits 0..14 = offset in doubles from the first array, bit 15 = "add the stage index s" (arrays laid out
// [element][stage]).  A term word = coefficient reference | operand reference << 16.  Every sub-step has <= 32 tasks,
// so the sweep runs on ONE warp with warp barriers only.
constexpr int TW = 0;              // 30 x 4  W = P At (24 entries: columns th, T, v, w) and pc = p - P c (6 entries)
constexpr int TF = TW + 30 * 4;    // 29 x 4  F = H + At^T W (21 entries), f = r + At^T pc (8 entries)
constexpr int TFH = TF + 29 * 4;   // 29      F entry index | dw class << 8 | (feedback slot + 1) << 12
constexpr int TB = TFH + 29;       // 27 x 3  elimination of (v, w): operand references (2 per word)
constexpr int TAB_N = TB + 27 * 3;
constexpr uint32_t REF_S = 0x8000u;   // reference flag: add the stage index

// Prefetch buffer: one slot per input array of an instance, each with room for the array plus one double of alignment
// shift, rounded to an even number of doubles so that every slot starts 16-byte aligned (bulk copies move 16-byte
// units: an array that starts on an odd double is shifted by one inside its slot).  Slot 0 holds the mbarrier.
enum { PF_MBAR = 0, PF_XREF, PF_UREF, PF_A, PF_B0, PF_DB, PF_X0, PF_U0, PF_TMAX, PF_TS, PF_TERM, PF_NSLOT };
OB_HD int pf_len(int slot, int N, int R, int has_uref) {
  const int n = slot == PF_MBAR ? 2 : slot == PF_XREF ? 3 * (N + 1) : slot == PF_UREF ? (has_uref ? 2 * N : 0) : slot == PF_A ? 2 * R :
                (slot == PF_B0 || slot == PF_DB) ? R : slot == PF_X0 ? 3 : slot == PF_U0 ? 2 : slot == PF_TERM ? 3 : 1;
  return n;
}
OB_HD int pf_off(int slot, int N, int R, int has_uref) {
  int o = 0;
  for (int j = 0; j < slot; ++j) o += (pf_len(j, N, R, has_uref) + 1 + 1) & ~1;
  return o;
}
OB_HD int pf_doubles(int N, int R, int has_uref) { return pf_off(PF_NSLOT, N, R, has_uref); }

OB_HD size_t sm_carve(Sm& s, double* base, int N, int no, int R, int nwarps, int has_uref) {
  const int S1 = N + 1, nb = no * S1;
  s.N = N; s.S1 = S1; s.no = no; s.R = R; s.nb = nb; s.nwarps = nwarps; s.T = 32 * nwarps; s.has_uref = has_uref;
  size_t o = 0;
  // (base may be null when only the size is wanted: the pointers are then meaningless but never dereferenced; no
  //  null test here - on the device it would put a select in front of every shared-memory access)
  auto take = [&](size_t n) { double* p = base + o; o += n; return p; };
  s.Z = take(3 * S1); s.U = take(2 * S1); s.YD = take(3 * S1);
  s.SXY = take(4 * S1); s.ZXY = take(4 * S1); s.SUB = take(8 * S1); s.ZUB = take(8 * S1);
  // step arrays first: together with the padding they double as the scratch of the 12-slot block reduction that
  // follows assemble (nothing of the previous step is alive then); same for H|RA and the 3-slot reductions
  const size_t d0 = o;
  s.DZ = take(3 * S1); s.DU = take(2 * S1); s.DYD = take(3 * S1); s.DSXY = take(4 * S1); s.DSUB = take(8 * S1);
  s.DLAM = take((size_t)R * S1); s.DMU = take(4 * (size_t)nb); s.DYE = take(2 * (size_t)nb);
  s.DSN = take(nb); s.DSD = take(nb); s.XI = take(6 * S1);
  if (o - d0 < (size_t)NPART * red_stride(s.T)) take((size_t)NPART * red_stride(s.T) - (o - d0));
  if (o - d0 < (size_t)EX_N * nb) take((size_t)EX_N * nb - (o - d0));
  if (o - d0 < (size_t)RIC_N) take((size_t)RIC_N - (o - d0));
  s.SCR_D = base + d0;
  s.EX = s.SCR_D;   // block -> stage exchange of assemble: consumed (combine) before the reduction scratch is written
  const size_t h0 = o;
  s.H = take(36 * S1); s.RA = take(8 * S1);   // H|RA (44 S1) is re-used by the roll-out as ACL(36)|CCL(6)
  if (o - h0 < (size_t)3 * red_stride(s.T)) take((size_t)3 * red_stride(s.T) - (o - h0));
  s.SCR_H = base + h0;
  s.RB = take(8 * S1);
  s.DYN = take(13 * S1); s.CD = s.DYN + 10 * S1;   // CD = elements 10..12 of DYN (one coefficient base)
  s.K = take(12 * S1); s.KAP = take(2 * S1); s.PM = take(21 * S1); s.PV = take(6 * S1);
  s.GL = s.K;       // Lagrangian gradient of assemble: dead before the sweep writes the feedback gains
  s.ETA = take(25 * (size_t)nb);
  s.A = take(2 * R); s.B0 = take(R); s.DB = take(R); s.XREF = take(3 * S1); s.UREF = take(has_uref ? 2 * N : 0);
  s.NXY = take(4 * S1); s.VXY = take(4 * S1); s.DNXY = take(4 * S1); s.ZR = take(3 * S1); s.UR = take(2 * S1);
  s.ND = take(nb); s.VD = take(nb); s.DND = take(nb);
  s.RIC = s.SCR_D;   // scratch of the sweep: the step arrays are dead while the sweep runs
  if (o & 1) take(1);
  s.PF = take(pf_doubles(N, R, has_uref));
  s.RED = take(NPART + 8 * NPART);
  s.TAB = (uint32_t*)take((TAB_N + 1) / 2);
  s.G = (Glob*)(base + o); o += (sizeof(Glob) + 7) / 8;
  return o;  // doubles
}

// registers of a block thread: the OBCA duals of one (stage, obstacle) pair
template <int EMAX>
struct BlockRegs {
  double lam[EMAX], Sl[EMAX], Zl[EMAX];
  double mu[4], Sm_[4], Zm[4];
  double ye[2], Sn, Zn, Sd, Zd;
  int i, k, r0, E;   // obstacle, stage, first row and edge count of this thread's block (set once per instance)
  double gf[8];      // stage lanes only: objective gradient of the stage, assemble -> directional derivative
};

// registers of a stage-warp lane during the Riccati sweep: the operands of its three tasks, decoded once per sweep from
// the task tables into BYTE offsets from the first shared array, and a word of flags.  Every lane runs every sub-step
// without a branch: idle lanes (and absent second destinations) read a per-stage zero and write a dummy scratch word,
// the two or three forms a sub-step has are selected by flag.  (Measured before this layout, cfg 3: 1,665 cycles per
// stage, a third of them in reconvergence after the `if (t < 21) ... else ...` bodies and in loads the compiler had
// serialised between the FMAs - profiles/r2_sweep_sass_before.txt.)
struct SweepRegs {
  uint32_t wa[4], wb[4], wp, wd;           // sub-step 1: coefficients (+s), operands (+s), base term (+s), destination
  uint32_t fa[4], fb[4], fx, fy, fd, fk;   // sub-step 2: coefficients (+s), operands (+s if SW_FS), x, y (+s), destination, feedback slot (+s if SW_FK)
  uint32_t bo[5], bd;                      // sub-step 3: operands (+s if bit SW_B0 + p), destination (+s if SW_BD)
  uint32_t ric, bad;                       // the sweep scratch, the "pivot not positive" flag of the instance
  uint32_t flags;
};
enum : uint32_t { SW_PC = 1u, SW_FS = 2u, SW_FF = 4u, SW_FK = 8u, SW_DW_ALL = 16u, SW_DW_GE1 = 32u, SW_DW_EQ0 = 64u,
                  SW_B0 = 256u, SW_BD = 1u << 13, SW_PV = 1u << 14 };
constexpr int RIC_DUMMY = 96;   // scratch words nobody reads, one per lane (RIC: 0..43 F|f, 44..67 W, 80..85 pc; the region it
                                // aliases holds at least 12 x 64 doubles, see sm_carve) - a word shared by the idle lanes would
                                // be a write-write hazard for racecheck

OB_HD void ob_sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}

// Branch-free reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / RSQ64H, ~2^-23) + two Newton steps
// in FMA arithmetic (~1 ulp).  No slow-path call, no divergence region: the IEEE division/sqrt sequences are what
// bloats the instruction stream of this kernel (see profiles/).  Arguments are positive normal numbers here.
OB_HD double ob_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}
OB_HD double ob_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-h * r, r, 0.5);
  return fma(r, e, r);
#else
  return 1.0 / sqrt(x);
#endif
}

// how the sub-steps reach shared memory: through the generic pointer of the first array (host emulation) or through a
// 32-bit shared-window address (the device's sweep function, which is a real call and would otherwise see a generic
// pointer and issue generic loads)
struct SwMemPtr {
  char* zb;
  OB_HD double ld(uint32_t o) const { return *(const double*)(zb + o); }
  OB_HD void st(uint32_t o, double v) const { *(double*)(zb + o) = v; }
  OB_HD void set(uint32_t o) const { *(int*)(zb + o) = 1; }
};
#if defined(__CUDACC__)
struct SwMemShared {
  uint32_t base;
  __device__ __forceinline__ double ld(uint32_t o) const { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + o)); return v; }
  __device__ __forceinline__ void st(uint32_t o, double v) const { asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + o), "d"(v) : "memory"); }
  __device__ __forceinline__ void set(uint32_t o) const { asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + o), "r"(1) : "memory"); }
};
#endif

// The three sub-steps of a stage of the Riccati sweep (Solver::fill_tables has the tasks).  Branch-free; all loads of a
// sub-step are issued before its arithmetic.
template <class Mem>
struct SweepOps {
  // sub-step 1:  W = P_{s+1} At (columns th, T, v, w),  pc = p_{s+1} - P_{s+1} c
  OB_HD static void w(const Mem& m, int s, const SweepRegs& r) {
    const uint32_t sb = 8u * (uint32_t)s;
    double a[4], b[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) { a[p] = m.ld(r.wa[p] + sb); b[p] = m.ld(r.wb[p] + sb); }
    const double pv = m.ld(r.wp + sb);
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < 4; ++p) acc = fma(a[p], b[p], acc);
    m.st(r.wd, (r.flags & SW_PC) ? pv - acc : acc);
  }
  // sub-step 2:  F = H + At^T W (+ dw on the regularised diagonal),  f = mu ra + rb + At^T pc
  OB_HD static void f(const Mem& m, int s, double mu, double dw, const SweepRegs& r) {
    const uint32_t sb = 8u * (uint32_t)s, fl = r.flags;
    const uint32_t sbo = (fl & SW_FS) ? sb : 0u, sbk = (fl & SW_FK) ? sb : 0u;
    double a[4], b[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) { a[p] = m.ld(r.fa[p] + sb); b[p] = m.ld(r.fb[p] + sbo); }
    const double x = m.ld(r.fx + sb), y = m.ld(r.fy + sb);
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < 4; ++p) acc = fma(a[p], b[p], acc);
    acc += (fl & SW_FF) ? fma(mu, x, y) : y;
    const bool reg = (fl & SW_DW_ALL) || ((fl & SW_DW_GE1) && s >= 1) || ((fl & SW_DW_EQ0) && s == 0);
    acc += reg ? dw : 0.0;
    m.st(r.fd, acc);
    m.st(r.fk + sbk, acc);
  }
  // sub-step 3: eliminate (v, w); cost-to-go of stage s (the feedback gains follow in fwd_prep, lane-parallel)
  OB_HD static void b(const Mem& m, int t, int s, const SweepRegs& r) {
    const uint32_t sb = 8u * (uint32_t)s, fl = r.flags;
    double v[5];
#pragma unroll
    for (int p = 0; p < 5; ++p) v[p] = m.ld(r.bo[p] + ((fl & (SW_B0 << p)) ? sb : 0u));
    const double q00 = m.ld(r.ric + 8u * 27u), q01 = m.ld(r.ric + 8u * 34u), q11 = m.ld(r.ric + 8u * 35u);   // (6,6) (7,6) (7,7)
    const double det = q00 * q11 - q01 * q01;
    if (t == 0 && (!(q00 > 0) || !(det > 0))) m.set(r.bad);
    const double idet = ob_rcp(det);
    const double i00 = q11 * idet, i01 = -q01 * idet, i11 = q00 * idet;
    // entries: v0 - (v1 X + v2 Y) with (X, Y) = Q^-1 (v3, v4);  p_s rows: the same with (v3, v4) = (f6, f7), i.e. (X, Y) = kap
    const double X = fma(i00, v[3], i01 * v[4]), Y = fma(i01, v[3], i11 * v[4]);
    const double pm = v[0] - fma(v[1], X, v[2] * Y);
    const double pv = fma(-v[2], Y, fma(-v[1], X, v[0]));
    m.st(r.bd + ((fl & SW_BD) ? sb : 0u), (fl & SW_PV) ? pv : pm);
  }
};
#if defined(__CUDACC__)
// The stages of the sweep for one lane of the stage warp, as a real call THROUGH A POINTER.  Inlined into the
// interior-point loop the sub-steps compete for registers with everything that is live across the sweep (~90 block-uniform
// scalars of the loop), and at the kernel's 168 registers ptxas then interleaves every load with the FMA that consumes it:
// four shared-memory latencies in a row per sub-step instead of one.  A direct call does not help (ptxas allocates the
// callee inside the caller's budget); through a pointer the call follows the ABI, the callee saves what it needs
// (38 registers, once per sweep, stage warp only) and issues the loads of a sub-step together.
// -DOBCA_SWEEP_DIRECT: the direct call (A/B).
static __device__ __noinline__ void sweep_lane_dev(uint32_t zbase, SweepRegs r, int t, int N, double mu, double dw) {
  const SwMemShared m{zbase};
  for (int s = N - 1; s >= 0; --s) {
    SweepOps<SwMemShared>::w(m, s, r); __syncwarp();
    SweepOps<SwMemShared>::f(m, s, mu, dw, r); __syncwarp();
    SweepOps<SwMemShared>::b(m, t, s, r); __syncwarp();
  }
}
typedef void (*sweep_fn_t)(uint32_t, SweepRegs, int, int, double, double);
#ifndef OBCA_SWEEP_DIRECT
static __constant__ sweep_fn_t c_sweep_fn = sweep_lane_dev;
#endif
#endif
// sum of logarithms as the logarithm of a product: mantissas are multiplied, exponents added (one log per thread
// and phase instead of one per inequality).  Non-positive arguments poison the mantissa with NaN.
struct LogAcc {
  double m = 1.0;
  int e = 0;
  OB_HD void add(double S) {
    long long b;
#if defined(__CUDA_ARCH__)
    b = __double_as_longlong(S);
#else
    memcpy(&b, &S, 8);
#endif
    e += (int)((b >> 52) & 0x7ff) - 1023;
    long long mb = (b & 0x800fffffffffffffLL) | 0x3ff0000000000000LL;
    double mm;
#if defined(__CUDA_ARCH__)
    mm = __longlong_as_double(mb);
#else
    memcpy(&mm, &mb, 8);
#endif
    m *= (S > 0.0) ? mm : (0.0 / 0.0);
    if (m > 1e200) { m *= 0x1p-512; e += 512; }
  }
  OB_HD double value() const { return log(m) + e * 0.6931471805599453; }
};

// 5x5 square-root factor R^T (lower triangular, packed) with Givens row insertion
struct Tri5 {
  double l[15];
  double id[5];   // 1 / diagonal (set by finish)
  OB_HD void zero() {
#pragma unroll
    for (int i = 0; i < 15; ++i) l[i] = 0.0;
  }
  OB_HD double& at(int r, int c) { return l[r * (r + 1) / 2 + c]; }
  OB_HD void insert(double row[5]) {
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      double a = at(c, c), b = row[c];
      if (b != 0.0) {
        const double n2 = a * a + b * b, ir = ob_rsqrt(n2), cs = a * ir, sn = b * ir;
        at(c, c) = n2 * ir;
#pragma unroll
        for (int q = c + 1; q < 5; ++q) {
          double u = at(q, c), w = row[q];
          at(q, c) = cs * u + sn * w;
          row[q] = -sn * u + cs * w;
        }
      }
    }
  }
  OB_HD bool finish() {
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      if (at(a, a) < 0) {
#pragma unroll
        for (int q = a; q < 5; ++q) at(q, a) = -at(q, a);
      }
      ok = ok && (at(a, a) > 0) && isfinite(at(a, a));
      id[a] = ob_rcp(at(a, a));
    }
    return ok;
  }
  OB_HD void solve(const double r[5], double x[5]) {  // (L L^T) x = r
    double t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double s = r[i];
#pragma unroll
      for (int q = 0; q < i; ++q) s -= at(i, q) * t[q];
      t[i] = s * id[i];
    }
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      double s = t[i];
#pragma unroll
      for (int q = i + 1; q < 5; ++q) s -= at(q, i) * x[q];
      x[i] = s * id[i];
    }
  }
};

struct BlkGeo {  // geometry of one (stage, obstacle) block at the current point
  double a1, a2, ct, st, tx, ty;
};

OB_HD int symi(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// statistics of one inequality (slack S, multiplier Z, value d); returns sigma and the mu-split of the
// step-form right-hand side  t - Z = mu * ta + tb
struct IneqAcc {
  double th = 0, cmax = 0, sumz = 0, szmax = 0, szmin = 1e300;
  LogAcc lg;
  OB_HD void add(double S, double Z, double d, double& sig, double& ta, double& tb) {
    double rd = d - S;
    ta = ob_rcp(S); sig = Z * ta; tb = -Z - sig * rd;
    th += fabs(rd); cmax = fmax(cmax, fabs(rd)); lg.add(S); sumz += Z;
    double sz = S * Z; szmax = fmax(szmax, sz); szmin = fmin(szmin, sz);
  }
  // restoration pass: row relaxed to  d + n - S = 0, n >= 0 with cost rho n and bound multiplier V (stationarity in n:
  // rho - Z - V = 0).  Eliminating dS, dn, dV from the Newton equations leaves  dZ = t - sigma (grad d . dX)  with
  //   sigma = 1 / (S/Z + n/V),   t = mu ta + tb = -sigma [(d + n - S) + (mu - n (rho - Z))/V - (mu - S Z)/Z]
  // which tends to the ordinary row as n -> 0.  e1: dual infeasibility of the n-row; tho: what the row adds to the
  // violation of the ORIGINAL problem beyond its own residual (|d - S| - |d + n - S|).
  double e1 = 0, tho = 0;
  OB_HD void add_relaxed(double S, double Z, double n, double V, double d, double rho, double& sig, double& ta, double& tb) {
    const double r = d + n - S, iZ = ob_rcp(Z), iV = ob_rcp(V);
    sig = ob_rcp(S * iZ + n * iV);
    ta = sig * (iZ - iV);
    tb = -sig * (r - n * (rho - Z) * iV + S);
    th += fabs(r); cmax = fmax(cmax, fabs(r)); lg.add(S); lg.add(n); sumz += Z + V;
    const double sz = S * Z, nv = n * V;
    szmax = fmax(szmax, fmax(sz, nv)); szmin = fmin(szmin, fmin(sz, nv));
    e1 = fmax(e1, fabs(rho - Z - V));
    tho += fabs(d - S) - fabs(r);
  }
};
// steps of slack and relaxation of a relaxed row from the step dZ of its multiplier
OB_HD void relaxed_steps(double S, double Z, double n, double V, double mu, double rho, double dZ, double& dS, double& dn) {
  dS = (mu - S * Z - S * dZ) / Z;
  const double dV = (rho - Z - V) - dZ;
  dn = (mu - n * V - n * dV) / V;
}
// weight of the proximity term: D_R^2 = min(1, 1/ref^2)
OB_HD double dr2(double ref) { const double a = fabs(ref); return a > 1.0 ? 1.0 / (a * a) : 1.0; }

// partial-result slots (sum / max / min groups are reduced separately)
enum { PS_F = 0, PS_TH, PS_LG, PS_SUMY, PS_SUMZ, PS_GT, PM_E1, PM_E2, PM_SZMAX, PM_CT, PM_BAD, PN_SZMIN, PS_THO };  // assemble (PS_THO: restoration)
enum { QS_DPHI = 0, QN_AMAX = 1, QN_AZ = 2 };                                                                 // backsub

struct StageVals {
  double f, cd[3], dxy[4], dub[8];
};

// ======================================================================================================
// The solver: all phase code.  `Exec` supplies the execution model (device block or host emulation).
// ======================================================================================================
template <int EMAX>
struct Solver {
  const KParams& kp;
  const obca_params& P;
  Sm sm;
  int N, S1, no, nb;
  bool free_, has_term, stacked;

  // The dimensions come from the shared-memory map, not from the parameter block: a kernel instantiated for a fixed
  // (N, n_obs, rows, warps) carves with literals, so every array offset and loop bound below folds to a constant
  // (on the generic kernel a quarter of the executed instructions were shared-memory address arithmetic).
  OB_HD Solver(const KParams& kp_, const Sm& sm_)
      : kp(kp_), P(kp_.P), sm(sm_), N(sm_.N), S1(sm_.S1), no(sm_.no), nb(sm_.nb),
        free_(kp_.free_ != 0), has_term(kp_.has_term != 0), stacked(kp_.stacked != 0) {}

  // ---- thread roles
  OB_HD bool is_block(int tid) const { return tid < nb; }
  OB_HD int stage_lane(int tid) const { return tid - 32 * (sm.nwarps - 1); }   // >= 0 in the stage warp
  OB_HD bool is_stage(int tid) const { int l = stage_lane(tid); return l >= 0 && l <= N; }

  OB_HD double bk(int k, int r) const { return sm.B0[r] + (stacked ? k * sm.DB[r] : 0.0); }
  OB_HD const double* xref(int k) const { return sm.XREF + 3 * k; }

  // row j of a dual block (lambda rows 0..E-1, then the four mu rows): variable, slack, multiplier.  The block state
  // lives in registers, so the (warp-uniform) row index is resolved by selects, which keeps the row loops ROLLED:
  // one copy of the Givens-insertion code instead of E+7 (instruction fetch, not arithmetic, limits this kernel)
  OB_HD static void row_regs(const BlockRegs<EMAX>& br, int j, int E, double& wv, double& S, double& Z) {
#if defined(__CUDA_ARCH__)
    // the per-thread state object is homed in local memory (see DevExec), so a dynamic index is one local load per
    // value - cheaper than resolving the row by selects over registers (which was 10 % of the executed instructions)
    if (j < E) { wv = br.lam[j]; S = br.Sl[j]; Z = br.Zl[j]; }
    else { const int q = j - E; wv = br.mu[q]; S = br.Sm_[q]; Z = br.Zm[q]; }
#else
    const int q = (j < E) ? j : EMAX + (j - E);
    wv = 0.0; S = 1.0; Z = 1.0;
    for (int c = 0; c < EMAX; ++c)
      if (c == q) { wv = br.lam[c]; S = br.Sl[c]; Z = br.Zl[c]; }
    for (int c = 0; c < 4; ++c)
      if (EMAX + c == q) { wv = br.mu[c]; S = br.Sm_[c]; Z = br.Zm[c]; }
#endif
  }
  OB_HD void row_y(const BlkGeo& b, int k, int r0, int E, int j, double yv[5]) const {
    const Glob& G = *sm.G;
    if (j < E) {
      int r = r0 + j;
      double A0 = sm.A[2 * r], A1 = sm.A[2 * r + 1];
      yv[0] = A0; yv[1] = A1;
      yv[2] = b.tx * A0 + b.ty * A1 - bk(k, r);
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
  OB_HD static void cn_inv(const BlkGeo& b, double ci0, double ci1, const double v[2], double o[2]) {
    double av = b.a1 * v[0] + b.a2 * v[1];
    o[0] = (v[0] - ci1 * b.a1 * av) * ci0;
    o[1] = (v[1] - ci1 * b.a2 * av) * ci0;
  }

  // ------------------------------------------------------------------------------------------------
  // instance load (all threads): inputs HBM -> shared, once
  // ------------------------------------------------------------------------------------------------
  // the input arrays of ONE instance (each pointer already at the instance; Tmax / Ts / term / uref / db may be null)
  struct InstPtrs { const double *xref, *uref, *A, *b0, *db, *x0, *u0, *Tmax, *Ts, *term; };
  OB_HD InstPtrs inst_ptrs(size_t b) const {
    const int R = sm.R;
    const size_t ob = kp.shared_obs ? 0 : b;
    InstPtrs q;
    q.xref = kp.xref + b * 3 * S1; q.uref = (sm.has_uref && kp.uref) ? kp.uref + b * 2 * N : nullptr;
    q.A = kp.A ? kp.A + ob * 2 * R : nullptr; q.b0 = kp.b0 ? kp.b0 + ob * R : nullptr; q.db = kp.db ? kp.db + ob * R : nullptr;
    q.x0 = kp.x0 + 3 * b; q.u0 = kp.u0 + 2 * b;
    q.Tmax = (free_ && kp.Tmax) ? kp.Tmax + b : nullptr; q.Ts = kp.Ts_inst ? kp.Ts_inst + b : nullptr;
    q.term = (has_term && kp.term) ? kp.term + 3 * b : nullptr;
    return q;
  }
  OB_HD void load(int tid, size_t b, bool load_obs) const { load_from(tid, b, inst_ptrs(b), load_obs); }
  // inputs of instance b -> shared memory, from global memory (inst_ptrs) or from the prefetch buffer the bulk copies
  // filled (obca_kernel.cuh)
  OB_HD void load_from(int tid, size_t b, const InstPtrs& q, bool load_obs) const {
    Glob& G = *sm.G;
    const int T = sm.T, R = sm.R;
    for (int i = tid; i < 3 * S1; i += T) sm.XREF[i] = q.xref[i];
    if ((P.init & 15) == OBCA_INIT_GUESS)   // the caller's poses, read from the output array before it is written
      for (int i = tid; i < 3 * S1; i += T) sm.ZR[(i % 3) * S1 + i / 3] = kp.x[b * 3 * S1 + i];
    if (sm.has_uref)
      for (int i = tid; i < 2 * N; i += T) sm.UREF[i] = q.uref[i];
    if (load_obs) {
      for (int i = tid; i < 2 * R; i += T) sm.A[i] = q.A[i];
      for (int i = tid; i < R; i += T) {
        sm.B0[i] = q.b0[i];
        sm.DB[i] = q.db ? q.db[i] : 0.0;
      }
    }
    if (tid == 0) {
      for (int j = 0; j < 3; ++j) G.x0[j] = q.x0[j];
      for (int j = 0; j < 2; ++j) G.u0[j] = q.u0[j];
      G.Tmax = q.Tmax ? *q.Tmax : 1.0;
      G.Ts = q.Ts ? *q.Ts : P.Ts;
      G.init = ((P.init & 15) == OBCA_INIT_GUESS) ? OBCA_INIT_GUESS : (P.init & 15) % 3;
      G.T = 1.0; G.dT = 0.0;
      for (int j = 0; j < 3; ++j) {
        G.term[j] = q.term ? q.term[j] : 0.0;
        G.Stm[j] = G.Ztm[j] = 1.0; G.dStm[j] = 0.0; G.yt[j] = 0.0; G.dyt[j] = 0.0;
      }
      G.STb[0] = G.STb[1] = G.ZTb[0] = G.ZTb[1] = 1.0; G.dSTb[0] = G.dSTb[1] = 0.0;
      const double Lc = P.ego[0] + P.ego[2], Wc = P.ego[1] + P.ego[3];
      G.g[0] = Lc / 2; G.g[1] = Wc / 2; G.g[2] = Lc / 2; G.g[3] = Wc / 2;
      G.off = Lc / 2 - P.ego[2];
      G.bad = 0;
    }
  }

  // pose of the start-point construction: x0 at stage 0, the reference window elsewhere
  OB_HD void pp_of(int k, double pp[3]) const {
    const Glob& G = *sm.G;
#pragma unroll
    for (int j = 0; j < 3; ++j) pp[j] = (k == 0) ? G.x0[j] : (G.init == OBCA_INIT_GUESS ? sm.st(sm.ZR, j, k) : xref(k)[j]);
  }
  OB_HD static bool warm_like(int init) { return init == OBCA_INIT_WARM || init == OBCA_INIT_GUESS; }

  // ------------------------------------------------------------------------------------------------
  // start point (oracle/obca_nlp.py start_point): init 0 = reference (all zero, T = 1: obca.py:856),
  // 1 = poses from xref, 2 = A* warm start.   Pass 1: poses + path length partial.
  // ------------------------------------------------------------------------------------------------
  OB_HD void start_a(int tid, double* part) const {
    part[0] = 0.0;
    if (!is_stage(tid)) return;
    const int k = stage_lane(tid);
    double pp[3];
    pp_of(k, pp);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (sm.G->init != OBCA_INIT_KEEP) sm.st(sm.Z, j, k) = (k == 0) ? pp[j] : (sm.G->init >= OBCA_INIT_XREF ? pp[j] : 0.0);
      sm.st(sm.YD, j, k) = 0.0;
    }
    if (warm_like(sm.G->init) && k < N) {
      double pn[3];
      pp_of(k + 1, pn);
      part[0] = sqrt((pn[0] - pp[0]) * (pn[0] - pp[0]) + (pn[1] - pp[1]) * (pn[1] - pp[1]));
    }
  }
  OB_HD double start_T(double len) const {
    const Glob& G = *sm.G;
    if (G.init == OBCA_INIT_KEEP) return free_ ? G.T : 1.0;
    if (!warm_like(G.init) || !free_) return 1.0;
    double T0 = len / (N * P.uU[0] * G.Ts);
    return fmin(fmax(T0, 1.0), fmax(G.Tmax, P.T_min));
  }
  // Pass 2: inputs (stage lanes) and duals (block threads)
  OB_HD void start_b(int tid, BlockRegs<EMAX>& br, double T0) const {
    Glob& G = *sm.G;
    if (is_stage(tid)) {
      const int k = stage_lane(tid);
      double u[2] = {0, 0};
      if (warm_like(sm.G->init) && k < N) {
        const double h = (free_ ? T0 : 1.0) * G.Ts;
        double pp[3], pn[3];
        pp_of(k, pp); pp_of(k + 1, pn);
        double dth = pn[2] - pp[2] + M_PI;
        dth = dth - 2 * M_PI * floor(dth / (2 * M_PI)) - M_PI;
        double fwd = cos(pp[2]) * (pn[0] - pp[0]) + sin(pp[2]) * (pn[1] - pp[1]);
        u[0] = fmin(fmax(fwd / h, P.uL[0]), P.uU[0]);
        u[1] = fmin(fmax(dth / h, P.uL[1]), P.uU[1]);
      }
      if (G.init != OBCA_INIT_KEEP) { sm.st(sm.U, 0, k) = u[0]; sm.st(sm.U, 1, k) = u[1]; }
      if (k == 0) { G.T = T0; G.yt[0] = G.yt[1] = G.yt[2] = 0.0; }
    }
    if (is_block(tid)) {
      const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
      br.ye[0] = br.ye[1] = 0.0;
      if (G.init == OBCA_INIT_KEEP) return;
#pragma unroll
      for (int j = 0; j < EMAX; ++j) br.lam[j] = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) br.mu[q] = 0.0;
      if (warm_like(sm.G->init)) {
        double pp[3];
        pp_of(k, pp);
        const double ct = cos(pp[2]), st = sin(pp[2]);
        const double tx = pp[0] + G.off * ct, ty = pp[1] + G.off * st;
        int jb = -1;
        double best = -1e300, nbst = 1;
#pragma unroll
        for (int j = 0; j < EMAX; ++j) {
          if (j < E) {
            const int r = r0 + j;
            const double A0 = sm.A[2 * r], A1 = sm.A[2 * r + 1];
            const double nr = sqrt(A0 * A0 + A1 * A1);
            const double sep = (A0 * tx + A1 * ty - bk(k, r)) / nr;
            if (sep > best) { best = sep; jb = j; nbst = nr; }
          }
        }
        if (jb >= 0) {
          const double l = 0.9 / nbst;
#pragma unroll
          for (int j = 0; j < EMAX; ++j)
            if (j == jb) br.lam[j] = l;
          const double a1 = sm.A[2 * (r0 + jb)] * l, a2 = sm.A[2 * (r0 + jb) + 1] * l;
          const double r1 = -(ct * a1 + st * a2), r2 = -(-st * a1 + ct * a2);
          br.mu[0] = fmax(r1, 0.0); br.mu[1] = fmax(r2, 0.0); br.mu[2] = fmax(-r1, 0.0); br.mu[3] = fmax(-r2, 0.0);
        }
      }
    }
  }

  // values of the stage's own (non-obstacle) constraints at a point
  OB_HD void stage_vals(int k, const double z[3], const double u[2], const double up[2], const double zn[3], double T,
                        StageVals& o) const {
    const Glob& G = *sm.G;
    const double h = T * G.Ts;
    const double* M = (k < N) ? P.Q : P.P;
    double e[3], f = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = z[j] - xref(k)[j];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) f += e[a] * M[3 * a + b] * e[b];
    if (k < N) {
      double uu[2] = {u[0], u[1]};
      if (sm.has_uref) { uu[0] -= sm.UREF[2 * k]; uu[1] -= sm.UREF[2 * k + 1]; }
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
      ob_sincos(z[2], &st, &ct);
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
    if (k == 0 && free_) f += (N + 1) * (P.time_cost[0] * T + P.time_cost[1] * T * T);
    o.f = f;
  }

  // stage k's pose / inputs and neighbours at step length a along the current direction (a = 0: the iterate)
  OB_HD void stage_point(int k, double a, double z[3], double u[2], double up[2], double zn[3]) const {
    const Glob& G = *sm.G;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      z[j] = sm.st(sm.Z, j, k) + ((a != 0.0 && k >= 1) ? a * sm.st(sm.DZ, j, k) : 0.0);
      zn[j] = (k < N) ? sm.st(sm.Z, j, k + 1) + ((a != 0.0) ? a * sm.st(sm.DZ, j, k + 1) : 0.0) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      u[j] = sm.st(sm.U, j, k) + ((a != 0.0) ? a * sm.st(sm.DU, j, k) : 0.0);
      up[j] = (k == 0) ? G.u0[j] : sm.st(sm.U, j, k - 1) + ((a != 0.0) ? a * sm.st(sm.DU, j, k - 1) : 0.0);
    }
  }

  // ------------------------------------------------------------------------------------------------
  // slack / multiplier initialisation: S = max(d(x0), bound_push), Z = 1
  // ------------------------------------------------------------------------------------------------
  OB_HD void init_slacks(int tid, BlockRegs<EMAX>& br) const {
    Glob& G = *sm.G;
    const double bp = P.bound_push;
    const double T = free_ ? G.T : 1.0;
    if (is_stage(tid)) {
      const int k = stage_lane(tid);
      double z[3], u[2], up[2], zn[3];
      stage_point(k, 0.0, z, u, up, zn);
      StageVals sv;
      stage_vals(k, z, u, up, zn, T, sv);
#pragma unroll
      for (int j = 0; j < 4; ++j) { sm.st(sm.SXY, j, k) = fmax(sv.dxy[j], bp); sm.st(sm.ZXY, j, k) = 1.0; }
#pragma unroll
      for (int j = 0; j < 8; ++j) { sm.st(sm.SUB, j, k) = (k < N) ? fmax(sv.dub[j], bp) : 1.0; sm.st(sm.ZUB, j, k) = 1.0; }
      if (k == 0 && free_) {
        G.STb[0] = fmax(T - P.T_min, bp); G.STb[1] = fmax(G.Tmax - T, bp);
        G.ZTb[0] = G.ZTb[1] = 1.0;
      }
      if (k == N && has_term) {
        G.Stm[0] = fmax(z[0] - G.term[0], bp); G.Stm[1] = fmax(z[1] - G.term[1], bp); G.Stm[2] = fmax(G.term[2] - z[1], bp);
        G.Ztm[0] = G.Ztm[1] = G.Ztm[2] = 1.0;
      }
    }
    if (is_block(tid)) {
      const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
      const double z0 = sm.st(sm.Z, 0, k), z1 = sm.st(sm.Z, 1, k), z2 = sm.st(sm.Z, 2, k);
      double st, ct;
      ob_sincos(z2, &st, &ct);
      const double tx = z0 + G.off * ct, ty = z1 + G.off * st;
      double a1 = 0, a2 = 0, bl = 0;
#pragma unroll
      for (int j = 0; j < EMAX; ++j) {
        if (j < E) {
          const int r = r0 + j;
          const double l = br.lam[j];
          a1 += sm.A[2 * r] * l; a2 += sm.A[2 * r + 1] * l; bl += bk(k, r) * l;
          br.Sl[j] = fmax(l, bp); br.Zl[j] = 1.0;
        } else { br.Sl[j] = 1.0; br.Zl[j] = 1.0; }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { br.Sm_[q] = fmax(br.mu[q], bp); br.Zm[q] = 1.0; }
      br.Sn = fmax(1.0 - a1 * a1 - a2 * a2, bp); br.Zn = 1.0;
      br.Sd = fmax(-(G.g[0] * br.mu[0] + G.g[1] * br.mu[1] + G.g[2] * br.mu[2] + G.g[3] * br.mu[3]) + tx * a1 + ty * a2 - bl - P.dmin, bp);
      br.Zd = 1.0;
    }
  }

  // ------------------------------------------------------------------------------------------------
  // restoration pass, start: the point where the ordinary pass failed becomes the reference of the proximity term;
  // it is then projected onto the rows the restoration problem keeps hard, so that the pass starts feasible for its
  // own constraints - poses by rolling the inputs out through the dynamics, OBCA duals pushed inside their sign
  // bounds, mu_1..4 so that the two OBCA equalities hold - and the relaxation variables absorb what remains.
  // ------------------------------------------------------------------------------------------------
  OB_HD void resto_ref(int tid) const {
    Glob& G = *sm.G;
    if (!is_stage(tid)) return;
    const int k = stage_lane(tid);
#pragma unroll
    for (int j = 0; j < 3; ++j) sm.st(sm.ZR, j, k) = sm.st(sm.Z, j, k);
#pragma unroll
    for (int j = 0; j < 2; ++j) sm.st(sm.UR, j, k) = sm.st(sm.U, j, k);
    if (k == 0) G.TR = free_ ? G.T : 1.0;
  }
  OB_HD void resto_rollout() const {   // one thread: sequential in the stage
    const Glob& G = *sm.G;
    const double h = (free_ ? G.T : 1.0) * G.Ts;
    for (int k = 0; k < N; ++k) {
      const double z0 = sm.st(sm.Z, 0, k), z1 = sm.st(sm.Z, 1, k), z2 = sm.st(sm.Z, 2, k);
      sm.st(sm.Z, 0, k + 1) = z0 + h * sm.st(sm.U, 0, k) * cos(z2);
      sm.st(sm.Z, 1, k + 1) = z1 + h * sm.st(sm.U, 0, k) * sin(z2);
      sm.st(sm.Z, 2, k + 1) = z2 + h * sm.st(sm.U, 1, k);
    }
  }
  // relaxation variable of a row with residual c = d - S: c - p + n = 0 with p n on the central path (IPOPT's
  // initialisation); the slack absorbs p (it carries no cost), so the row starts satisfied
  OB_HD static void relax_start(double d, double bp, double mu, double& S, double& Z, double& n, double& V) {
    S = fmax(d, bp); Z = 1.0;
    const double c = d - S, hh = (mu - RESTO_RHO * c) / (2 * RESTO_RHO);
    n = hh + sqrt(hh * hh + mu * c / (2 * RESTO_RHO));
    S += c + n;
    V = mu / n;
  }
  OB_HD void resto_init(int tid, BlockRegs<EMAX>& br, double mu) const {
    Glob& G = *sm.G;
    const double bp = P.bound_push;
    const double T = free_ ? G.T : 1.0;
    if (is_stage(tid)) {
      const int k = stage_lane(tid);
      double z[3], u[2], up[2], zn[3];
      stage_point(k, 0.0, z, u, up, zn);
      StageVals sv;
      stage_vals(k, z, u, up, zn, T, sv);
#pragma unroll
      for (int j = 0; j < 3; ++j) sm.st(sm.YD, j, k) = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k >= 1) relax_start(sv.dxy[j], bp, mu, sm.st(sm.SXY, j, k), sm.st(sm.ZXY, j, k), sm.st(sm.NXY, j, k), sm.st(sm.VXY, j, k));
        else { sm.st(sm.SXY, j, k) = fmax(sv.dxy[j], bp); sm.st(sm.ZXY, j, k) = 1.0; sm.st(sm.NXY, j, k) = 1.0; sm.st(sm.VXY, j, k) = 1.0; }
        sm.st(sm.DNXY, j, k) = 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { sm.st(sm.SUB, j, k) = (k < N) ? fmax(sv.dub[j], bp) : 1.0; sm.st(sm.ZUB, j, k) = 1.0; }
      if (k == 0 && free_) {
        G.STb[0] = fmax(T - P.T_min, bp); G.STb[1] = fmax(G.Tmax - T, bp);
        G.ZTb[0] = G.ZTb[1] = 1.0;
      }
      if (k == N) {
        const double dtm[3] = {z[0] - G.term[0], z[1] - G.term[1], G.term[2] - z[1]};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (has_term) relax_start(dtm[j], bp, mu, G.Stm[j], G.Ztm[j], G.ntm[j], G.Vtm[j]);
          G.dntm[j] = 0.0; G.yt[j] = 0.0; G.dpt[j] = G.dnt[j] = 0.0;
          if (free_) {
            const double c = z[j] - xref(N)[j], hh = (mu - RESTO_RHO * c) / (2 * RESTO_RHO);
            G.nt[j] = hh + sqrt(hh * hh + mu * c / (2 * RESTO_RHO));
            G.pt[j] = c + G.nt[j];
            G.Vpt[j] = mu / G.pt[j]; G.Vnt[j] = mu / G.nt[j];
          }
        }
      }
    }
    if (is_block(tid)) {
      const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
      const double z0 = sm.st(sm.Z, 0, k), z1 = sm.st(sm.Z, 1, k), z2 = sm.st(sm.Z, 2, k);
      double st, ct;
      ob_sincos(z2, &st, &ct);
      const double tx = z0 + G.off * ct, ty = z1 + G.off * st;
      double a1 = 0, a2 = 0, bl = 0;
#pragma unroll
      for (int j = 0; j < EMAX; ++j) {
        if (j < E) {
          const int r = r0 + j;
          const double l = fmax(br.lam[j], bp);
          br.lam[j] = l;
          a1 += sm.A[2 * r] * l; a2 += sm.A[2 * r + 1] * l; bl += bk(k, r) * l;
          br.Sl[j] = l; br.Zl[j] = 1.0;
        } else { br.Sl[j] = 1.0; br.Zl[j] = 1.0; }
      }
      const double c1 = ct * a1 + st * a2, c2 = -st * a1 + ct * a2;
      const double b1 = fmax(fmin(br.mu[0], br.mu[2]), bp), b2 = fmax(fmin(br.mu[1], br.mu[3]), bp);
      br.mu[0] = b1 + fmax(-c1, 0.0); br.mu[2] = b1 + fmax(c1, 0.0);
      br.mu[1] = b2 + fmax(-c2, 0.0); br.mu[3] = b2 + fmax(c2, 0.0);
#pragma unroll
      for (int q = 0; q < 4; ++q) { br.Sm_[q] = br.mu[q]; br.Zm[q] = 1.0; }
      br.ye[0] = br.ye[1] = 0.0;
      br.Sn = fmax(1.0 - a1 * a1 - a2 * a2, bp); br.Zn = 1.0;
      const double dd = -(G.g[0] * br.mu[0] + G.g[1] * br.mu[1] + G.g[2] * br.mu[2] + G.g[3] * br.mu[3]) + tx * a1 + ty * a2 - bl - P.dmin;
      relax_start(dd, bp, mu, br.Sd, br.Zd, sm.bl(sm.ND, 0, tid), sm.bl(sm.VD, 0, tid));
      sm.bl(sm.DND, 0, tid) = 0.0;
    }
  }

  // ------------------------------------------------------------------------------------------------
  // assemble, stage part (lane k): everything that does not depend on the barrier parameter; right-hand
  // sides are split as  mu * (a part) + (b part)  so that mu can be chosen from this pass's own error.
  // Leaves H, RA, RB (without the Lagrangian gradient), GL, GF, CD, DYN in shared memory.
  // ------------------------------------------------------------------------------------------------
  template <bool RESTO>
  OB_HD void assemble_stage(int k, BlockRegs<EMAX>& br, double* part) const {
    Glob& G = *sm.G;
    double z[3], u[2], up[2], zn[3], yd[3], ydm[3];
    stage_point(k, 0.0, z, u, up, zn);
#pragma unroll
    for (int j = 0; j < 3; ++j) { yd[j] = sm.st(sm.YD, j, k); ydm[j] = (k >= 1) ? sm.st(sm.YD, j, k - 1) : 0.0; }
    const double T = free_ ? G.T : 1.0, h = T * G.Ts;
    double st, ct;
    ob_sincos(z[2], &st, &ct);
    double H[36], ra[8], rb[8], gL[8], gf[8];
#pragma unroll
    for (int i = 0; i < 36; ++i) H[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ra[i] = rb[i] = gL[i] = gf[i] = 0; }
#define HH(a, b) H[((a) >= (b)) ? ((a) * ((a) + 1) / 2 + (b)) : ((b) * ((b) + 1) / 2 + (a))]
    StageVals sv;
    stage_vals(k, z, u, up, zn, T, sv);
    IneqAcc acc;
    double sumy = 0, ceq_th = 0, ceq_max = 0, ctmax = 0;
    // (1) tracking cost; restoration: proximity to the reference point instead of the objective
    double fR = 0.0;   // restoration objective of this stage: zeta/2 |D_R (. - ref)|^2 + rho (relaxation variables)
    if constexpr (RESTO) {
      if (k >= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double ref = sm.st(sm.ZR, a, k), w = G.zeta * dr2(ref), e = z[a] - ref;
          gf[a] += w * e; HH(a, a) += w; fR += 0.5 * w * e * e;
        }
      }
    } else {
      const double* M = (k < N) ? P.Q : P.P;
      double e[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) e[j] = z[j] - xref(k)[j];
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
    double dyn[10] = {0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0.0};
    if (k < N) {
      // (2) input cost
      double uu[2] = {u[0], u[1]};
      if (sm.has_uref) { uu[0] -= sm.UREF[2 * k]; uu[1] -= sm.UREF[2 * k + 1]; }
      if constexpr (RESTO) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const double ref = sm.st(sm.UR, a, k), w = G.zeta * dr2(ref), e = u[a] - ref;
          gf[6 + a] += w * e; HH(6 + a, 6 + a) += w; fR += 0.5 * w * e * e;
        }
      } else {
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
      }
      // (3) acceleration cost between u_{k-1} (state 3,4) and u_k, k >= 1 (the t == 0 term is identically 0)
      if (k >= 1 && !RESTO) {
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
      const double fT0 = free_ ? G.Ts * vv * ct : 0.0, fT1 = free_ ? G.Ts * vv * st : 0.0, fT2 = free_ ? G.Ts * ww : 0.0;
      dyn[0] = fth0; dyn[1] = fth1; dyn[2] = h * ct; dyn[3] = h * st; dyn[4] = h; dyn[5] = fT0; dyn[6] = fT1; dyn[7] = fT2;
      HH(2, 2) += h * vv * (-yd[0] * ct - yd[1] * st);
      HH(6, 2) += h * (-yd[0] * st + yd[1] * ct);
      if (free_) {
        HH(5, 2) += G.Ts * vv * (-yd[0] * st + yd[1] * ct);
        HH(6, 5) += G.Ts * (yd[0] * ct + yd[1] * st);
        HH(7, 5) += G.Ts * yd[2];
      }
      gL[0] += yd[0]; gL[1] += yd[1]; gL[2] += yd[2] + fth0 * yd[0] + fth1 * yd[1];
      gL[6] += h * ct * yd[0] + h * st * yd[1]; gL[7] += h * yd[2];
      gL[5] += fT0 * yd[0] + fT1 * yd[1] + fT2 * yd[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) { sumy += fabs(yd[j]); ceq_th += fabs(sv.cd[j]); ceq_max = fmax(ceq_max, fabs(sv.cd[j])); }
      // (7) input bounds and acceleration rows
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double sg, ta, tb;
        const double Z0 = sm.st(sm.ZUB, j, k), Z2 = sm.st(sm.ZUB, 2 + j, k);
        acc.add(sm.st(sm.SUB, j, k), Z0, sv.dub[j], sg, ta, tb);
        HH(6 + j, 6 + j) += sg; ra[6 + j] += ta; rb[6 + j] += tb; gL[6 + j] -= Z0;
        acc.add(sm.st(sm.SUB, 2 + j, k), Z2, sv.dub[2 + j], sg, ta, tb);
        HH(6 + j, 6 + j) += sg; ra[6 + j] -= ta; rb[6 + j] -= tb; gL[6 + j] += Z2;
        double ga = (up[j] - u[j]) / h;
        double s4, ta4, tb4, s6, ta6, tb6;
        const double Z4 = sm.st(sm.ZUB, 4 + j, k), Z6 = sm.st(sm.ZUB, 6 + j, k);
        acc.add(sm.st(sm.SUB, 4 + j, k), Z4, sv.dub[4 + j], s4, ta4, tb4);
        acc.add(sm.st(sm.SUB, 6 + j, k), Z6, sv.dub[6 + j], s6, ta6, tb6);
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
        const double Z0 = sm.st(sm.ZXY, j, k), Z2 = sm.st(sm.ZXY, 2 + j, k);
        if constexpr (RESTO) {
          acc.add_relaxed(sm.st(sm.SXY, j, k), Z0, sm.st(sm.NXY, j, k), sm.st(sm.VXY, j, k), sv.dxy[j], RESTO_RHO, sg, ta, tb);
          fR += RESTO_RHO * sm.st(sm.NXY, j, k);
        } else
          acc.add(sm.st(sm.SXY, j, k), Z0, sv.dxy[j], sg, ta, tb);
        HH(j, j) += sg; ra[j] += ta; rb[j] += tb; gL[j] -= Z0;
        if constexpr (RESTO) {
          acc.add_relaxed(sm.st(sm.SXY, 2 + j, k), Z2, sm.st(sm.NXY, 2 + j, k), sm.st(sm.VXY, 2 + j, k), sv.dxy[2 + j], RESTO_RHO, sg, ta, tb);
          fR += RESTO_RHO * sm.st(sm.NXY, 2 + j, k);
        } else
          acc.add(sm.st(sm.SXY, 2 + j, k), Z2, sv.dxy[2 + j], sg, ta, tb);
        HH(j, j) += sg; ra[j] -= ta; rb[j] -= tb; gL[j] += Z2;
      }
    }
    if (k == 0 && free_) {
      // (4) time cost and (8) T bounds live in stage 0
      if constexpr (RESTO) {
        const double w = G.zeta * dr2(G.TR), e = T - G.TR;
        gf[5] += w * e; HH(5, 5) += w; fR += 0.5 * w * e * e;
      } else {
        gf[5] += (N + 1) * (P.time_cost[0] + 2 * P.time_cost[1] * T);
        HH(5, 5) += 2 * (N + 1) * P.time_cost[1];
      }
      double sg, ta, tb;
      acc.add(G.STb[0], G.ZTb[0], T - P.T_min, sg, ta, tb);
      HH(5, 5) += sg; ra[5] += ta; rb[5] += tb; gL[5] -= G.ZTb[0];
      acc.add(G.STb[1], G.ZTb[1], G.Tmax - T, sg, ta, tb);
      HH(5, 5) += sg; ra[5] -= ta; rb[5] -= tb; gL[5] += G.ZTb[1];
    }
    if (k == N && has_term) {
      double sg, ta, tb;
      const double dtm[3] = {z[0] - G.term[0], z[1] - G.term[1], G.term[2] - z[1]};
      const int ix[3] = {0, 1, 1};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if constexpr (RESTO) {
          acc.add_relaxed(G.Stm[j], G.Ztm[j], G.ntm[j], G.Vtm[j], dtm[j], RESTO_RHO, sg, ta, tb);
          fR += RESTO_RHO * G.ntm[j];
        } else
          acc.add(G.Stm[j], G.Ztm[j], dtm[j], sg, ta, tb);
        HH(ix[j], ix[j]) += sg;
        if (j < 2) { ra[ix[j]] += ta; rb[ix[j]] += tb; gL[ix[j]] -= G.Ztm[j]; }
        else { ra[ix[j]] -= ta; rb[ix[j]] -= tb; gL[ix[j]] += G.Ztm[j]; }
      }
    }
    double tho_eq = 0.0;
    if (k == N && free_) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        gL[j] += G.yt[j];
        double c = z[j] - xref(N)[j];
        if constexpr (RESTO) {
          // terminal equality relaxed to  z_N - r_N - pt + nt = 0,  pt, nt >= 0 with cost rho each; eliminating pt, nt
          // and their multipliers leaves  dz_N - tdc dy = -(mu tcta + tctb)
          const double pp = G.pt[j], nn = G.nt[j], Vp = G.Vpt[j], Vn = G.Vnt[j], y = G.yt[j];
          tho_eq += fabs(c);
          c += -pp + nn;
          tho_eq -= fabs(c);
          const double iVp = ob_rcp(Vp), iVn = ob_rcp(Vn);
          G.tdc[j] = pp * iVp + nn * iVn;
          G.tcta[j] = iVn - iVp;
          G.tctb[j] = c + pp * (RESTO_RHO - y) * iVp - nn * (RESTO_RHO + y) * iVn;
          acc.lg.add(pp); acc.lg.add(nn); acc.sumz += Vp + Vn;
          const double sp = pp * Vp, sn_ = nn * Vn;
          acc.szmax = fmax(acc.szmax, fmax(sp, sn_)); acc.szmin = fmin(acc.szmin, fmin(sp, sn_));
          acc.e1 = fmax(acc.e1, fmax(fabs(RESTO_RHO - y - Vp), fabs(RESTO_RHO + y - Vn)));
          fR += RESTO_RHO * (pp + nn);
        }
        ceq_th += fabs(c); ceq_max = fmax(ceq_max, fabs(c)); ctmax = fmax(ctmax, fabs(c));
        sumy += fabs(G.yt[j]);
      }
    }
#undef HH
#pragma unroll
    for (int i = 0; i < 36; ++i) sm.st(sm.H, i, k) = H[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sm.st(sm.RA, i, k) = ra[i]; sm.st(sm.RB, i, k) = rb[i]; sm.st(sm.GL, i, k) = gL[i] + gf[i];
      br.gf[i] = gf[i]; sm.st(sm.DYN, i, k) = dyn[i];
    }
    sm.st(sm.DYN, 8, k) = dyn[8]; sm.st(sm.DYN, 9, k) = dyn[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) sm.st(sm.CD, j, k) = (k < N) ? sv.cd[j] : 0.0;
    part[PS_F] = RESTO ? fR : sv.f; part[PS_TH] = acc.th + ceq_th; part[PS_LG] = acc.lg.value(); part[PS_SUMY] = sumy; part[PS_SUMZ] = acc.sumz;
    part[PS_GT] = 0.0; part[PM_E1] = 0.0; part[PM_E2] = fmax(acc.cmax, ceq_max); part[PM_SZMAX] = acc.szmax;
    part[PM_CT] = ctmax; part[PM_BAD] = 0.0; part[PN_SZMIN] = acc.szmin;
    if constexpr (RESTO) { part[PS_THO] = acc.tho + tho_eq; part[PM_CT] = acc.e1; }   // PM_CT carries the n-rows' dual infeasibility
  }

  // ------------------------------------------------------------------------------------------------
  // assemble, block part (thread = one (obstacle, stage) pair): square-root factorisation of the 5x5 block
  // system, solves for the mu-split right-hand side and the three pose columns, Schur complement onto the pose
  // ------------------------------------------------------------------------------------------------
  template <bool RESTO>
  OB_HD void assemble_block(int tid, const BlockRegs<EMAX>& br, double* part) const {
    const Glob& G = *sm.G;
    const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
    const double z0 = sm.st(sm.Z, 0, k), z1 = sm.st(sm.Z, 1, k), z2 = sm.st(sm.Z, 2, k);
    BlkGeo b;
    ob_sincos(z2, &b.st, &b.ct);
    const double st = b.st, ct = b.ct;
    b.tx = z0 + G.off * ct; b.ty = z1 + G.off * st;
    double a1 = 0, a2 = 0, bl = 0;
#pragma unroll
    for (int j = 0; j < EMAX; ++j) {
      if (j < E) {
        const int r = r0 + j;
        a1 += sm.A[2 * r] * br.lam[j]; a2 += sm.A[2 * r + 1] * br.lam[j]; bl += bk(k, r) * br.lam[j];
      }
    }
    b.a1 = a1; b.a2 = a2;
    const double m0 = br.mu[0], m1 = br.mu[1], m2 = br.mu[2], m3 = br.mu[3];
    const double ce1 = m0 - m2 + ct * a1 + st * a2, ce2 = m1 - m3 - st * a1 + ct * a2;
    const double dn = 1.0 - a1 * a1 - a2 * a2;
    const double dd = -(G.g[0] * m0 + G.g[1] * m1 + G.g[2] * m2 + G.g[3] * m3) + b.tx * a1 + b.ty * a2 - bl - P.dmin;
    const double y1 = br.ye[0], y2 = br.ye[1];
    const double Sn = br.Sn, Zn = br.Zn, Sd = br.Sd, Zd = br.Zd;
    IneqAcc acc;
    double e1 = 0;
    const double ceq_th = fabs(ce1) + fabs(ce2), ceq_max = fmax(fabs(ce1), fabs(ce2));
    const double sumy = fabs(y1) + fabs(y2);
    double sn, tna, tnb, sd, tda, tdb;
    acc.add(Sn, Zn, dn, sn, tna, tnb);
    // right-hand side of the distance row divided by its sigma, split in mu:  (t - Z)/sd = mu rda + rdb
    double rda, rdb;
    if constexpr (RESTO) {
      acc.add_relaxed(Sd, Zd, sm.bl(sm.ND, 0, tid), sm.bl(sm.VD, 0, tid), dd, RESTO_RHO, sd, tda, tdb);
      const double isd = ob_rcp(sd);
      rda = tda * isd; rdb = tdb * isd;
    } else {
      acc.add(Sd, Zd, dd, sd, tda, tdb);
      rda = ob_rcp(Zd); rdb = -dd;
    }
    Tri5 Lf;
    Lf.zero();
    double g0a[5] = {0, 0, 0, 0, 0}, g0b[5] = {0, 0, 0, 0, 0};
    const double aa = a1 * a1 + a2 * a2, lam1 = 2 * Zn + 4 * sn * aa;
    const double ci0 = ob_rcp(2 * Zn), ci1 = 4 * sn * ob_rcp(lam1);
    const int nrow = E + 4;
#pragma unroll 1
    for (int j = 0; j < nrow + 3; ++j) {
      double yh[5];
      if (j < nrow) {
        double wv, S, Z, yv[5];
        row_regs(br, j, E, wv, S, Z);
        row_y(b, k, r0, E, j, yv);
        double sig, ta, tb;
        acc.add(S, Z, wv, sig, ta, tb);
        const double gl = y1 * yv[3] + y2 * yv[4] - Z + 2 * Zn * (a1 * yv[0] + a2 * yv[1]) - Zd * yv[2];
        e1 = fmax(e1, fabs(gl));
        tb -= gl;
        const double sq = ob_rsqrt(fmax(sig, SIG_MIN)), di = sq * sq;
#pragma unroll
        for (int a = 0; a < 5; ++a) { g0a[a] += yv[a] * ta * di; g0b[a] += yv[a] * tb * di; yh[a] = yv[a] * sq; }
      } else {
        // the three rows of C = blockdiag(Cn^-1, 1/sd, 0, 0) in square-root form
#pragma unroll
        for (int a = 0; a < 5; ++a) yh[a] = 0.0;
        if (j == nrow + 2) {
          yh[2] = ob_rsqrt(sd);
        } else if (aa > 0) {
          const double ina = ob_rsqrt(aa), e1_ = a1 * ina, e2_ = a2 * ina;
          if (j == nrow) { const double s1 = ob_rsqrt(lam1); yh[0] = s1 * e1_; yh[1] = s1 * e2_; }
          else { const double s2 = ob_rsqrt(2 * Zn); yh[0] = -s2 * e2_; yh[1] = s2 * e1_; }
        } else {
          const double s2 = ob_rsqrt(2 * Zn);
          if (j == nrow) yh[0] = s2; else yh[1] = s2;
        }
      }
      Lf.insert(yh);
    }
    const bool ok = Lf.finish();
    // right-hand sides: column 0 split in mu (h0 = -2 tn a), then the three pose columns
    double h0a[2] = {-2 * tna * a1, -2 * tna * a2}, h0b[2] = {-2 * tnb * a1, -2 * tnb * a2};
    double cha[2], chb[2];
    cn_inv(b, ci0, ci1, h0a, cha);
    cn_inv(b, ci0, ci1, h0b, chb);
    const double offt = G.off * (-st * a1 + ct * a2);
    const double dpose[3] = {a1, a2, offt};
    const double jt0 = -st * a1 + ct * a2, jt1 = -ct * a1 - st * a2;
    double gLz[3] = {-Zd * dpose[0], -Zd * dpose[1], y1 * jt0 + y2 * jt1 - Zd * dpose[2]};
    const double ydv = -Zd, c1 = y1 + ydv * G.off;
    const double hc[3][2] = {{ydv, 0.0}, {0.0, ydv}, {-c1 * st - y2 * ct, c1 * ct - y2 * st}};
    double bc[3][5];
    sm.bl(sm.ETA, 0, tid) = g0a[0] - cha[0]; sm.bl(sm.ETA, 1, tid) = g0a[1] - cha[1]; sm.bl(sm.ETA, 2, tid) = g0a[2] - rda;
    sm.bl(sm.ETA, 3, tid) = g0a[3]; sm.bl(sm.ETA, 4, tid) = g0a[4];
    sm.bl(sm.ETA, 5, tid) = g0b[0] - chb[0]; sm.bl(sm.ETA, 6, tid) = g0b[1] - chb[1]; sm.bl(sm.ETA, 7, tid) = g0b[2] - rdb;
    sm.bl(sm.ETA, 8, tid) = g0b[3] + ce1; sm.bl(sm.ETA, 9, tid) = g0b[4] + ce2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cn_inv(b, ci0, ci1, hc[c], bc[c]);
      bc[c][2] = dpose[c];
      bc[c][3] = (c == 2) ? jt0 : 0.0;
      bc[c][4] = (c == 2) ? jt1 : 0.0;
#pragma unroll
      for (int a = 0; a < 5; ++a) sm.bl(sm.ETA, 10 + 5 * c + a, tid) = bc[c][a];
    }
    // five solves with the same factor, rolled: right-hand sides are staged in (and overwritten by the solutions in)
    // this thread's ETA column
#pragma unroll 1
    for (int c = 0; c < 5; ++c) {
      double r[5], x[5];
#pragma unroll
      for (int a = 0; a < 5; ++a) r[a] = sm.bl(sm.ETA, 5 * c + a, tid);
      Lf.solve(r, x);
#pragma unroll
      for (int a = 0; a < 5; ++a) sm.bl(sm.ETA, 5 * c + a, tid) = x[a];
    }
    double eta0a[5], eta0b[5], etc[3][5];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      eta0a[a] = sm.bl(sm.ETA, a, tid); eta0b[a] = sm.bl(sm.ETA, 5 + a, tid);
#pragma unroll
      for (int c = 0; c < 3; ++c) etc[c][a] = sm.bl(sm.ETA, 10 + 5 * c + a, tid);
    }
    const double h22 = y1 * (-ct * a1 - st * a2) + y2 * (st * a1 - ct * a2) + ydv * G.off * (-ct * a1 - st * a2);
    int e = 0;
#pragma unroll
    for (int cp = 0; cp < 3; ++cp) {
#pragma unroll
      for (int c = 0; c <= cp; ++c) {
        double Gm = hc[cp][0] * bc[c][0] + hc[cp][1] * bc[c][1];
#pragma unroll
        for (int a = 0; a < 5; ++a) Gm -= bc[cp][a] * etc[c][a];
        sm.bl(sm.EX, e++, tid) = Gm;        // order (0,0) (1,0) (1,1) (2,0) (2,1) (2,2) = packed symmetric
      }
    }
#pragma unroll
    for (int cp = 0; cp < 3; ++cp) {
      double Ga = -(bc[cp][0] * h0a[0] + bc[cp][1] * h0a[1]), Gb = -(bc[cp][0] * h0b[0] + bc[cp][1] * h0b[1]);
#pragma unroll
      for (int a = 0; a < 5; ++a) { Ga -= bc[cp][a] * eta0a[a]; Gb -= bc[cp][a] * eta0b[a]; }
      sm.bl(sm.EX, 6 + cp, tid) = Ga; sm.bl(sm.EX, 9 + cp, tid) = Gb; sm.bl(sm.EX, 12 + cp, tid) = gLz[cp];
    }
    sm.bl(sm.EX, 15, tid) = h22;
    part[PS_F] = 0.0; part[PS_TH] = acc.th + ceq_th; part[PS_LG] = acc.lg.value(); part[PS_SUMY] = sumy; part[PS_SUMZ] = acc.sumz;
    part[PS_GT] = 0.0; part[PM_E1] = e1; part[PM_E2] = fmax(acc.cmax, ceq_max); part[PM_SZMAX] = acc.szmax;
    part[PM_CT] = 0.0; part[PM_BAD] = ok ? 0.0 : 1.0; part[PN_SZMIN] = acc.szmin;
    if constexpr (RESTO) { part[PS_F] = RESTO_RHO * sm.bl(sm.ND, 0, tid); part[PS_THO] = acc.tho; part[PM_CT] = acc.e1; }
  }

  // fold the block contributions into the stage QP (lane k), finish the step-form right-hand side and the
  // stage's share of the optimality error
  OB_HD void assemble_combine(int k, double* part) const {
    double Gs[EX_N];
#pragma unroll
    for (int e = 0; e < EX_N; ++e) Gs[e] = 0.0;
    for (int i = 0; i < no; ++i) {
      const int t = i * S1 + k;
#pragma unroll
      for (int e = 0; e < EX_N; ++e) Gs[e] += sm.bl(sm.EX, e, t);
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) sm.st(sm.H, e, k) -= Gs[e];     // packed (cp,c), cp,c < 3, is the head of H
    sm.st(sm.H, 5, k) += Gs[15];                                // H(2,2)
    double gL[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) gL[a] = sm.st(sm.GL, a, k);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      sm.st(sm.RA, a, k) += Gs[6 + a];
      sm.st(sm.RB, a, k) += Gs[9 + a];
      gL[a] += Gs[12 + a];
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) sm.st(sm.RB, a, k) -= gL[a];
    // stage share of the dual infeasibility; the (v_prev, w_prev) rows of stage k+1 belong to u_k
    // (rows 3, 4 of GL are final after assemble_stage, so the neighbour's can be read here)
    double e1 = 0;
    if (k >= 1) e1 = fmax(fabs(gL[0]), fmax(fabs(gL[1]), fabs(gL[2])));
    if (k < N) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double gun = (k + 1 < N) ? sm.st(sm.GL, 3 + j, k + 1) : 0.0;
        e1 = fmax(e1, fabs(gL[6 + j] + gun));
      }
    }
    part[PM_E1] = e1;
    part[PS_GT] = gL[5];
  }

  // ------------------------------------------------------------------------------------------------
  // Riccati sweep, cooperative: the 32 lanes of the stage warp share the entries of the 8x8 stage system.
  // Variables (x, y, th, v_prev, w_prev, T | v, w).  Sub-step A forms F = H + At^T P At and f = r + At^T pc;
  // sub-step B eliminates (v, w) and writes the cost-to-go of stage s.  Pivots double as the inertia test.
  // ------------------------------------------------------------------------------------------------
  // Column a of At = [A B] as <= 4 (row, coefficient) pairs; the coefficients are entries of the per-stage vector
  // DYN = (fth0, fth1, bv0, bv1, bw, fT0, fT1, fT2, 1, 0); columns 3, 4 (v_prev, w_prev) are empty.
  OB_HD static void at_col_tab(int a, int rows[4], int cfs[4]) {
    for (int p = 0; p < 4; ++p) { rows[p] = 0; cfs[p] = 9; }
    switch (a) {
      case 0: rows[0] = 0; cfs[0] = 8; break;
      case 1: rows[0] = 1; cfs[0] = 8; break;
      case 2: rows[0] = 0; cfs[0] = 0; rows[1] = 1; cfs[1] = 1; rows[2] = 2; cfs[2] = 8; break;
      case 5: rows[0] = 0; cfs[0] = 5; rows[1] = 1; cfs[1] = 6; rows[2] = 2; cfs[2] = 7; rows[3] = 5; cfs[3] = 8; break;
      case 6: rows[0] = 0; cfs[0] = 2; rows[1] = 1; cfs[1] = 3; rows[2] = 3; cfs[2] = 8; break;
      case 7: rows[0] = 2; cfs[0] = 4; rows[1] = 4; cfs[1] = 8; break;
      default: break;
    }
  }
  // shared-memory reference of element e of a stage array (bit 15: + stage index) / of a scratch word
  OB_HD uint32_t ref_st(const double* arr, int e, int dstage = 0) const { return (uint32_t)((arr - sm.Z) + e * S1 + dstage) | REF_S; }
  OB_HD uint32_t ref_abs(const double* p) const { return (uint32_t)(p - sm.Z); }
  OB_HD double ld(uint32_t ref, int s) const { return sm.Z[(ref & 0x7fffu) + ((ref & REF_S) ? s : 0)]; }
  // feedback-gain inputs kept per stage for fwd_prep (in the K|KAP region): F(6,b), F(7,b) for b in {0,1,2,5},
  // the 2x2 pivot block, f6, f7
  OB_HD static int kraw_slot(int e) {   // e = packed index of the 8x8 system, or 36 + a for f[a]
    switch (e) {
      case 21: return 0; case 22: return 1; case 23: return 2; case 26: return 3;
      case 28: return 4; case 29: return 5; case 30: return 6; case 33: return 7;
      case 27: return 8; case 34: return 9; case 35: return 10; case 42: return 11; case 43: return 12;
      default: return -1;
    }
  }
  // lane -> task tables of the sweep (every thread fills a few words; depends on the sizes only)
  OB_HD void fill_tables(int tid) const {
    const int I6[6] = {0, 1, 2, 5, 6, 7};
    uint32_t* tab = sm.TAB;
    int rows[4], cfs[4];
    const double* Wb = sm.RIC + 44;   // W[r][jj-2], jj = 2..5
    const double* pcb = sm.RIC + 80;
    for (int t = tid; t < 30; t += sm.T) {
      if (t < 24) {                                  // W[r][jj] = sum_p cf[c_p] P_{s+1}[r, row_p],  jj = 2..5
        const int r = t / 4, jj = 2 + t % 4;
        at_col_tab(I6[jj], rows, cfs);
        for (int p = 0; p < 4; ++p) tab[TW + 4 * t + p] = ref_st(sm.DYN, cfs[p]) | (ref_st(sm.PM, symi(r, rows[p]), 1) << 16);
      } else {                                       // pc[r] = p_{s+1}[r] - sum_j cd[j] P_{s+1}[r, j]   (cd = DYN 10..12)
        const int r = t - 24;
        for (int p = 0; p < 4; ++p)
          tab[TW + 4 * t + p] = ref_st(sm.DYN, p < 3 ? 10 + p : 9) | (ref_st(sm.PM, symi(r, p < 3 ? p : 0), 1) << 16);
      }
    }
    for (int t = tid; t < 29; t += sm.T) {
      if (t < 21) {                                  // F[a][b] = H[a][b] + sum_p cf[c_p] (P At)[row_p][b],  a, b in I6
        int ia = 0;
        while ((ia + 1) * (ia + 2) / 2 <= t) ++ia;
        const int ib = t - ia * (ia + 1) / 2, a = I6[ia], b = I6[ib];
        at_col_tab(a, rows, cfs);
        for (int p = 0; p < 4; ++p) {
          // columns x, y of At are unit vectors: (P At)[r][x|y] is P itself
          const uint32_t op = (ib >= 2) ? ref_abs(Wb + rows[p] * 4 + (ib - 2)) : ref_st(sm.PM, symi(rows[p], ib), 1);
          tab[TF + 4 * t + p] = ref_st(sm.DYN, cfs[p]) | (op << 16);
        }
        const int cls = (a != b) ? 0 : (a < 3 ? 1 : (a >= 6 ? 2 : 3));
        const int e = symi(a, b);
        tab[TFH + t] = (uint32_t)e | ((uint32_t)cls << 8) | ((uint32_t)(kraw_slot(e) + 1) << 12);
      } else {                                       // f[a] = r[a] + sum_p cf[c_p] pc[row_p]
        const int a = t - 21;
        at_col_tab(a, rows, cfs);
        for (int p = 0; p < 4; ++p) tab[TF + 4 * t + p] = ref_st(sm.DYN, cfs[p]) | (ref_abs(pcb + rows[p]) << 16);
        tab[TFH + t] = (uint32_t)(36 + a) | ((uint32_t)(kraw_slot(36 + a) + 1) << 12);
      }
    }
    for (int t = tid; t < 27; t += sm.T) {          // elimination of (v, w) = variables 6, 7
      // entry (x, y) of the 8x8 system: computed entries sit in the scratch, rows/columns v_prev, w_prev are H itself
      auto fref = [&](int x, int y) -> uint32_t {
        const int e = symi(x, y);
        const bool inH = (x == 3 || x == 4 || y == 3 || y == 4);
        return inH ? ref_st(sm.H, e) : ref_abs(sm.RIC + e);
      };
      if (t < 21) {
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= t) ++a;
        const int b = t - a * (a + 1) / 2;
        tab[TB + 3 * t] = fref(a, b) | (fref(6, a) << 16);
        tab[TB + 3 * t + 1] = fref(7, a) | (fref(6, b) << 16);
        tab[TB + 3 * t + 2] = fref(7, b);
      } else {
        // p_s[a] = f[a] - F(6,a) kap0 - F(7,a) kap1, kap = Q^-1 (f6, f7): same operand pattern as the entries above
        const int a = t - 21;
        tab[TB + 3 * t] = ref_abs(sm.RIC + 36 + a) | (fref(6, a) << 16);
        tab[TB + 3 * t + 1] = fref(7, a) | (ref_abs(sm.RIC + 42) << 16);
        tab[TB + 3 * t + 2] = ref_abs(sm.RIC + 43);
      }
    }
  }
  // The sweep is the longest dependent chain of an iteration (3 sub-steps x N stages).  Each sub-step has <= 32 tasks
  // and runs on the stage warp alone, separated by warp barriers; the other warps wait at one block barrier.
  template <bool RESTO>
  OB_HD void ric_terminal(int t, double mu, double dw, double dc) const {
    // stage N: cost-to-go = its own 6x6 block (terminal equality folded in Levenberg-Marquardt style; in the restoration
    // pass the fold is the elimination of pt, nt: per-component regularisation tdc and residual mu tcta + tctb)
    const int s = N;
    const Glob& G = *sm.G;
    if (t < 21) {
      double v = sm.st(sm.H, t, s);
      if (t == 0 || t == 2 || t == 5) {   // (0,0) (1,1) (2,2)
        v += dw;
        if (free_) v += RESTO ? 1.0 / G.tdc[t == 0 ? 0 : (t == 2 ? 1 : 2)] : 1.0 / dc;
      }
      sm.st(sm.PM, t, s) = v;
    } else if (t < 27) {
      const int a = t - 21;
      double r = mu * sm.st(sm.RA, a, s) + sm.st(sm.RB, a, s);
      if (a < 3 && free_) {
        if constexpr (RESTO) r -= (mu * G.tcta[a] + G.tctb[a]) / G.tdc[a];
        else r -= (sm.st(sm.Z, a, s) - xref(N)[a]) / dc;
      }
      sm.st(sm.PV, a, s) = r;
    }
  }
  // decode this lane's three tasks (once per sweep)
  OB_HD void sweep_load(int t, SweepRegs& r) const {
    const uint32_t* tab = sm.TAB;
    const uint32_t zero = ref_st(sm.DYN, 9);            // DYN element 9 is 0 at every stage
    const uint32_t dummy = ref_abs(sm.RIC + RIC_DUMMY + t);
    auto off = [](uint32_t ref) -> uint32_t { return (ref & 0x7fffu) * 8u; };
    uint32_t fl = 0;
    // sub-step 1
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t w = (t < 30) ? tab[TW + 4 * t + p] : (zero | (zero << 16));
      r.wa[p] = off(w); r.wb[p] = off(w >> 16);
    }
    const bool pc = (t >= 24 && t < 30);
    r.wp = off(pc ? ref_st(sm.PV, t - 24, 1) : zero);
    r.wd = off((t < 24) ? ref_abs(sm.RIC + 44 + t) : (pc ? ref_abs(sm.RIC + 80 + (t - 24)) : dummy));
    if (pc) fl |= SW_PC;
    // sub-step 2
    uint32_t f0 = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t f = (t < 29) ? tab[TF + 4 * t + p] : (zero | (zero << 16));
      if (p == 0) f0 = f;
      r.fa[p] = off(f); r.fb[p] = off(f >> 16);
    }
    if ((f0 >> 31) & 1u) fl |= SW_FS;     // all four operands of a task live in the same array
    const uint32_t fmeta = (t < 29) ? tab[TFH + t] : 0u;
    const int e = (int)(fmeta & 255u), cls = (int)((fmeta >> 8) & 15u), slot = (int)(fmeta >> 12) - 1;
    if (t < 21) {
      r.fx = r.fy = off(ref_st(sm.H, e));
      if (cls == 2) fl |= SW_DW_ALL;
      if (cls == 1) fl |= SW_DW_GE1;
      if (cls == 3 && free_) fl |= SW_DW_EQ0;
    } else if (t < 29) {
      r.fx = off(ref_st(sm.RA, e - 36)); r.fy = off(ref_st(sm.RB, e - 36));
      fl |= SW_FF;
    } else {
      r.fx = r.fy = off(zero);
    }
    r.fd = off((t < 29) ? ref_abs(sm.RIC + e) : dummy);
    r.fk = off((t < 29 && slot >= 0) ? ref_st(sm.K, slot) : dummy);
    if (t < 29 && slot >= 0) fl |= SW_FK;
    // sub-step 3
    const uint32_t zz = zero | (zero << 16);
    const uint32_t b0 = (t < 27) ? tab[TB + 3 * t] : zz, b1 = (t < 27) ? tab[TB + 3 * t + 1] : zz, b2 = (t < 27) ? tab[TB + 3 * t + 2] : zero;
    const uint32_t refs[5] = {b0 & 0xffffu, b0 >> 16, b1 & 0xffffu, b1 >> 16, b2 & 0xffffu};
#pragma unroll
    for (int p = 0; p < 5; ++p) { r.bo[p] = off(refs[p]); if (refs[p] & REF_S) fl |= SW_B0 << p; }
    r.bd = off((t < 21) ? ref_st(sm.PM, t) : ((t < 27) ? ref_st(sm.PV, t - 21) : dummy));
    if (t < 27) fl |= SW_BD;
    if (t >= 21 && t < 27) fl |= SW_PV;
    r.ric = off(ref_abs(sm.RIC));
    r.bad = (uint32_t)((const char*)&sm.G->bad - (const char*)sm.Z);
    r.flags = fl;
  }
  OB_HD SwMemPtr sweep_mem() const { return SwMemPtr{(char*)sm.Z}; }
  OB_HD void ric_finish(int t) const {
    Glob& G = *sm.G;
    if (t == 0 && free_ && !(sm.st(sm.PM, 20, 0) > 0)) G.bad = 1;
  }

  // ------------------------------------------------------------------------------------------------
  // roll-out.  fwd_prep (lane k): closed-loop map  xi_{k+1} = ACL_k xi_k + CCL_k  (ACL|CCL overwrite H|RA);
  // fwd_step (lanes 0..5, sequential in s); fwd_post (lane k): du_k, dz_k, dynamics-multiplier steps
  // ------------------------------------------------------------------------------------------------
  OB_HD void fwd_prep(int k) const {
    Glob& G = *sm.G;
    if (k == 0) {
      double dT = 0.0;
      if (free_) dT = sm.st(sm.PV, 5, 0) / sm.st(sm.PM, 20, 0);
      G.dT = dT;
#pragma unroll
      for (int a = 0; a < 6; ++a) sm.st(sm.XI, a, 0) = (a == 5) ? dT : 0.0;
    }
    if (k >= N) return;
    // feedback gains K = -Q^-1 F_ux, kap = Q^-1 f_u from what the sweep left in the K|KAP region (kraw_slot) and the
    // untouched (v_prev, w_prev) columns of H - read before ACL|CCL overwrite H|RA below
    double Kr[2][6], kap[2];
    {
      const double q00 = sm.st(sm.K, 8, k), q01 = sm.st(sm.K, 9, k), q11 = sm.st(sm.K, 10, k);
      const double idet = ob_rcp(q00 * q11 - q01 * q01);
      const double i00 = q11 * idet, i01 = -q01 * idet, i11 = q00 * idet;
      const double F6[6] = {sm.st(sm.K, 0, k), sm.st(sm.K, 1, k), sm.st(sm.K, 2, k), sm.st(sm.H, symi(6, 3), k),
                            sm.st(sm.H, symi(6, 4), k), sm.st(sm.K, 3, k)};
      const double F7[6] = {sm.st(sm.K, 4, k), sm.st(sm.K, 5, k), sm.st(sm.K, 6, k), sm.st(sm.H, symi(7, 3), k),
                            sm.st(sm.H, symi(7, 4), k), sm.st(sm.K, 7, k)};
      const double f6 = sm.st(sm.K, 11, k), f7 = sm.st(sm.K, 12, k);
#pragma unroll
      for (int b = 0; b < 6; ++b) { Kr[0][b] = -(i00 * F6[b] + i01 * F7[b]); Kr[1][b] = -(i01 * F6[b] + i11 * F7[b]); }
      kap[0] = i00 * f6 + i01 * f7; kap[1] = i01 * f6 + i11 * f7;
    }
    const double fth0 = sm.st(sm.DYN, 0, k), fth1 = sm.st(sm.DYN, 1, k), bv0 = sm.st(sm.DYN, 2, k), bv1 = sm.st(sm.DYN, 3, k);
    const double bw = sm.st(sm.DYN, 4, k), fT0 = sm.st(sm.DYN, 5, k), fT1 = sm.st(sm.DYN, 6, k), fT2 = sm.st(sm.DYN, 7, k);
    const double cd0 = sm.st(sm.CD, 0, k), cd1 = sm.st(sm.CD, 1, k), cd2 = sm.st(sm.CD, 2, k);
    double* ACL = sm.H;            // [36][S1]
    double* CCL = sm.H + 36 * S1;  // [6][S1]  (the RA region)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const double A0 = (b == 0 ? 1.0 : 0.0) + (b == 2 ? fth0 : 0.0) + (b == 5 ? fT0 : 0.0);
      const double A1 = (b == 1 ? 1.0 : 0.0) + (b == 2 ? fth1 : 0.0) + (b == 5 ? fT1 : 0.0);
      const double A2 = (b == 2 ? 1.0 : 0.0) + (b == 5 ? fT2 : 0.0);
      ACL[(0 * 6 + b) * S1 + k] = A0 + bv0 * Kr[0][b];
      ACL[(1 * 6 + b) * S1 + k] = A1 + bv1 * Kr[0][b];
      ACL[(2 * 6 + b) * S1 + k] = A2 + bw * Kr[1][b];
      ACL[(3 * 6 + b) * S1 + k] = Kr[0][b];
      ACL[(4 * 6 + b) * S1 + k] = Kr[1][b];
      ACL[(5 * 6 + b) * S1 + k] = (b == 5) ? 1.0 : 0.0;
    }
    CCL[0 * S1 + k] = bv0 * kap[0] + cd0;
    CCL[1 * S1 + k] = bv1 * kap[0] + cd1;
    CCL[2 * S1 + k] = bw * kap[1] + cd2;
    CCL[3 * S1 + k] = kap[0];
    CCL[4 * S1 + k] = kap[1];
    CCL[5 * S1 + k] = 0.0;
  }
  OB_HD void fwd_step(int lane, int s) const {
    if (lane < 6) {
      const double* ACL = sm.H;
      const double* CCL = sm.H + 36 * S1;
      double v = CCL[lane * S1 + s];
#pragma unroll
      for (int b = 0; b < 6; ++b) v += ACL[(lane * 6 + b) * S1 + s] * sm.st(sm.XI, b, s);
      sm.st(sm.XI, lane, s + 1) = v;
    }
  }
  template <bool RESTO>
  OB_HD void fwd_post(int k, double dc, double mu) const {
    Glob& G = *sm.G;
    double xi[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) xi[a] = sm.st(sm.XI, a, k);
    if (k < N) {
      // du_k = xi_{k+1}[3:5]
      sm.st(sm.DU, 0, k) = sm.st(sm.XI, 3, k + 1);
      sm.st(sm.DU, 1, k) = sm.st(sm.XI, 4, k + 1);
    } else {
      sm.st(sm.DU, 0, k) = 0.0; sm.st(sm.DU, 1, k) = 0.0;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) sm.st(sm.DZ, a, k) = (k >= 1) ? xi[a] : 0.0;
    if (k >= 1) {
      // dy_{k-1} = (P_k d xi_k - p_k)[0:3]
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double sacc = -sm.st(sm.PV, a, k);
#pragma unroll
        for (int b = 0; b < 6; ++b) sacc += sm.st(sm.PM, symi(a, b), k) * xi[b];
        sm.st(sm.DYD, a, k - 1) = sacc;
      }
    }
    if (k == N) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        sm.st(sm.DYD, a, N) = 0.0;
        if (free_) {
          if constexpr (RESTO) G.dyt[a] = (xi[a] + (mu * G.tcta[a] + G.tctb[a])) / G.tdc[a];
          else G.dyt[a] = (xi[a] + (sm.st(sm.Z, a, N) - xref(N)[a])) / dc;
        }
      }
    }
  }

  OB_HD static void ftb(double S, double Z, double dS, double mu, double tau, double& amax, double& az, double& sls) {
    const double iS = ob_rcp(S);
    const double dZ = (mu - Z * dS) * iS - Z;
    // tau*S/(-dS) < amax  <=>  tau*S < amax*(-dS) for dS < 0: compare by cross-multiplication, divide only when it binds
    if (dS < 0 && tau * S < amax * -dS) amax = -tau * S / dS;
    if (dZ < 0 && tau * Z < az * -dZ) az = -tau * Z / dZ;
    sls += dS * iS;
  }

  // the same for a relaxation variable n >= 0 with bound multiplier V and steps dn, dV
  OB_HD static void ftb_n(double n, double V, double dn, double dV, double tau, double& amax, double& az, double& sls) {
    if (dn < 0 && tau * n < amax * -dn) amax = -tau * n / dn;
    if (dV < 0 && tau * V < az * -dV) az = -tau * V / dV;
    sls += dn / n;
  }
  // restoration pass: steps of a relaxed row from g = grad d . dX; returns dn (dS through the reference)
  OB_HD static double relaxed_row_steps(double S, double Z, double n, double V, double d, double g, double mu, double tau,
                                        double& dS, double& amax, double& az, double& sls) {
    const double iZ = ob_rcp(Z), iV = ob_rcp(V), sig = ob_rcp(S * iZ + n * iV);
    const double t = -sig * ((d + n - S) + (mu - n * (RESTO_RHO - Z)) * iV - (mu - S * Z) * iZ);
    const double dZ = t - sig * g;
    double dn;
    relaxed_steps(S, Z, n, V, mu, RESTO_RHO, dZ, dS, dn);
    ftb(S, Z, dS, mu, tau, amax, az, sls);
    ftb_n(n, V, dn, (RESTO_RHO - Z - V) - dZ, tau, amax, az, sls);
    return dn;
  }

  // ------------------------------------------------------------------------------------------------
  // steps of the slacks, fraction to the boundary, directional derivative - stage part (lane k)
  // ------------------------------------------------------------------------------------------------
  template <bool RESTO>
  OB_HD void backsub_stage(int k, const BlockRegs<EMAX>& br, double mu, double tau, double* part) const {
    Glob& G = *sm.G;
    double z[3], u[2], up[2], zn[3], dz[3], du[2], dup[2];
    stage_point(k, 0.0, z, u, up, zn);
#pragma unroll
    for (int j = 0; j < 3; ++j) dz[j] = sm.st(sm.DZ, j, k);
#pragma unroll
    for (int j = 0; j < 2; ++j) { du[j] = sm.st(sm.DU, j, k); dup[j] = (k >= 1) ? sm.st(sm.DU, j, k - 1) : 0.0; }
    const double T = free_ ? G.T : 1.0, h = T * G.Ts, dT = G.dT;
    StageVals sv;
    stage_vals(k, z, u, up, zn, T, sv);
    double amax = 1.0, az = 1.0, sls = 0.0, dphi = 0.0;
    {
      const double* gf = br.gf;
      if (k >= 1) dphi += gf[0] * dz[0] + gf[1] * dz[1] + gf[2] * dz[2];
      if (k >= 1 && k < N) dphi += gf[3] * dup[0] + gf[4] * dup[1];
      if (free_) dphi += gf[5] * dT;
      if (k < N) dphi += gf[6] * du[0] + gf[7] * du[1];
    }
    if (k >= 1) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double S0 = sm.st(sm.SXY, j, k), S1_ = sm.st(sm.SXY, 2 + j, k);
        if constexpr (RESTO) {
          double d0, d1;
          const double n0 = relaxed_row_steps(S0, sm.st(sm.ZXY, j, k), sm.st(sm.NXY, j, k), sm.st(sm.VXY, j, k), sv.dxy[j], dz[j], mu, tau, d0, amax, az, sls);
          const double n1 = relaxed_row_steps(S1_, sm.st(sm.ZXY, 2 + j, k), sm.st(sm.NXY, 2 + j, k), sm.st(sm.VXY, 2 + j, k), sv.dxy[2 + j], -dz[j], mu, tau, d1, amax, az, sls);
          sm.st(sm.DSXY, j, k) = d0; sm.st(sm.DSXY, 2 + j, k) = d1;
          sm.st(sm.DNXY, j, k) = n0; sm.st(sm.DNXY, 2 + j, k) = n1;
          dphi += RESTO_RHO * (n0 + n1);
        } else {
        double d0 = dz[j] + (sv.dxy[j] - S0), d1 = -dz[j] + (sv.dxy[2 + j] - S1_);
        sm.st(sm.DSXY, j, k) = d0; sm.st(sm.DSXY, 2 + j, k) = d1;
        ftb(S0, sm.st(sm.ZXY, j, k), d0, mu, tau, amax, az, sls);
        ftb(S1_, sm.st(sm.ZXY, 2 + j, k), d1, mu, tau, amax, az, sls);
        }
      }
    }
    if (k < N) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double ga = (up[j] - u[j]) / h;
        double dga = (dup[j] - du[j]) / h - (free_ ? ga / T * dT : 0.0);
        double dS[4] = {du[j] + (sv.dub[j] - sm.st(sm.SUB, j, k)), -du[j] + (sv.dub[2 + j] - sm.st(sm.SUB, 2 + j, k)),
                        dga + (sv.dub[4 + j] - sm.st(sm.SUB, 4 + j, k)), -dga + (sv.dub[6 + j] - sm.st(sm.SUB, 6 + j, k))};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          sm.st(sm.DSUB, 2 * q + j, k) = dS[q];
          ftb(sm.st(sm.SUB, 2 * q + j, k), sm.st(sm.ZUB, 2 * q + j, k), dS[q], mu, tau, amax, az, sls);
        }
      }
    }
    if (k == 0 && free_) {
      G.dSTb[0] = dT + ((T - P.T_min) - G.STb[0]);
      G.dSTb[1] = -dT + ((G.Tmax - T) - G.STb[1]);
      ftb(G.STb[0], G.ZTb[0], G.dSTb[0], mu, tau, amax, az, sls);
      ftb(G.STb[1], G.ZTb[1], G.dSTb[1], mu, tau, amax, az, sls);
    }
    if (k == N && has_term) {
      if constexpr (RESTO) {
        const double dtm[3] = {z[0] - G.term[0], z[1] - G.term[1], G.term[2] - z[1]}, gtm[3] = {dz[0], dz[1], -dz[1]};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          G.dntm[j] = relaxed_row_steps(G.Stm[j], G.Ztm[j], G.ntm[j], G.Vtm[j], dtm[j], gtm[j], mu, tau, G.dStm[j], amax, az, sls);
          dphi += RESTO_RHO * G.dntm[j];
        }
      } else {
      G.dStm[0] = dz[0] + ((z[0] - G.term[0]) - G.Stm[0]);
      G.dStm[1] = dz[1] + ((z[1] - G.term[1]) - G.Stm[1]);
      G.dStm[2] = -dz[1] + ((G.term[2] - z[1]) - G.Stm[2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) ftb(G.Stm[j], G.Ztm[j], G.dStm[j], mu, tau, amax, az, sls);
      }
    }
    if constexpr (RESTO) {
      if (k == N && free_) {   // pt, nt of the relaxed terminal equality from the step of its multiplier
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double dy = G.dyt[j], y = G.yt[j];
          const double dVp = (RESTO_RHO - y - G.Vpt[j]) - dy, dVn = (RESTO_RHO + y - G.Vnt[j]) + dy;
          G.dpt[j] = (mu - G.pt[j] * G.Vpt[j] - G.pt[j] * dVp) / G.Vpt[j];
          G.dnt[j] = (mu - G.nt[j] * G.Vnt[j] - G.nt[j] * dVn) / G.Vnt[j];
          ftb_n(G.pt[j], G.Vpt[j], G.dpt[j], dVp, tau, amax, az, sls);
          ftb_n(G.nt[j], G.Vnt[j], G.dnt[j], dVn, tau, amax, az, sls);
          dphi += RESTO_RHO * (G.dpt[j] + G.dnt[j]);
        }
      }
    }
    part[QS_DPHI] = dphi - mu * sls; part[QN_AMAX] = amax; part[QN_AZ] = az;
  }

  // block part: back-substitution of the dual block
  template <bool RESTO>
  OB_HD void backsub_block(int tid, const BlockRegs<EMAX>& br, double mu, double tau, double* part) const {
    const Glob& G = *sm.G;
    const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
    const double z0 = sm.st(sm.Z, 0, k), z1 = sm.st(sm.Z, 1, k), z2 = sm.st(sm.Z, 2, k);
    BlkGeo b;
    ob_sincos(z2, &b.st, &b.ct);
    const double st = b.st, ct = b.ct;
    b.tx = z0 + G.off * ct; b.ty = z1 + G.off * st;
    const double dp[3] = {(k >= 1) ? sm.st(sm.DZ, 0, k) : 0.0, (k >= 1) ? sm.st(sm.DZ, 1, k) : 0.0, (k >= 1) ? sm.st(sm.DZ, 2, k) : 0.0};
    double a1 = 0, a2 = 0, bl = 0;
#pragma unroll
    for (int j = 0; j < EMAX; ++j) {
      if (j < E) {
        const int r = r0 + j;
        a1 += sm.A[2 * r] * br.lam[j]; a2 += sm.A[2 * r + 1] * br.lam[j]; bl += bk(k, r) * br.lam[j];
      }
    }
    b.a1 = a1; b.a2 = a2;
    const double m0 = br.mu[0], m1 = br.mu[1], m2 = br.mu[2], m3 = br.mu[3];
    const double dn = 1.0 - a1 * a1 - a2 * a2;
    const double dd = -(G.g[0] * m0 + G.g[1] * m1 + G.g[2] * m2 + G.g[3] * m3) + b.tx * a1 + b.ty * a2 - bl - P.dmin;
    const double y1 = br.ye[0], y2 = br.ye[1];
    const double Sn = br.Sn, Zn = br.Zn, Sd = br.Sd, Zd = br.Zd;
    const double sn = Zn / Sn, sd = Zd / Sd;
    const double tn = (mu - Sn * Zn) / Sn - sn * (dn - Sn);
    const double ydv = -Zd, c1 = y1 + ydv * G.off;
    const double hc[3][2] = {{ydv, 0.0}, {0.0, ydv}, {-c1 * st - y2 * ct, c1 * ct - y2 * st}};
    const double offt = G.off * (-st * a1 + ct * a2);
    const double dpose[3] = {a1, a2, offt};
    double et[5], ht[2] = {-2 * tn * a1, -2 * tn * a2};
#pragma unroll
    for (int a = 0; a < 5; ++a) et[a] = mu * sm.bl(sm.ETA, a, tid) + sm.bl(sm.ETA, 5 + a, tid);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int a = 0; a < 5; ++a) et[a] += sm.bl(sm.ETA, 10 + 5 * c + a, tid) * dp[c];
      ht[0] -= hc[c][0] * dp[c]; ht[1] -= hc[c][1] * dp[c];
    }
    double amax = 1.0, az = 1.0, sls = 0.0;
    double da1 = 0, da2 = 0, qdw = 0;
    const int nrow = E + 4;
#pragma unroll 1
    for (int j = 0; j < nrow; ++j) {
      double wv, S, Z, yv[5];
      row_regs(br, j, E, wv, S, Z);
      row_y(b, k, r0, E, j, yv);
      const double iS = ob_rcp(S), sig = Z * iS;
      const double gl = y1 * yv[3] + y2 * yv[4] - Z + 2 * Zn * (a1 * yv[0] + a2 * yv[1]) - Zd * yv[2];
      double s_ = (mu - S * Z) * iS - sig * (wv - S) - gl;
#pragma unroll
      for (int a = 0; a < 5; ++a) s_ -= yv[a] * et[a];
      const double dwv = s_ * ob_rcp(fmax(sig, SIG_MIN));
      if (j < E) { sm.st(sm.DLAM, r0 + j, k) = dwv; da1 += yv[0] * dwv; da2 += yv[1] * dwv; }
      else sm.bl(sm.DMU, j - E, tid) = dwv;
      qdw += yv[2] * dwv;
      ftb(S, Z, dwv + (wv - S), mu, tau, amax, az, sls);
    }
    sm.bl(sm.DYE, 0, tid) = et[3]; sm.bl(sm.DYE, 1, tid) = et[4];
    double ada = a1 * da1 + a2 * da2;
    if (sn >= 1.0) ada = (a1 * (et[0] + ht[0]) + a2 * (et[1] + ht[1])) / (2 * Zn + 4 * sn * (a1 * a1 + a2 * a2));
    const double dSn = -2 * ada + (dn - Sn);
    double dSd, rdn = 0.0;
    if constexpr (RESTO) {
      // relaxed distance row: multiplier step from the solve when the row is active, from the primal direction otherwise
      const double nd = sm.bl(sm.ND, 0, tid), Vd = sm.bl(sm.VD, 0, tid);
      const double iZ = ob_rcp(Zd), iV = ob_rcp(Vd), sde = ob_rcp(Sd * iZ + nd * iV);
      const double tds = -((dd + nd - Sd) + (mu - nd * (RESTO_RHO - Zd)) * iV - (mu - Sd * Zd) * iZ);
      const double gd = qdw + dpose[0] * dp[0] + dpose[1] * dp[1] + dpose[2] * dp[2];
      const double dZ = (sde >= 1.0) ? -et[2] : sde * (tds - gd);
      double dnd;
      relaxed_steps(Sd, Zd, nd, Vd, mu, RESTO_RHO, dZ, dSd, dnd);
      sm.bl(sm.DND, 0, tid) = dnd;
      ftb_n(nd, Vd, dnd, (RESTO_RHO - Zd - Vd) - dZ, tau, amax, az, sls);
      rdn = RESTO_RHO * dnd;
    } else {
      if (sd >= 1.0) dSd = (mu - Sd * Zd + Sd * et[2]) / Zd;
      else dSd = qdw + (dd - Sd) + dpose[0] * dp[0] + dpose[1] * dp[1] + dpose[2] * dp[2];
    }
    sm.bl(sm.DSN, 0, tid) = dSn; sm.bl(sm.DSD, 0, tid) = dSd;
    ftb(Sn, Zn, dSn, mu, tau, amax, az, sls);
    ftb(Sd, Zd, dSd, mu, tau, amax, az, sls);
    part[QS_DPHI] = rdn - mu * sls; part[QN_AMAX] = amax; part[QN_AZ] = az;
  }

  // ------------------------------------------------------------------------------------------------
  // trial point X + a dX, S + a dS: partial objective, constraint violation theta, sum log S
  // ------------------------------------------------------------------------------------------------
  template <bool RESTO>
  OB_HD void trial_stage(int k, double a, double* part) const {
    const Glob& G = *sm.G;
    double z[3], u[2], up[2], zn[3];
    stage_point(k, a, z, u, up, zn);
    if (a == 0.0) { /* stage_point skips the direction at a == 0 */ }
    const double T = free_ ? G.T + a * G.dT : 1.0;
    StageVals sv;
    stage_vals(k, z, u, up, zn, T, sv);
    double th = 0;
    LogAcc lg;
    if (k < N) {
      th += fabs(sv.cd[0]) + fabs(sv.cd[1]) + fabs(sv.cd[2]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { double S = sm.st(sm.SUB, j, k) + a * sm.st(sm.DSUB, j, k); th += fabs(sv.dub[j] - S); lg.add(S); }
    }
    double fR = 0.0;
    if constexpr (RESTO) {
      if (k >= 1) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { const double ref = sm.st(sm.ZR, j, k), e = z[j] - ref; fR += 0.5 * G.zeta * dr2(ref) * e * e; }
      }
      if (k < N) {
#pragma unroll
        for (int j = 0; j < 2; ++j) { const double ref = sm.st(sm.UR, j, k), e = u[j] - ref; fR += 0.5 * G.zeta * dr2(ref) * e * e; }
      }
      if (k == 0 && free_) { const double e = T - G.TR; fR += 0.5 * G.zeta * dr2(G.TR) * e * e; }
    }
    if (k >= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double S = sm.st(sm.SXY, j, k) + a * sm.st(sm.DSXY, j, k);
        if constexpr (RESTO) {
          const double n = sm.st(sm.NXY, j, k) + a * sm.st(sm.DNXY, j, k);
          th += fabs(sv.dxy[j] + n - S); lg.add(S); lg.add(n); fR += RESTO_RHO * n;
        } else { th += fabs(sv.dxy[j] - S); lg.add(S); }
      }
    }
    if (k == 0 && free_) {
      double S = G.STb[0] + a * G.dSTb[0]; th += fabs(T - P.T_min - S); lg.add(S);
      S = G.STb[1] + a * G.dSTb[1]; th += fabs(G.Tmax - T - S); lg.add(S);
    }
    if (k == N && free_) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if constexpr (RESTO) {
          const double pp = G.pt[j] + a * G.dpt[j], nn = G.nt[j] + a * G.dnt[j];
          th += fabs(z[j] - xref(N)[j] - pp + nn); lg.add(pp); lg.add(nn); fR += RESTO_RHO * (pp + nn);
        } else th += fabs(z[j] - xref(N)[j]);
      }
    }
    if (k == N && has_term) {
      const double dtm[3] = {z[0] - G.term[0], z[1] - G.term[1], G.term[2] - z[1]};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double S = G.Stm[j] + a * G.dStm[j];
        if constexpr (RESTO) {
          const double n = G.ntm[j] + a * G.dntm[j];
          th += fabs(dtm[j] + n - S); lg.add(S); lg.add(n); fR += RESTO_RHO * n;
        } else { th += fabs(dtm[j] - S); lg.add(S); }
      }
    }
    part[0] = RESTO ? fR : sv.f; part[1] = th; part[2] = lg.value();
  }
  template <bool RESTO>
  OB_HD void trial_block(int tid, const BlockRegs<EMAX>& br, double a, double* part) const {
    const Glob& G = *sm.G;
    const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
    const double ak = (k >= 1) ? a : 0.0;
    const double z0 = sm.st(sm.Z, 0, k) + ak * sm.st(sm.DZ, 0, k), z1 = sm.st(sm.Z, 1, k) + ak * sm.st(sm.DZ, 1, k);
    const double z2 = sm.st(sm.Z, 2, k) + ak * sm.st(sm.DZ, 2, k);
    double st, ct;
    ob_sincos(z2, &st, &ct);
    const double tx = z0 + G.off * ct, ty = z1 + G.off * st;
    double th = 0, a1 = 0, a2 = 0, bl = 0;
    LogAcc lg;
#pragma unroll
    for (int j = 0; j < EMAX; ++j) {
      if (j < E) {
        const int r = r0 + j;
        const double l0 = br.lam[j], dl = sm.st(sm.DLAM, r, k), S0 = br.Sl[j];
        const double l = l0 + a * dl, S = S0 + a * (dl + (l0 - S0));
        a1 += sm.A[2 * r] * l; a2 += sm.A[2 * r + 1] * l; bl += bk(k, r) * l;
        th += fabs(l - S); lg.add(S);
      }
    }
    double m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double m0 = br.mu[q], dm = sm.bl(sm.DMU, q, tid), S0 = br.Sm_[q];
      m[q] = m0 + a * dm;
      const double S = S0 + a * (dm + (m0 - S0));
      th += fabs(m[q] - S); lg.add(S);
    }
    th += fabs(m[0] - m[2] + ct * a1 + st * a2) + fabs(m[1] - m[3] - st * a1 + ct * a2);
    {
      const double S = br.Sn + a * sm.bl(sm.DSN, 0, tid);
      th += fabs(1.0 - a1 * a1 - a2 * a2 - S); lg.add(S);
    }
    {
      const double S = br.Sd + a * sm.bl(sm.DSD, 0, tid);
      const double d = -(G.g[0] * m[0] + G.g[1] * m[1] + G.g[2] * m[2] + G.g[3] * m[3]) + tx * a1 + ty * a2 - bl - P.dmin;
      if constexpr (RESTO) {
        const double n = sm.bl(sm.ND, 0, tid) + a * sm.bl(sm.DND, 0, tid);
        th += fabs(d + n - S); lg.add(S); lg.add(n);
        part[0] = RESTO_RHO * n;
      } else { th += fabs(d - S); lg.add(S); part[0] = 0.0; }
    }
    part[1] = th; part[2] = lg.value();
  }

  // ------------------------------------------------------------------------------------------------
  // accept the step: primal / slacks / equality multipliers with a, inequality multipliers with a_z (then
  // clipped into [mu/(ks S), ks mu/S] as IPOPT does)
  // ------------------------------------------------------------------------------------------------
  OB_HD static void upd(double& S, double& Z, double dS, double a, double az, double mu) {
    const double ks = 1e10;
    const double dZ = (mu - Z * dS) * ob_rcp(S) - Z;
    double Sn = S + a * dS, Zn = Z + az * dZ;
    const double mS = mu * ob_rcp(Sn);
    Zn = fmin(fmax(Zn, mS * 1e-10), ks * mS);
    S = Sn; Z = Zn;
  }
  // relaxation variable and its bound multiplier (restoration pass; before upd of the row: the step of V needs the
  // row's current S and Z)
  OB_HD static void upd_n(double S, double Z, double dS, double& n, double& V, double dn, double a, double az, double mu) {
    const double ks = 1e10;
    const double dZ = (mu - Z * dS) * ob_rcp(S) - Z;
    const double nn = n + a * dn;
    double Vn = V + az * ((RESTO_RHO - Z - V) - dZ);
    const double mN = mu * ob_rcp(nn);
    V = fmin(fmax(Vn, mN * 1e-10), ks * mN);
    n = nn;
  }
  // NOTE: reads neighbours' DZ/DU only through its own column, so no barrier is needed inside
  template <bool RESTO>
  OB_HD void update(int tid, BlockRegs<EMAX>& br, double a, double az, double mu) const {
    Glob& G = *sm.G;
    if (is_stage(tid)) {
      const int k = stage_lane(tid);
      if (k >= 1) {
#pragma unroll
        for (int j = 0; j < 3; ++j) sm.st(sm.Z, j, k) += a * sm.st(sm.DZ, j, k);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (RESTO) upd_n(sm.st(sm.SXY, j, k), sm.st(sm.ZXY, j, k), sm.st(sm.DSXY, j, k), sm.st(sm.NXY, j, k), sm.st(sm.VXY, j, k), sm.st(sm.DNXY, j, k), a, az, mu);
          upd(sm.st(sm.SXY, j, k), sm.st(sm.ZXY, j, k), sm.st(sm.DSXY, j, k), a, az, mu);
        }
      }
      if (k < N) {
#pragma unroll
        for (int j = 0; j < 2; ++j) sm.st(sm.U, j, k) += a * sm.st(sm.DU, j, k);
#pragma unroll
        for (int j = 0; j < 8; ++j) upd(sm.st(sm.SUB, j, k), sm.st(sm.ZUB, j, k), sm.st(sm.DSUB, j, k), a, az, mu);
#pragma unroll
        for (int j = 0; j < 3; ++j) sm.st(sm.YD, j, k) += a * sm.st(sm.DYD, j, k);
      }
      if constexpr (RESTO) {
        if (k == N && free_) {   // needs the current yt: the same lane moves yt below
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double ks = 1e10, dy = G.dyt[j], y = G.yt[j];
            const double pn = G.pt[j] + a * G.dpt[j], nn = G.nt[j] + a * G.dnt[j];
            const double Vp = G.Vpt[j] + az * ((RESTO_RHO - y - G.Vpt[j]) - dy), Vn = G.Vnt[j] + az * ((RESTO_RHO + y - G.Vnt[j]) + dy);
            G.Vpt[j] = fmin(fmax(Vp, mu / (ks * pn)), ks * mu / pn);
            G.Vnt[j] = fmin(fmax(Vn, mu / (ks * nn)), ks * mu / nn);
            G.pt[j] = pn; G.nt[j] = nn;
          }
        }
      }
      if (k == N && free_) {
#pragma unroll
        for (int j = 0; j < 3; ++j) G.yt[j] += a * G.dyt[j];
      }
      if (k == 0 && free_) {
        G.T += a * G.dT;
        upd(G.STb[0], G.ZTb[0], G.dSTb[0], a, az, mu);
        upd(G.STb[1], G.ZTb[1], G.dSTb[1], a, az, mu);
      }
      if (k == N && has_term) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if constexpr (RESTO) upd_n(G.Stm[j], G.Ztm[j], G.dStm[j], G.ntm[j], G.Vtm[j], G.dntm[j], a, az, mu);
          upd(G.Stm[j], G.Ztm[j], G.dStm[j], a, az, mu);
        }
      }
    }
    if (is_block(tid)) {
      const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
#pragma unroll
      for (int j = 0; j < EMAX; ++j) {
        if (j < E) {
          const double l0 = br.lam[j], dl = sm.st(sm.DLAM, r0 + j, k);
          upd(br.Sl[j], br.Zl[j], dl + (l0 - br.Sl[j]), a, az, mu);
          br.lam[j] = l0 + a * dl;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double m0 = br.mu[q], dm = sm.bl(sm.DMU, q, tid);
        upd(br.Sm_[q], br.Zm[q], dm + (m0 - br.Sm_[q]), a, az, mu);
        br.mu[q] = m0 + a * dm;
      }
      upd(br.Sn, br.Zn, sm.bl(sm.DSN, 0, tid), a, az, mu);
      if constexpr (RESTO) upd_n(br.Sd, br.Zd, sm.bl(sm.DSD, 0, tid), sm.bl(sm.ND, 0, tid), sm.bl(sm.VD, 0, tid), sm.bl(sm.DND, 0, tid), a, az, mu);
      upd(br.Sd, br.Zd, sm.bl(sm.DSD, 0, tid), a, az, mu);
      br.ye[0] += a * sm.bl(sm.DYE, 0, tid);
      br.ye[1] += a * sm.bl(sm.DYE, 1, tid);
    }
  }

  // ------------------------------------------------------------------------------------------------
  // watchdog checkpoint of the whole iterate (primal, slacks, multipliers) in HBM: the on-chip budget has no room
  // for a second copy, and a checkpoint is written only when a full step is taken on trust (rare), read back only
  // when that trust was misplaced (rarer).  Layout: block registers [element][thread] | stage arrays | scalars.
  // ------------------------------------------------------------------------------------------------
  OB_HD static int wd_doubles(int T, int S1) { return (6 * EMAX + 6) * T + 32 * S1 + 16; }
  OB_HD void wd_save(int tid, const BlockRegs<EMAX>& br, double* buf) const {
    const Glob& G = *sm.G;
    const int T = sm.T;
    int e = 0;
#pragma unroll
    for (int j = 0; j < EMAX; ++j) { buf[(e++) * T + tid] = br.lam[j]; buf[(e++) * T + tid] = br.Sl[j]; buf[(e++) * T + tid] = br.Zl[j]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { buf[(e++) * T + tid] = br.mu[q]; buf[(e++) * T + tid] = br.Sm_[q]; buf[(e++) * T + tid] = br.Zm[q]; }
    buf[(e++) * T + tid] = br.ye[0]; buf[(e++) * T + tid] = br.ye[1];
    buf[(e++) * T + tid] = br.Sn; buf[(e++) * T + tid] = br.Zn; buf[(e++) * T + tid] = br.Sd; buf[(e++) * T + tid] = br.Zd;
    double* sb = buf + (6 * EMAX + 6) * T;
    for (int i = tid; i < 32 * S1; i += T) sb[i] = sm.Z[i];     // Z U YD SXY ZXY SUB ZUB are contiguous
    if (tid == 0) {
      double* g = sb + 32 * S1;
      g[0] = G.T; g[1] = G.STb[0]; g[2] = G.STb[1]; g[3] = G.ZTb[0]; g[4] = G.ZTb[1];
      for (int j = 0; j < 3; ++j) { g[5 + j] = G.Stm[j]; g[8 + j] = G.Ztm[j]; g[11 + j] = G.yt[j]; }
    }
  }
  OB_HD void wd_restore(int tid, BlockRegs<EMAX>& br, const double* buf) const {
    Glob& G = *sm.G;
    const int T = sm.T;
    int e = 0;
#pragma unroll
    for (int j = 0; j < EMAX; ++j) { br.lam[j] = buf[(e++) * T + tid]; br.Sl[j] = buf[(e++) * T + tid]; br.Zl[j] = buf[(e++) * T + tid]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { br.mu[q] = buf[(e++) * T + tid]; br.Sm_[q] = buf[(e++) * T + tid]; br.Zm[q] = buf[(e++) * T + tid]; }
    br.ye[0] = buf[(e++) * T + tid]; br.ye[1] = buf[(e++) * T + tid];
    br.Sn = buf[(e++) * T + tid]; br.Zn = buf[(e++) * T + tid]; br.Sd = buf[(e++) * T + tid]; br.Zd = buf[(e++) * T + tid];
    const double* sb = buf + (6 * EMAX + 6) * T;
    for (int i = tid; i < 32 * S1; i += T) sm.Z[i] = sb[i];
    if (tid == 0) {
      const double* g = sb + 32 * S1;
      G.T = g[0]; G.STb[0] = g[1]; G.STb[1] = g[2]; G.ZTb[0] = g[3]; G.ZTb[1] = g[4];
      for (int j = 0; j < 3; ++j) { G.Stm[j] = g[5 + j]; G.Ztm[j] = g[8 + j]; G.yt[j] = g[11 + j]; }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // results -> HBM, once:  x [B,N+1,3]  u [B,N,2]  lam [B,N+1,R]  mu [B,N+1,4 no]  T  obj  status  iters
  // ------------------------------------------------------------------------------------------------
  OB_HD void store(int tid, const BlockRegs<EMAX>& br, size_t b, int status, int iters, double obj) const {
    store_stage(tid, b, status, iters, obj, kp.u + b * N * 2);
    store_blocks(tid, br, kp.lam + b * S1 * sm.R, kp.mu + b * S1 * 4 * no);
  }
  // poses, time scale, objective, status, iterations straight to the result arrays; inputs to `uo` [N,2]
  OB_HD void store_stage(int tid, size_t b, int status, int iters, double obj, double* uo) const {
    const Glob& G = *sm.G;
    if (is_stage(tid)) {
      const int k = stage_lane(tid);
      double* xo = kp.x + (b * S1 + k) * 3;
#pragma unroll
      for (int j = 0; j < 3; ++j) xo[j] = sm.st(sm.Z, j, k);
      if (k < N) { uo[2 * k] = sm.st(sm.U, 0, k); uo[2 * k + 1] = sm.st(sm.U, 1, k); }
      if (k == 0) {
        kp.T[b] = free_ ? G.T : 1.0;
        kp.obj[b] = obj;
        kp.status[b] = status;
        kp.iters[b] = iters;
      }
    }
  }
  // OBCA duals of the thread's (obstacle, stage) block to `lo` [N+1,R] / `mo` [N+1,4*n_obs] of this instance
  OB_HD void store_blocks(int tid, const BlockRegs<EMAX>& br, double* lo, double* mo) const {
    if (is_block(tid)) {
      const int i = br.i, k = br.k, r0 = br.r0, E = br.E;
      double* l = lo + k * sm.R + r0;
#pragma unroll
      for (int j = 0; j < EMAX; ++j)
        if (j < E) l[j] = br.lam[j];
      double* m = mo + k * 4 * no + 4 * i;
#pragma unroll
      for (int q = 0; q < 4; ++q) m[q] = br.mu[q];
    }
  }
};

// ======================================================================================================
// The interior-point loop of one instance.  `Exec` provides:
//   par(f)            run f(tid, BlockRegs&, part*) for every thread of the block, then a block barrier
//   reduce<S0,NS,M0,NM,N0,NN>(scratch)  one block reduction over the threads' part[] slots: sum of slots S0..S0+NS-1,
//                     max of M0.., min of N0.. (results in ex.red[] at the same slots)
//   all(f)            run f(tid) on every thread of the block, then a block barrier (no per-thread state)
//   sweep(f)          like stage(f) with the lane's SweepRegs: f(lane, SweepRegs&)
//   sweep_stages(S, mu, dw)  for s = N-1 .. 0: the three sub-steps (SweepOps) on the stage warp, a warp barrier after each
//   stage(f)          run f(lane) on the 32 lanes of the stage warp, then a warp barrier (no block barrier)
//   stage_end()       block barrier closing a run of stage() calls
//   trace(...), tick(i)  per-iteration / per-phase hooks (no-ops unless profiling)
//   once(f)           run f() on one thread (block-uniform shared state), visible after the next barrier
// ======================================================================================================
// start point of attempt a for a context whose first choice is `base` (OBCA_INIT_RETRY, include/obca_b200.h):
// ZERO -> WARM -> XREF, XREF -> WARM -> ZERO, WARM -> XREF -> ZERO, GUESS -> WARM -> XREF -> ZERO; -1 = no more
OB_HD int retry_init(int base, int a) {
  if (base == OBCA_INIT_GUESS) return a == 0 ? OBCA_INIT_GUESS : (a == 1 ? OBCA_INIT_WARM : (a == 2 ? OBCA_INIT_XREF : (a == 3 ? OBCA_INIT_ZERO : -1)));
  if (a > 2) return -1;
  return a == 0 ? base : (a == 1 ? (base == OBCA_INIT_WARM ? OBCA_INIT_XREF : OBCA_INIT_WARM)
                                 : (base == OBCA_INIT_ZERO ? OBCA_INIT_XREF : OBCA_INIT_ZERO));
}
OB_HD bool failed_search(int st) { return st == OBCA_ST_LSFAIL || st == OBCA_ST_REGFAIL || st == OBCA_ST_STALL; }
// outcomes after which another start point may still succeed (the problems are non-convex)
OB_HD bool failed_attempt(int st) { return failed_search(st) || st == OBCA_ST_INFEASIBLE || st == OBCA_ST_RESTOFAIL; }
// Recovery sequence after a failed attempt (include/obca_b200.h, OBCA_INIT_SOFT / OBCA_INIT_RETRY): up to n soft
// restarts from the point reached (multipliers, slacks, barrier parameter and filter start afresh), then the next
// start point.  `seq` packs (start point index << 4 | soft restarts used); returns the next start code or -1.
OB_HD int next_attempt(int init_word, int& seq) {
  const int nsoft = OBCA_SOFT_RESTARTS(init_word), soft = seq & 15, a = seq >> 4;
  if (soft < nsoft) { seq += 1; return OBCA_INIT_KEEP; }
  if (init_word & OBCA_INIT_RETRY) {
    const int base = ((init_word & 15) == OBCA_INIT_GUESS) ? OBCA_INIT_GUESS : (init_word & 15) % 3;
    const int nx = retry_init(base, a + 1);
    if (nx >= 0) { seq = (a + 1) << 4; return nx; }
  }
  return -1;
}

// One pass of the interior-point method.  RESTO = false: the NLP itself, from the start point G.init.  RESTO = true: the
// feasibility-restoration problem at the point the previous pass reached (see solve_with_recovery), a separate
// instantiation so that the ordinary pass - the hot loop - carries none of its code or state.
template <int EMAX, bool RESTO, class Exec>
OB_HD int solve_pass(const Solver<EMAX>& S, Exec& ex, size_t inst, double* wd_buf, int& iters_out, double& obj_out) {
  const obca_params& P = S.P;
  Glob& G = *S.sm.G;
  const double s_max = 100.0, kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
  const double dw_first = 1e-4, dw_min = 1e-20, dw_max = 1e20, kw_plus_first = 100.0, kw_plus = 8.0, kw_minus = 1.0 / 3.0;
  const double dc_min = 1e-8, lm_cap = 1e4, stall_alpha = 1e-3;
  const int stall_iters = 10;
  const double g_th = 1e-5, g_ph = 1e-8, s_th = 1.1, s_ph = 2.3, eta_ph = 1e-8;
  const double tol = P.tol;
  const int N = S.N, no = S.no;
  (void)theta_mu;
  const bool free_ = S.free_, has_term = S.has_term;
  const int m_eq = 3 * N + (free_ ? 3 : 0) + 2 * no * (N + 1);
  // the bounds n >= 0 (and pt, nt >= 0) count as inequalities of the restoration problem
  const int q_in = 12 * N + (free_ ? 2 : 0) + (has_term ? 3 : 0) + (S.sm.R + 6 * no) * (N + 1) +
                   (RESTO ? 4 * N + (has_term ? 3 : 0) + S.nb + (free_ ? 6 : 0) : 0);
  typedef BlockRegs<EMAX> BR;

  double mu = P.mu_init;
  if constexpr (RESTO) {
    mu = G.mu0;
    ex.par([&](int tid, BR& br, double* part) { (void)br; (void)part; S.resto_ref(tid); });
    ex.once([&]() { S.resto_rollout(); G.zeta = sqrt(mu); });
    ex.stage_end();
    ex.par([&](int tid, BR& br, double* part) { (void)part; S.resto_init(tid, br, mu); });
  } else {
  ex.par([&](int tid, BR& br, double* part) {
    br.i = (tid < S.nb) ? tid / S.S1 : 0; br.k = (tid < S.nb) ? tid % S.S1 : 0;
    br.r0 = S.kp.eptr[br.i]; br.E = S.kp.eptr[br.i + 1] - br.r0;
    S.start_a(tid, part); S.fill_tables(tid);
  });
  ex.template reduce<0, 1, 0, 0, 0, 0>(S.sm.SCR_H);
  const double T0 = S.start_T(ex.red[0]);
  ex.par([&](int tid, BR& br, double* part) { (void)part; S.start_b(tid, br, T0); });
  ex.par([&](int tid, BR& br, double* part) { (void)part; S.init_slacks(tid, br); });
  }

  int f_n = 0, f_wr = 0;
  bool f_active = false;
  int nstall = 0, acc_count = 0, iter = 0, status = OBCA_ST_MAXITER;
  double E0 = 0.0, fcur = 0.0;
  ex.once([&]() { G.c_best_E0 = 1e300; G.c_best_f = 0.0; G.c_dw_last = 0.0; });   // visible after the barriers of the start phase
  // watchdog (IPOPT: watchdog_shortened_iter_trigger = 10, watchdog_trial_iter_max = 3): after 10 consecutive shortened
  // steps a rejected full step is taken on trust from a checkpointed reference iterate; if 3 further full steps reach no
  // point acceptable to the reference, the reference is restored and ordinary backtracking resumes there.  This is what
  // ends the Maratos-type crawl (hundreds of 2^-9 steps) that otherwise dominates the tail of a batch.
  const int WD_TRIGGER = 10, WD_MAX = 3, ACC_STALL = 10;
  bool have_best = false;
  double e_min = 1e300;
  int e_min_iter = 0, best_lvl = 0;
  int in_wd = 0, wd_count = 0, wd_block = 0, n_short = 0;
  double rs_best = 1e300;   // restoration: lowest violation of the original problem so far, and when it was reached
  int rs_iter = 0;
  double th_last = 0.0, cmax_last = 0.0;

  ex.tick(0);
  for (;;) {
    ex.align();   // (device, first-pass kernel: rendezvous of the instances that share a block; nothing elsewhere)
    // ---- assemble
    ex.par([&](int tid, BR& br, double* part) {
      if (S.is_block(tid)) S.template assemble_block<RESTO>(tid, br, part);
      else if (S.is_stage(tid)) S.template assemble_stage<RESTO>(S.stage_lane(tid), br, part);
      else { for (int q = 0; q < NPART; ++q) part[q] = (q == PN_SZMIN) ? 1e300 : 0.0; if constexpr (RESTO) part[PS_THO] = 0.0; }
    });
    ex.tick(1);
    ex.par([&](int tid, BR& br, double* part) { (void)br; if (S.is_stage(tid)) S.assemble_combine(S.stage_lane(tid), part); });
    ex.tick(2);
    ex.template reduce<PS_F, 6, PM_E1, 5, PN_SZMIN, 1>(S.sm.SCR_D);
    ex.tick(3);
    const double Ef = ex.red[PS_F], Eth = ex.red[PS_TH], ElgS = ex.red[PS_LG], Esumy = ex.red[PS_SUMY], Esumz = ex.red[PS_SUMZ];
    double Ee1 = ex.red[PM_E1];
    if (free_) Ee1 = fmax(Ee1, fabs(ex.red[PS_GT]));
    double th_orig = 0.0;
    if constexpr (RESTO) {
      Ee1 = fmax(Ee1, ex.red[PM_CT]);   // dual infeasibility of the relaxation variables (PM_CT is theirs in this pass)
      const double Eth_ = Eth;
      ex.template reduce<PS_THO, 1, 0, 0, 0, 0>(S.sm.SCR_D);   // (SCR_H is the live Hessian here)
      th_orig = Eth_ + ex.red[PS_THO];
    }
    th_last = Eth; cmax_last = ex.red[PM_E2];
    const double Ee2 = ex.red[PM_E2], Eszmax = ex.red[PM_SZMAX], Ectmax = ex.red[PM_CT], Eszmin = ex.red[PN_SZMIN];
    const bool Eok = !(ex.red[PM_BAD] > 0.0);
    fcur = Ef;
    if (!Eok) { status = OBCA_ST_REGFAIL; break; }
    const double sd = fmax(s_max, (Esumy + Esumz) / (m_eq + q_in)) / s_max, sc = fmax(s_max, Esumz / q_in) / s_max;
    E0 = fmax(fmax(Ee1 / sd, Ee2), Eszmax / sc);
    // noise-floor level (status 2): IPOPT's scaled error AND the unscaled dual infeasibility within 1e3 acceptable_tol
    // (diverging multipliers make the scaled error small at points that are not stationary)
    const bool floor_err = fmax(E0, Ee1) <= 1e3 * P.acceptable_tol;
    if constexpr (RESTO) {
      // the restoration pass ends as soon as the violation of the original problem is below kappa times what it has to
      // beat (IPOPT: 0.9 plus acceptance by the original filter - here that filter starts afresh, so more is asked);
      // if instead its own problem converges, or the violation stops decreasing (1 % in RESTO_STALL iterations) while
      // its own constraints hold, the point is a local minimiser of the violation (the dual error of the regularised
      // steps sits on a noise floor and would never reach tol)
      if (th_orig <= fmax(RESTO_KAPPA * G.th_ref, RESTO_FEAS_TOL)) { status = OBCA_ST_OK; break; }
      if (E0 <= fmax(tol, RESTO_TOL)) { status = OBCA_ST_INFEASIBLE; break; }
      if (th_orig < 0.99 * rs_best) { rs_best = th_orig; rs_iter = iter; }
      if (iter - rs_iter >= RESTO_STALL && Eth <= 1e-6 * fmax(1.0, th_orig)) { status = OBCA_ST_INFEASIBLE; break; }
      if (iter >= RESTO_MAXITER) { status = OBCA_ST_RESTOFAIL; break; }
    } else {
    if (E0 <= tol) { status = OBCA_ST_OK; break; }
    if (E0 <= P.acceptable_tol) {
      if (++acc_count >= P.acceptable_iter) { status = OBCA_ST_ACCEPTABLE; break; }
    } else
      acc_count = 0;
    }
    // stall at the acceptable level: an acceptable point is stored, the barrier parameter is final and the error has not
    // halved for ACC_STALL iterations - the iterate wanders on the noise floor (objective constant to 10 digits).  End
    // like IPOPT does when it cannot progress from an acceptable point: with the stored point.  This, with the
    // watchdog, bounds the iteration tail of a batch (cfg 3: max 363 -> 84 -> 51 on the oracle).
    if (have_best && mu <= tol / 10 * (1 + 1e-12) && iter - e_min_iter >= ACC_STALL) { status = OBCA_ST_LSFAIL; break; }
    if (iter >= P.max_iter) { status = OBCA_ST_MAXITER; break; }
    // ---- barrier update (monotone Fiacco-McCormick)
    bool changed = false;
    for (;;) {
      const double e3 = fmax(Eszmax - mu, mu - Eszmin) / sc;
      const double Emu = fmax(fmax(Ee1 / sd, Ee2), e3);
      if (Emu <= kappa_eps * mu && mu > tol / 10) {
        mu = fmax(tol / 10, fmin(kappa_mu * mu, mu * sqrt(mu)));   // mu^theta_mu, theta_mu = 1.5
        changed = true;
      } else
        break;
    }
    // acceptable level: IPOPT's acceptable tolerance, or - at the final barrier parameter - primal feasible to 1e-6,
    // complementary, with only the dual infeasibility above tol (the rounding-noise floor of a degenerate vertex of the
    // OBCA dual polytope; same condition as at_floor below).  Judged after the barrier update: the iteration that lowers
    // mu to its final value already counts (otherwise a point that is left again one noisy step later is never stored)
    const int acc_lvl = RESTO ? 0 : (E0 <= P.acceptable_tol) ? OBCA_ST_ACCEPTABLE : (mu <= 1e-6 && Eth <= 1e-6 && floor_err) ? OBCA_ST_FLOOR : 0;
    if (acc_lvl) {
      // IPOPT stores the best acceptable iterate and ends there ("Solved To Acceptable Level") if the run fails later
      // on.  The store goes straight to the result arrays: no on-chip copy is kept.
      if (E0 < 0.1 * G.c_best_E0) {   // a store per decade of improvement keeps the HBM writes near the algorithmic figure
        ex.par([&](int tid, BR& br, double* part) { (void)part; S.store(tid, br, inst, acc_lvl, iter, Ef); });
        ex.once([&]() { G.c_best_E0 = E0; G.c_best_f = Ef; });   // after the barrier: everyone has evaluated the test
        have_best = true; best_lvl = acc_lvl;
      }
      if (E0 < 0.5 * e_min) { e_min = E0; e_min_iter = iter; }
    }
    if (changed && f_active) { f_n = 0; f_wr = 0; }
    if (changed) in_wd = 0;   // a new barrier problem: the current point becomes an ordinary iterate
    const double th = Eth, ph0 = Ef - mu * ElgS;
    const double tau = fmax(tau_min, 1 - mu);
    const double dc = free_ ? fmax(dc_min, Ectmax / lm_cap) : 0.0;
    ex.tick(4);
    // ---- Riccati with inertia correction: the pivots are the inertia test
    double dw = 0.0;
    const double dw_last = G.c_dw_last;
    bool regfail = false;
    for (;;) {
      ex.once([&]() { G.bad = 0; });
      ex.stage_end();
      ex.sweep([&](int t, SweepRegs& sr) { S.sweep_load(t, sr); S.template ric_terminal<RESTO>(t, mu, dw, dc); });
      ex.sweep_stages(S, mu, dw);
      ex.stage([&](int t) { S.ric_finish(t); });
      ex.stage_end();
      if (!G.bad) break;
      if (dw == 0.0) dw = (dw_last == 0.0) ? dw_first : fmax(dw_min, kw_minus * dw_last);
      else dw = dw * ((dw_last == 0.0) ? kw_plus_first : kw_plus);
      if (dw > dw_max) { regfail = true; break; }
      ex.stage_end();   // everyone has read G.bad before it is cleared again
    }
    if (regfail) { status = OBCA_ST_REGFAIL; break; }
    if (dw > 0) ex.once([&]() { G.c_dw_last = dw; });   // read again only after the barriers of the roll-out
    ex.tick(5);
    // ---- roll-out
    ex.stage([&](int lane) { if (lane <= N) S.fwd_prep(lane); });
    for (int s = 0; s < N; ++s) ex.stage([&](int lane) { S.fwd_step(lane, s); });
    ex.stage([&](int lane) { if (lane <= N) S.template fwd_post<RESTO>(lane, dc, mu); });
    ex.stage_end();
    ex.tick(6);
    // ---- steps of the duals / slacks, fraction to the boundary
    ex.par([&](int tid, BR& br, double* part) {
      if (S.is_block(tid)) S.template backsub_block<RESTO>(tid, br, mu, tau, part);
      else if (S.is_stage(tid)) S.template backsub_stage<RESTO>(S.stage_lane(tid), br, mu, tau, part);
      else { part[QS_DPHI] = 0.0; part[QN_AMAX] = 1.0; part[QN_AZ] = 1.0; }
    });
    ex.tick(7);
    ex.template reduce<QS_DPHI, 1, 0, 0, QN_AMAX, 2>(S.sm.SCR_H);
    ex.tick(8);
    const double Dphi = ex.red[QS_DPHI], a_max = ex.red[QN_AMAX], a_z = ex.red[QN_AZ];
    double thmax, thmin;
    if (!f_active) {
      thmax = 1e4 * fmax(1.0, th); thmin = 1e-4 * fmax(1.0, th);
      ex.once([&]() { G.c_thmax = thmax; G.c_thmin = thmin; });
      f_active = true; f_n = 0; f_wr = 0;
    } else {
      thmax = G.c_thmax; thmin = G.c_thmin;
    }
    // the two powers of the switching condition are loop invariants of the line search
    const double pw_th = (th > 0) ? pow(th, s_th) : 0.0, pw_dphi = (Dphi < 0) ? pow(-Dphi, s_ph) : 0.0;
    double a_min;
    if (Dphi < 0 && th <= thmin) a_min = fmin(g_th, fmin(g_ph * th / (-Dphi), (th > 0) ? pw_th / pw_dphi : g_th));
    else if (Dphi < 0) a_min = fmin(g_th, g_ph * th / (-Dphi));
    else a_min = g_th;
    a_min *= 0.05;
    // ---- filter line search
    double a = a_max;
    int accepted = 0;
    bool restored = false, first = true;
    // acceptance of a trial against a reference (th_r, ph_r, Dphi_r; powers pre-computed) reached with step a_r:
    // 0 rejected, 1 sufficient decrease (the reference enters the filter), 2 Armijo on the barrier function
    auto accept_test = [&](double tht, double pht, double th_r, double ph_r, double dphi_r, double a_r, double pwt, double pwd) -> int {
      if (!(isfinite(pht) && tht < thmax)) return 0;
      for (int q = 0; q < f_n; ++q)
        if (tht >= G.fth[q] && pht >= G.fph[q]) return 0;
      const bool sw = (dphi_r < 0) && (a_r * pwd > pwt);
      if (th_r <= thmin && sw) return (pht <= ph_r + eta_ph * a_r * dphi_r + 10 * 2.220446049250313e-16 * fabs(ph_r)) ? 2 : 0;
      return (tht <= (1 - g_th) * th_r || pht <= ph_r - g_ph * th_r) ? 1 : 0;
    };
    ex.tick(9);
    while (a >= a_min * (1 - 1e-12)) {
      ex.par([&](int tid, BR& br, double* part) {
        if (S.is_block(tid)) S.template trial_block<RESTO>(tid, br, a, part);
        else if (S.is_stage(tid)) S.template trial_stage<RESTO>(S.stage_lane(tid), a, part);
        else { part[0] = part[1] = part[2] = 0.0; }
      });
      ex.tick(10);
      ex.template reduce<0, 3, 0, 0, 0, 0>(S.sm.SCR_H);
      const double tht = ex.red[1], pht = ex.red[0] - mu * ex.red[2];
      if (in_wd) {
        // watchdog: only full steps, judged against the reference iterate
        const double wd_th = G.c_wd_th, wd_ph = G.c_wd_ph;
        accepted = accept_test(tht, pht, wd_th, wd_ph, G.c_wd_dphi, G.c_wd_alpha, G.c_wd_pw_th, G.c_wd_pw_dphi);
        if (accepted) {
          in_wd = 0;
          if (accepted == 1) {   // the reference point enters the filter
            const int slot = (f_n < FILT_MAX) ? f_n++ : (f_wr % FILT_MAX);
            ex.once([&]() { G.fth[slot] = (1 - g_th) * wd_th; G.fph[slot] = wd_ph - g_ph * wd_th; });
            f_wr++;
          }
          accepted = 3;
        } else if (++wd_count >= WD_MAX) {
          ex.par([&](int tid, BR& br, double* part) { (void)part; S.wd_restore(tid, br, wd_buf); });
          in_wd = 0; wd_block = 1; restored = true;   // give up: back to the reference iterate
        } else
          accepted = 3;                               // one more full step on trust
        break;
      }
      accepted = accept_test(tht, pht, th, ph0, Dphi, a, pw_th, pw_dphi);
      if (accepted) break;
      if (!RESTO && first && !wd_block && n_short >= WD_TRIGGER && isfinite(pht)) {   // (no watchdog in the restoration pass)
        ex.par([&](int tid, BR& br, double* part) { (void)part; S.wd_save(tid, br, wd_buf); });
        ex.once([&]() { G.c_wd_th = th; G.c_wd_ph = ph0; G.c_wd_dphi = Dphi; G.c_wd_alpha = a; G.c_wd_pw_th = pw_th; G.c_wd_pw_dphi = pw_dphi; G.c_wd_cmax = Ee2; });
        in_wd = 1; wd_count = 0; accepted = 3;
        break;
      }
      first = false;
      a *= 0.5;
    }
    if (restored) { iter++; continue; }
    // IPOPT ends with Solved_To_Acceptable_Level when it cannot progress from an acceptable point; the second
    // clause is the rounding-noise floor of a degenerate vertex of the OBCA dual polytope: barrier parameter at most
    // 1e-6, primal feasible to 1e-6, only the dual infeasibility (non-unique multipliers, block elimination in fp64)
    // above tol.  Every cfg-3 instance that ends here has its objective constant to 11 digits over the last ten steps.
    const int at_floor = RESTO ? 0 : (E0 <= P.acceptable_tol) ? OBCA_ST_ACCEPTABLE : (mu <= 1e-6 && th <= 1e-6 && floor_err) ? OBCA_ST_FLOOR : 0;
    ex.tick(11);
    ex.trace(iter, RESTO ? th_orig : Ef, th, E0, mu, dw, accepted ? a : -1.0);
    if (!accepted) { status = at_floor ? at_floor : OBCA_ST_LSFAIL; break; }
    nstall = (a < stall_alpha) ? nstall + 1 : 0;
    if (nstall >= stall_iters) { status = at_floor ? at_floor : OBCA_ST_STALL; break; }
    if (accepted != 3) { wd_block = 0; n_short = (a < a_max) ? n_short + 1 : 0; }
    else if (!in_wd) n_short = 0;   // watchdog succeeded
    if (accepted == 1) {
      const int slot = (f_n < FILT_MAX) ? f_n++ : (f_wr % FILT_MAX);
      ex.once([&]() { G.fth[slot] = (1 - g_th) * th; G.fph[slot] = ph0 - g_ph * th; });
      f_wr++;
    }
    ex.par([&](int tid, BR& br, double* part) { (void)part; S.template update<RESTO>(tid, br, a, a_z, mu); });
    ex.tick(12);
    iter++;
  }
  ex.tick(13);
  iters_out = iter;
  if (status < 0 && G.c_best_E0 < 1e300) {
    // x, u, lam, mu, T of the stored acceptable point are already in the result arrays
    ex.once([&]() { S.kp.status[inst] = best_lvl; });   // obj and iters: the caller (totals over the attempts)
    obj_out = G.c_best_f;
    return OBCA_ST_STORED;
  }
  if (!RESTO && status < 0 && in_wd) {
    // failed on a step taken on trust: the point reached is the watchdog's reference
    ex.par([&](int tid, BR& br, double* part) { (void)part; S.wd_restore(tid, br, wd_buf); });
    th_last = G.c_wd_th; cmax_last = G.c_wd_cmax;
  }
  ex.once([&]() { G.c_th_end = th_last; G.c_cmax_end = cmax_last; G.c_mu_end = mu; });
  ex.stage_end();
  obj_out = fcur;
  return status;
}

// One attempt = the interior-point pass and, where its line search / regularisation / progress fails, IPOPT's remedy:
// the feasibility-restoration phase from the point reached - the same algorithm on
//     min  rho sum(n) + rho sum(pt + nt) + zeta/2 |D_R ((z,u,T) - (z,u,T)_R)|^2
//     s.t. dynamics, OBCA equalities, sign / norm / input / acceleration / T rows as they are,
//          d_i(X) + n_i - S_i = 0, n_i >= 0 for the state box, the terminal set and the OBCA distance rows,
//          z_N - r_N - pt + nt = 0, pt, nt >= 0 (free-time modes)
// (IPOPT's restoration problem with the rows that can always be satisfied kept hard, so that the structure of the
// linear algebra is that of the NLP) - then the NLP again from the restored point with multipliers, slacks, barrier
// parameter and filter afresh.  Where the restoration phase cannot reduce the violation (a local minimiser of the
// violation - e.g. a predicted pose inside an obstacle, where the OBCA distance has no gradient - or its own line
// search fails) the fresh start is taken from the point of failure instead.  At most RESTO_ROUNDS (2) rounds, within the
// iteration budget; every call has to get below RESTO_KAPPA times the lowest violation seen so far.
template <int EMAX, class Exec>
OB_HD int solve_attempt(const Solver<EMAX>& S, Exec& ex, size_t inst, double* wd_buf, double* fail_buf, int& iters, double& obj, int budget) {
  typedef BlockRegs<EMAX> BR;
  Glob& G = *S.sm.G;
  int it_a = 0;
  int st = solve_pass<EMAX, false>(S, ex, inst, wd_buf, it_a, obj);
  iters += it_a;
  const bool use_resto = !(S.P.init & OBCA_INIT_NORESTO);
  double th_goal = 1e300;
  bool infeasible = false;
  for (int nres = 0; use_resto && failed_search(st) && nres < RESTO_ROUNDS && iters < budget; ++nres) {
    // (block-uniform: written before the last barrier of the pass)
    const double th_orig = G.c_th_end, cmax = G.c_cmax_end, mu_end = G.c_mu_end;
    // called at an almost feasible point (IPOPT aborts there): nothing to restore, only the fresh start below
    if (th_orig > RESTO_FEAS_TOL) {
      ex.par([&](int tid, BR& br, double* part) { (void)part; S.wd_save(tid, br, fail_buf); });
      th_goal = fmin(th_goal, th_orig);
      ex.once([&]() { G.th_ref = th_goal; G.mu0 = fmax(mu_end, cmax); });
      ex.stage_end();
      double obj_r = 0.0;
      const int st_r = solve_pass<EMAX, true>(S, ex, inst, wd_buf, it_a, obj_r);
      iters += it_a;
      if (st_r < 0) ex.par([&](int tid, BR& br, double* part) { (void)part; S.wd_restore(tid, br, fail_buf); });
      else th_goal *= RESTO_KAPPA;
      infeasible = (st_r == OBCA_ST_INFEASIBLE);
    }
    ex.once([&]() { G.init = OBCA_INIT_KEEP; });
    ex.stage_end();
    st = solve_pass<EMAX, false>(S, ex, inst, wd_buf, it_a, obj);
    iters += it_a;
  }
  if (failed_search(st) && infeasible) st = OBCA_ST_INFEASIBLE;
  return st;
}

// The instance: attempts from the start points / soft restarts the caller's flags allow (cold path: one attempt
// unless it failed).  Returns the final status (OBCA_ST_STORED: the result arrays already hold the answer).
template <int EMAX, class Exec>
OB_HD int solve_with_recovery(const Solver<EMAX>& S, Exec& ex, size_t inst, double* wd_buf, double* fail_buf, int& iters, double& obj) {
  Glob& G = *S.sm.G;
  int status;
  iters = 0;
  int budget = OBCA_RECOVERY_BUDGET;
  for (int seq = 0;;) {
    status = solve_attempt(S, ex, inst, wd_buf, fail_buf, iters, obj, budget);
    if (!failed_attempt(status) || iters >= budget) break;
    const int next = next_attempt(S.P.init, seq);
    if (next < 0) break;
    if ((S.P.init & OBCA_INIT_PATIENT) && next != OBCA_INIT_KEEP) budget = iters + OBCA_RECOVERY_BUDGET;   // per start point
    ex.stage_end();
    ex.once([&]() { G.init = next; });
    ex.stage_end();
  }
  return status;
}

// does a failed first pass have anything to follow (restoration phase, soft restart, another start point)?
OB_HD bool recovery_follows(int init_word, int st) {
  int seq = 0;
  return failed_search(st) && (!(init_word & OBCA_INIT_NORESTO) || next_attempt(init_word, seq) >= 0);
}

}  // namespace obca

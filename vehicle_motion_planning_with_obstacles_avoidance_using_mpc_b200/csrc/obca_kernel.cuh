// obca_kernel.cuh - device execution model (one CTA per instance) and the kernel entry of the batched OBCA-MPC solver.
// Included by obca_variant.cu, which is compiled once per kernel variant (one translation unit each, so the variants
// build in parallel); obca_b200.cu (the C-ABI) only sees the variants' host-side handles.
#pragma once
#include <cuda_runtime.h>

#include "obca_cta.cuh"

namespace obca {

// Optional in-kernel phase timing (-DOBCA_PROFILE; tools/phase_profile.py): cycles per phase summed over blocks.
// (counters: KParams::prof, 48 words - phase cycles (16) | par-body cycles of warp 0 (16) | of the stage warp (16))

// Block reduction, two stages through shared memory.  `buf` holds nt rows (one per slot) of one value per thread
// (row stride T).  Stage A: 8 threads per
// slot each fold T/8 consecutive values; stage B: one thread per slot folds the 8 partials into RED[slot].
// Slots [0, ns) are sums, [ns, ns+nm) maxima, the rest minima.  (Inlined: as a real call it cost ~5 k cycles per
// reduction in caller-saved register traffic - the block threads carry their iterate in registers.)
__device__ __forceinline__ void cta_reduce(const double* buf, double* RED, int T, int tid, int ns, int nm, int nt) {
  const int L = T >> 3;
  double* P2 = RED + NPART;
  for (int j = tid; j < nt * 8; j += T) {
    const int q = j >> 3, seg = j & 7;
    const double* row = buf + q * T + seg * L;
    // the 8 segments of a slot start a multiple of 32 words apart: start each at a different offset (rotation) so
    // that the lanes of a warp hit different banks
    double a = row[seg];              // seg < 8 <= L
    for (int i = 1; i < L; ++i) {
      int idx = seg + i;
      if (idx >= L) idx -= L;
      const double b = row[idx];
      a = (q < ns) ? a + b : (q < ns + nm) ? fmax(a, b) : fmin(a, b);
    }
    P2[j] = a;
  }
  __syncthreads();
  if (tid < nt) {
    double a = P2[tid * 8];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const double b = P2[tid * 8 + i];
      a = (tid < ns) ? a + b : (tid < ns + nm) ? fmax(a, b) : fmin(a, b);
    }
    RED[tid] = a;
  }
  __syncthreads();
}

// Execution model of solve_instance() on the device: one CTA, registers for the per-thread state
template <int EMAX>
struct DevExec {
  BlockRegs<EMAX> br;
  double part[NPART_X];
  double* red;   // block-reduced values (shared memory), valid after reduce()
  int tid, lane, warp, nwarps;
  bool stage_warp;
#ifdef OBCA_PROFILE
#define OBCA_P_TICK
#define OBCA_P_PAR
#define OBCA_P_SWEEP
#endif
#if defined(OBCA_P_TICK) || defined(OBCA_P_PAR) || defined(OBCA_P_SWEEP)
  long long prof[16], prof_t;
  long long work[16];   // cycles this warp spent inside par() bodies of the current phase group (before the barrier)
  int phase;
#endif
#ifdef OBCA_P_PAR
  template <class F> __device__ __forceinline__ void par(F&& f) {
    const long long t0 = clock64();
    f(tid, br, part);
    work[phase] += clock64() - t0;
    __syncthreads();
  }
#else
  // Home of the per-thread state.  A dynamically indexed member makes this object addressable, so ptxas keeps it in
  // (L1-resident) local memory and loads what a phase needs at its start instead of holding the block registers (80)
  // live across phases that do not touch them - the sweep, the control code, the reductions - and spilling at random
  // inside the hot loops: 1.0 KB of spill stores per thread instead of 2.3 KB, 15-20 % more throughput.  (Found by
  // accident: the -DOBCA_PROFILE build, whose phase timers are such a member, was the faster one.)
  int phase_hits[4];
  int phase_id;
  template <class F> __device__ __forceinline__ void par(F&& f) {
    f(tid, br, part);
    phase_hits[phase_id & 3] += 1;
    __syncthreads();
  }
#endif
  template <class F> __device__ __forceinline__ void all(F&& f) { f(tid); __syncthreads(); }
  SweepRegs sr;
  template <class F> __device__ __forceinline__ void sweep(F&& f) {
#ifdef OBCA_P_SWEEP
    if (stage_warp) { const long long t0 = clock64(); f(lane, sr); __syncwarp(); work[15] += clock64() - t0; work[14] += 1; }
#else
    if (stage_warp) { f(lane, sr); __syncwarp(); }
#endif
  }
  template <class F> __device__ __forceinline__ void stage(F&& f) {
    if (stage_warp) { f(lane); __syncwarp(); }
  }
  __device__ __forceinline__ void stage_end() { __syncthreads(); }
  template <class F> __device__ __forceinline__ void once(F&& f) { if (tid == 0) f(); }
  __device__ __forceinline__ void trace(int, double, double, double, double, double, double) {}
  __device__ __forceinline__ void tick(int i) {
#ifdef OBCA_P_TICK
    long long t = clock64(); prof[i] += t - prof_t; prof_t = t;
    phase = (i + 1) & 15;
#else
    (void)i;
#endif
  }
  // one block reduction: sums of part[S0..], maxima of part[M0..], minima of part[N0..] -> red[] (same slots).
  // Slot ranges must be laid out S | M | N consecutively in part[] (they are: see the PS_/PM_/PN_ enums).
  template <int S0, int NS, int M0, int NM, int N0, int NN> __device__ __forceinline__ void reduce(double* scratch) {
    const int T = 32 * nwarps, rs = T, pos = tid;
#pragma unroll
    for (int q = 0; q < NS; ++q) scratch[q * rs + pos] = part[S0 + q];
#pragma unroll
    for (int q = 0; q < NM; ++q) scratch[(NS + q) * rs + pos] = part[M0 + q];
#pragma unroll
    for (int q = 0; q < NN; ++q) scratch[(NS + NM + q) * rs + pos] = part[N0 + q];
    __syncthreads();
    // results land at red[slot] = RED[slot]: shift the base so that RED[0] is slot S0 (or M0 / N0 when NS == 0)
    constexpr int first = (NS > 0) ? S0 : ((NM > 0) ? M0 : N0);
    cta_reduce(scratch, red + first, T, tid, NS, NM, NS + NM + NN);
  }
};

extern __shared__ double obca_smem[];

// NT/NOT/RT > 0: kernel specialised for horizon NT, NOT obstacles, RT half-space rows (sizes are literals);
// 0: generic kernel, sizes read from the parameter block.
// FULL = false: the first-pass kernel - one interior-point pass per instance, nothing else, so that the hot loop is the
// whole kernel (with the restoration pass compiled into the same kernel the hot loop lost 10 % to a 50 % larger stack
// frame); an instance whose pass fails and that may recover is appended to kp.fail_list instead of being stored.
// FULL = true: the recovery kernel - the complete sequence (pass, restoration phase, fresh starts, other start points)
// over that list, started from scratch per instance (the first pass is deterministic, so the sequence is the one a
// single kernel would run; failures are rare, and the list spreads them over all blocks instead of leaving them as the
// tail of the block that met them).
template <int EMAX, int MAXT, int MINB, int NT = 0, int NOT = 0, int RT = 0, bool FULL = false>
__global__ void __launch_bounds__(MAXT, MINB) obca_solve_kernel(const __grid_constant__ KParams kp, int nwarps_rt, int has_uref) {
  __shared__ unsigned int s_inst;
  Sm sm;
  constexpr bool fixed = NT > 0;
  const int nwarps = fixed ? (NOT * (NT + 1) + 31) / 32 + 1 : nwarps_rt;
  if (fixed) sm_carve(sm, obca_smem, NT, NOT, RT, (NOT * (NT + 1) + 31) / 32 + 1, has_uref);
  else sm_carve(sm, obca_smem, kp.P.N, kp.P.n_obs, kp.P.rows, nwarps_rt, has_uref);
  const Solver<EMAX> S(kp, sm);
  DevExec<EMAX> ex;
  ex.red = sm.RED;
  ex.tid = threadIdx.x; ex.lane = threadIdx.x & 31; ex.warp = threadIdx.x >> 5; ex.nwarps = nwarps;
  ex.stage_warp = (ex.warp == nwarps - 1);
  bool first = true;
  const unsigned int n_items = kp.count_dev ? (unsigned)*kp.count_dev : (unsigned)kp.batch;
  for (;;) {
    if (threadIdx.x == 0) {
      const unsigned int w = atomicAdd(kp.counter, 1u);
      s_inst = (w < n_items) ? (kp.index ? (unsigned)kp.index[w] : w) : 0xffffffffu;
    }
    __syncthreads();
    const unsigned int inst = s_inst;
    if (inst == 0xffffffffu) break;
#if defined(OBCA_P_TICK) || defined(OBCA_P_PAR) || defined(OBCA_P_SWEEP)
    for (int i = 0; i < 16; ++i) { ex.prof[i] = 0; ex.work[i] = 0; }
    ex.prof_t = clock64(); ex.phase = 0;
#endif
    S.load(ex.tid, inst, first || !kp.shared_obs);
    first = false;
    __syncthreads();
#if !defined(OBCA_P_PAR)
    ex.phase_id = (int)(inst & 3u);
    for (int i = 0; i < 4; ++i) ex.phase_hits[i] = 0;
#endif
    int iters = 0;
    double obj = 0.0;
    double* const ckpt = kp.wd_buf + (size_t)blockIdx.x * 2 * kp.wd_stride;   // watchdog reference | point of failure
    int status;
    if constexpr (FULL) status = solve_with_recovery(S, ex, (size_t)inst, ckpt, ckpt + kp.wd_stride, iters, obj);
    else status = solve_pass<EMAX, false>(S, ex, (size_t)inst, ckpt, iters, obj);
    if (!FULL && kp.fail_list && recovery_follows(kp.P.init, status)) {   // (block-uniform)
      if (ex.tid == 0) kp.fail_list[atomicAdd(kp.fail_count, 1u)] = (int32_t)inst;
    } else if (status != OBCA_ST_STORED) S.store(ex.tid, ex.br, inst, status, iters, obj);
    else if (ex.tid == 0) { kp.obj[inst] = obj; kp.iters[inst] = iters; }
#if !defined(OBCA_P_PAR)
    if (ex.phase_hits[ex.phase_id & 3] < 0) kp.iters[inst] = -1;   // never true: keeps the member alive
#endif
#if defined(OBCA_P_TICK) || defined(OBCA_P_PAR) || defined(OBCA_P_SWEEP)
    if (ex.tid == 0 && kp.prof)
      for (int i = 0; i < 16; ++i) { atomicAdd(&kp.prof[i], (unsigned long long)ex.prof[i]); atomicAdd(&kp.prof[16 + i], (unsigned long long)ex.work[i]); }
    if (ex.stage_warp && ex.lane == 0 && kp.prof)
      for (int i = 0; i < 16; ++i) atomicAdd(&kp.prof[32 + i], (unsigned long long)ex.work[i]);
#endif
    __syncthreads();
  }
}

}  // namespace obca

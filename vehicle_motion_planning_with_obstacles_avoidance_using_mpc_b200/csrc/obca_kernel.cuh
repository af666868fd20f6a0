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
// Barrier of one GROUP of threads.  A first-pass block hosts G instances side by side (G groups of T threads, each with
// its own shared-memory region and its own named barrier 1 + group); the recovery kernel and G = 1 use barrier 0.
template <int G>
struct GroupBar {
  int id, n;   // barrier number, threads taking part
  __device__ __forceinline__ void sync() const {
    if constexpr (G == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
  }
  __device__ __forceinline__ bool sync_and(bool p) const {
    if constexpr (G == 1) return __syncthreads_and(p);
    unsigned r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r) : "r"((unsigned)p), "r"(id), "r"(n) : "memory");
    return r != 0;
  }
};

template <class GB>
__device__ __forceinline__ void cta_reduce(const double* buf, double* RED, int T, int tid, int ns, int nm, int nt, const GB& gb) {
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
  gb.sync();
  if (tid < nt) {
    double a = P2[tid * 8];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const double b = P2[tid * 8 + i];
      a = (tid < ns) ? a + b : (tid < ns + nm) ? fmax(a, b) : fmin(a, b);
    }
    RED[tid] = a;
  }
  gb.sync();
}

// Execution model of solve_instance() on the device: one CTA, registers for the per-thread state
template <int EMAX, int G, bool LOCKSTEP>
struct DevExec {
  BlockRegs<EMAX> br;
  double part[NPART_X];
  double* red;   // block-reduced values (shared memory), valid after reduce()
  int tid, lane, warp, nwarps;   // (within the group)
  bool stage_warp;
  GroupBar<G> gb;
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
    gb.sync();
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
    gb.sync();
  }
#endif
  template <class F> __device__ __forceinline__ void all(F&& f) { f(tid); gb.sync(); }
  SweepRegs sr;
  template <class F> __device__ __forceinline__ void sweep(F&& f) {
#ifdef OBCA_P_SWEEP
    if (stage_warp) { const long long t0 = clock64(); f(lane, sr); __syncwarp(); work[15] += clock64() - t0; work[14] += 1; }
#else
    if (stage_warp) { f(lane, sr); __syncwarp(); }
#endif
  }
  // the stages of the Riccati sweep: one warp-uniform branch (the other warps go straight to the block barrier that
  // follows) around one call per lane
  template <class SolverT> __device__ __forceinline__ void sweep_stages(const SolverT& S, double mu, double dw) {
    if (!stage_warp) return;
#ifdef OBCA_P_SWEEP
    const long long t0 = clock64();
#endif
#ifndef OBCA_SWEEP_DIRECT
    c_sweep_fn((uint32_t)__cvta_generic_to_shared(S.sm.Z), sr, lane, S.N, mu, dw);
#else
    sweep_lane_dev((uint32_t)__cvta_generic_to_shared(S.sm.Z), sr, lane, S.N, mu, dw);
#endif
#ifdef OBCA_P_SWEEP
    work[15] += clock64() - t0; work[14] += 3 * S.N;
#endif
  }
  template <class F> __device__ __forceinline__ void stage(F&& f) {
    if (stage_warp) { f(lane); __syncwarp(); }
  }
  __device__ __forceinline__ void stage_end() { gb.sync(); }
  // Top of an interior-point iteration.  The groups of a block run the same code on different instances; left alone
  // they drift apart and the SM's instruction caches (32 KB against ~150 KB of straight-line code per iteration) serve
  // three different streams - measured: 3.3 stall cycles per issue waiting for instructions against 1.3 with one block
  // per SM.  Meeting once per iteration keeps the groups within a few hundred instructions of each other.  The
  // rendezvous counts the threads of groups that have run out of work (they keep arriving until everybody has).
  __device__ __forceinline__ void align() { if constexpr (LOCKSTEP && G > 1) __syncthreads_count(0); }
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
    gb.sync();
    // results land at red[slot] = RED[slot]: shift the base so that RED[0] is slot S0 (or M0 / N0 when NS == 0)
    constexpr int first = (NS > 0) ? S0 : ((NM > 0) ? M0 : N0);
    cta_reduce(scratch, red + first, T, tid, NS, NM, NS + NM + NN, gb);
  }
};

extern __shared__ __align__(16) double obca_smem[];

// ======================================================================================================
// Bulk-copy staging (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS).  The inputs of an instance are ten small arrays
// (0.9 KB at the headline shape): while one instance is being solved the copy engine fetches the inputs of the next one
// into the prefetch buffer Sm::PF, and the OBCA duals of a finished instance - 5.4 KB that the threads would otherwise
// write as scattered 32-byte pieces - leave through a shared-memory tile as two bulk stores.
// Bulk copies move multiples of 16 bytes between 16-byte aligned addresses; rows of an [B, n] double array start on an
// odd double for every other instance, so an array is fetched as  [head double] + 16-byte aligned middle + [tail double]
// with the middle by bulk copy and the (at most two) end doubles by ordinary loads, shifted by one double inside its
// slot when the source starts on an odd double.
// ======================================================================================================
namespace bulk {
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(void* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(saddr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s2g_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one input array: n doubles from src into slot `dst` (16-byte aligned).  Returns the bytes the bulk copy will deliver.
// Element i ends up at dst[shift + i], shift = 1 if src starts on an odd double.
__device__ __forceinline__ unsigned fetch(double* dst, const double* src, int n, void* bar, bool issue) {
  if (!src || n <= 0) return 0;
  const int shift = (int)(((uintptr_t)src >> 3) & 1);
  const int tail = (n - shift) & 1;                    // a double left over after the 16-byte units
  const int mid = n - shift - tail;                    // doubles moved by the bulk copy
  if (!issue) return (unsigned)mid * 8u;
  if (shift) dst[1] = src[0];
  if (tail) dst[shift + n - 1] = src[n - 1];
  if (mid > 0) g2s(dst + 2 * shift, src + shift, (unsigned)mid * 8u, bar);
  return (unsigned)mid * 8u;
}
__device__ __forceinline__ const double* fetched(const double* slot, const double* src) {
  return src ? slot + (((uintptr_t)src >> 3) & 1) : nullptr;
}
}  // namespace bulk

// Prefetch of the inputs of instance b by the stage warp: lane j < 10 owns input array j.
template <int EMAX>
__device__ __forceinline__ void prefetch_inputs(const Solver<EMAX>& S, const Sm& sm, size_t b, int lane, bool with_obs) {
  const typename Solver<EMAX>::InstPtrs q = S.inst_ptrs(b);
  const int N = sm.N, R = sm.R, hu = sm.has_uref;
  const double* src[10] = {q.xref, q.uref, with_obs ? q.A : nullptr, with_obs ? q.b0 : nullptr, with_obs ? q.db : nullptr,
                           q.x0, q.u0, q.Tmax, q.Ts, q.term};
  const double* my = nullptr;
  int slot = PF_XREF, n = 0;
#pragma unroll
  for (int j = 0; j < 10; ++j)
    if (lane == j) { my = src[j]; slot = PF_XREF + j; }
  if (lane < 10) n = pf_len(slot, N, R, hu);
  double* dst = sm.PF + pf_off(slot, N, R, hu);
  void* bar = sm.PF;   // slot PF_MBAR
  const unsigned mine = (lane < 10) ? bulk::fetch(dst, my, n, bar, false) : 0u;
  const unsigned total = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) bulk::mbar_expect(bar, total);   // the one arrival of this phase + the bytes the copies will deliver
  __syncwarp();
  if (lane < 10) bulk::fetch(dst, my, n, bar, true);
}

// NT/NOT/RT > 0: kernel specialised for horizon NT, NOT obstacles, RT half-space rows (sizes are literals);
// 0: generic kernel, sizes read from the parameter block.
// FULL = false: the first-pass kernel - one interior-point pass per instance, nothing else, so that the hot loop is the
// whole kernel (with the restoration pass compiled into the same kernel the hot loop lost 10 % to a 50 % larger stack
// frame); an instance whose pass fails and that may recover is appended to kp.fail_list instead of being stored.
// FULL = true: the recovery kernel - the complete sequence (pass, restoration phase, fresh starts, other start points)
// over that list, started from scratch per instance (the first pass is deterministic, so the sequence is the one a
// single kernel would run).  It runs twice per solve: as ONE block on an SM the first-pass launch leaves free, polling
// the list while the first pass is still running (kp.poll; failures are rare but long - 100 to 300 iterations - and
// would otherwise be a serial tail after the launch), and then over the whole device for what is left.
template <int EMAX, int MAXT, int MINB, int NT = 0, int NOT = 0, int RT = 0, bool FULL = false>
__global__ void __launch_bounds__(MAXT * MINB, 1)
obca_solve_kernel(const __grid_constant__ KParams kp, int nwarps_rt, int has_uref) {
  // G = MINB instances side by side in one block per SM; in the first-pass kernel the groups run in lockstep (see
  // DevExec::align), in the recovery kernel they are independent
  constexpr int G = MINB;
  __shared__ unsigned int s_inst_g[G];
  Sm sm;
  constexpr bool fixed = NT > 0;
  const int nwarps = fixed ? (NOT * (NT + 1) + 31) / 32 + 1 : nwarps_rt;
  const int Tg = 32 * nwarps;                       // threads per group
  const int gid = (G == 1) ? 0 : (int)threadIdx.x / Tg;
  const int ltid = (int)threadIdx.x - gid * Tg;
  double* const base = obca_smem + (size_t)gid * kp.smem_stride;
  if (fixed) sm_carve(sm, base, NT, NOT, RT, (NOT * (NT + 1) + 31) / 32 + 1, has_uref);
  else sm_carve(sm, base, kp.P.N, kp.P.n_obs, kp.P.rows, nwarps_rt, has_uref);
  unsigned int& s_inst = s_inst_g[gid];
  const Solver<EMAX> S(kp, sm);
  DevExec<EMAX, G, !FULL> ex;
  ex.red = sm.RED;
  ex.tid = ltid; ex.lane = ltid & 31; ex.warp = ltid >> 5; ex.nwarps = nwarps;
  ex.stage_warp = (ex.warp == nwarps - 1);
  ex.gb.id = (G == 1) ? 0 : 1 + gid; ex.gb.n = Tg;
  bool first = true;
  const unsigned int n_items = kp.count_dev ? (unsigned)*kp.count_dev : (unsigned)kp.batch;
#ifndef OBCA_NO_BULK
  if (ltid == 0) bulk::mbar_init(sm.PF, 1);
  unsigned pf_parity = 0;
  unsigned int next = 0xfffffffeu;   // 0xfffffffe: no item claimed ahead; 0xffffffff: the queue is empty
  ex.gb.sync();
#endif
  for (;;) {
#ifdef OBCA_NO_BULK
    if (ltid == 0) {
      const unsigned int w = atomicAdd(kp.counter, 1u);
      s_inst = (w < n_items) ? (kp.index ? (unsigned)kp.index[w] : w) : 0xffffffffu;
    }
    ex.gb.sync();
    const unsigned int inst = s_inst;
    if (inst == 0xffffffffu) break;
#else
    // work item: the one claimed (and prefetched) during the previous solve, else claim now and fetch
    unsigned int inst = next;
    if (inst == 0xfffffffeu) {
      if (ltid == 0) {
        if (FULL && kp.poll) {
          // consumer of a list that is still being written: claim entry c only once it exists (compare-and-swap, no
          // overshoot), wait for its value to be published, leave when the producer has finished and nothing is unclaimed
          // The producer must be RUNNING for this to make sense: under a profiler or sanitizer kernels are serialised and
          // the first pass would only start after this block has left - so the block also leaves when the producer's
          // work counter (kp.heartbeat) has not moved for 4 ms; the launch over the whole device that follows takes over.
          unsigned int got = 0xffffffffu;
          unsigned int beat = *(volatile unsigned int*)kp.heartbeat;
          unsigned long long t_beat;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_beat));
          for (;;) {
            if (*(volatile unsigned int*)kp.poll) break;      // the first pass has finished: the whole device takes over
            const unsigned int n = *(volatile unsigned int*)kp.count_dev, c = *(volatile unsigned int*)kp.counter;
            if (c < n) {
              if (atomicCAS(kp.counter, c, c + 1u) == c) {
                int v;
                while ((v = *(volatile int32_t*)(kp.index + c)) < 0) __nanosleep(200);
                got = (unsigned)v;
                break;
              }
              continue;
            }
            const unsigned int b2 = *(volatile unsigned int*)kp.heartbeat;
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (b2 != beat) { beat = b2; t_beat = t; }
            else if (t - t_beat > 4000000ull) break;
            __nanosleep(1000);
          }
          s_inst = got;
        } else {
          const unsigned int w = atomicAdd(kp.counter, 1u);
          const unsigned int n = (FULL && kp.count_dev) ? *(volatile unsigned int*)kp.count_dev : n_items;
          s_inst = (w < n) ? (kp.index ? (unsigned)kp.index[w] : w) : 0xffffffffu;
        }
      }
      ex.gb.sync();
      inst = s_inst;
      if (inst != 0xffffffffu && ex.stage_warp) prefetch_inputs(S, sm, inst, ex.lane, first || !kp.shared_obs);
    }
    if (inst == 0xffffffffu) break;
#endif
#if defined(OBCA_P_TICK) || defined(OBCA_P_PAR) || defined(OBCA_P_SWEEP)
    for (int i = 0; i < 16; ++i) { ex.prof[i] = 0; ex.work[i] = 0; }
    ex.prof_t = clock64(); ex.phase = 0;
#endif
#ifdef OBCA_NO_BULK
    S.load(ex.tid, inst, first || !kp.shared_obs);
    first = false;
    ex.gb.sync();
#else
    {
      // wait for the bulk copies of this instance's inputs (bounded: a copy that never lands must not hang the GPU -
      // the plain loads take over and the event is counted), then move them from the prefetch buffer to their places
      bool landed = false;
      for (int spin = 0; spin < (1 << 20) && !landed; ++spin) landed = bulk::mbar_try(sm.PF, pf_parity);
      pf_parity ^= 1u;
      landed = ex.gb.sync_and(landed);
      const bool with_obs = first || !kp.shared_obs;
      if (landed) {
        const int N_ = sm.N, R_ = sm.R, hu = sm.has_uref;
        const typename Solver<EMAX>::InstPtrs g = S.inst_ptrs(inst);
        typename Solver<EMAX>::InstPtrs q;
        auto at = [&](int slot, const double* src) { return bulk::fetched(sm.PF + pf_off(slot, N_, R_, hu), src); };
        q.xref = at(PF_XREF, g.xref); q.uref = at(PF_UREF, g.uref);
        q.A = at(PF_A, g.A); q.b0 = at(PF_B0, g.b0); q.db = at(PF_DB, g.db);
        q.x0 = at(PF_X0, g.x0); q.u0 = at(PF_U0, g.u0); q.Tmax = at(PF_TMAX, g.Tmax); q.Ts = at(PF_TS, g.Ts); q.term = at(PF_TERM, g.term);
        S.load_from(ex.tid, inst, q, with_obs);
      } else {
        if (ex.tid == 0 && kp.bulk_timeouts) atomicAdd(kp.bulk_timeouts, 1u);
        S.load(ex.tid, inst, with_obs);
      }
      first = false;
      ex.gb.sync();
      // claim the next item now and let the copy engine fetch its inputs under this solve - but only while the queue
      // is long: the last items are claimed when a group is free, so that the end of the batch stays balanced
      if (ltid == 0) {
        unsigned int nx = 0xfffffffeu;
        if (!FULL && landed && *(volatile unsigned int*)kp.counter + 2u * gridDim.x * G < n_items) {
          const unsigned int w = atomicAdd(kp.counter, 1u);
          nx = (w < n_items) ? (kp.index ? (unsigned)kp.index[w] : w) : 0xffffffffu;
        }
        s_inst = nx;
      }
      ex.gb.sync();
      next = s_inst;
      if (next < 0xfffffffeu && ex.stage_warp) prefetch_inputs(S, sm, next, ex.lane, !kp.shared_obs);
    }
#endif
#if !defined(OBCA_P_PAR)
    ex.phase_id = (int)(inst & 3u);
    for (int i = 0; i < 4; ++i) ex.phase_hits[i] = 0;
#endif
    int iters = 0;
    double obj = 0.0;
    double* const ckpt = kp.wd_buf + ((size_t)(kp.wd_block0 + blockIdx.x) * G + gid) * 2 * kp.wd_stride;   // watchdog reference | point of failure
    int status;
    if constexpr (FULL) status = solve_with_recovery(S, ex, (size_t)inst, ckpt, ckpt + kp.wd_stride, iters, obj);
    else status = solve_pass<EMAX, false>(S, ex, (size_t)inst, ckpt, iters, obj);
    if (!FULL && kp.fail_list && recovery_follows(kp.P.init, status)) {   // (group-uniform)
      if (ex.tid == 0) {   // (the entry is published after the count: consumers wait for a value >= 0)
        *(volatile int32_t*)(kp.fail_list + atomicAdd(kp.fail_count, 1u)) = (int32_t)inst;
        __threadfence();
      }
    } else if (status != OBCA_ST_STORED) {
#ifdef OBCA_NO_BULK
      S.store(ex.tid, ex.br, inst, status, iters, obj);
#else
      // results: the small per-stage part straight to HBM; the duals (and inputs) through a shared-memory tile laid out
      // like the result arrays (in the dual-block scratch, dead now) and out as bulk stores - or, where the instance's
      // rows are not 16-byte aligned, as a coalesced copy of the tile
      const int S1_ = sm.S1, R_ = sm.R, no_ = sm.no, N_ = sm.N;
      double* tile = (double*)(((uintptr_t)sm.ETA + 15) & ~(uintptr_t)15);
      const int n_lam = S1_ * R_, n_mu = S1_ * 4 * no_, n_u = 2 * N_;
      double* t_lam = tile; double* t_mu = tile + ((n_lam + 1) & ~1); double* t_u = t_mu + n_mu;
      double* g_lam = kp.lam + (size_t)inst * n_lam; double* g_mu = kp.mu + (size_t)inst * n_mu; double* g_u = kp.u + (size_t)inst * n_u;
      S.store_stage(ex.tid, inst, status, iters, obj, t_u);
      S.store_blocks(ex.tid, ex.br, t_lam, t_mu);
      bulk::fence_async_smem();
      ex.gb.sync();
      double* const tl[3] = {t_lam, t_mu, t_u}; double* const gl[3] = {g_lam, g_mu, g_u}; const int nn[3] = {n_lam, n_mu, n_u};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const bool can = (((uintptr_t)gl[j] & 15) == 0) && ((nn[j] & 1) == 0) && nn[j] > 0;
        if (can) { if (ex.tid == 0) bulk::s2g(gl[j], tl[j], (unsigned)nn[j] * 8u); }
        else for (int i = ex.tid; i < nn[j]; i += sm.T) gl[j][i] = tl[j][i];
      }
      if (ex.tid == 0) bulk::s2g_commit_wait_read();
#endif
    } else if (ex.tid == 0) { kp.obj[inst] = obj; kp.iters[inst] = iters; }
#if !defined(OBCA_P_PAR)
    if (ex.phase_hits[ex.phase_id & 3] < 0) kp.iters[inst] = -1;   // never true: keeps the member alive
#endif
#if defined(OBCA_P_TICK) || defined(OBCA_P_PAR) || defined(OBCA_P_SWEEP)
    if (ex.tid == 0 && kp.prof)
      for (int i = 0; i < 16; ++i) { atomicAdd(&kp.prof[i], (unsigned long long)ex.prof[i]); atomicAdd(&kp.prof[16 + i], (unsigned long long)ex.work[i]); }
    if (ex.stage_warp && ex.lane == 0 && kp.prof)
      for (int i = 0; i < 16; ++i) atomicAdd(&kp.prof[32 + i], (unsigned long long)ex.work[i]);
#endif
    ex.gb.sync();
  }
  // out of work: keep the block's rendezvous going until every group is (the count is taken by the barrier itself, so
  // all threads of the block see the same number in the same round and leave together)
  if constexpr (G > 1 && !FULL)
    while (__syncthreads_count(1) < (int)blockDim.x) {}
}

}  // namespace obca

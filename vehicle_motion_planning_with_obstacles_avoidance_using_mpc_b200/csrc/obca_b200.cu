// obca_b200.cu - kernel entry + C-ABI (include/obca_b200.h) of the batched OBCA-MPC solver for sm_100a.
//
// Replaces the CasADi/IPOPT work behind obca.obca_mpc4 / obca_mpc6 / obca_mpc8 / obca2 of the reference
// (src/obca.py:828-1071, 1361-1562, 1564-1758, 338-629; called at src/closed_loop.py:118,131,137,170,...).
// One thread block per NLP instance (obca_cta.cuh), persistent blocks pulling instances from a work queue.
// No host fallback: every entry point fails with OBCA_E_NODEVICE / OBCA_E_CUDA when there is no GPU.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "obca_cta.cuh"

// ======================================================================================================
// C-ABI
// ======================================================================================================
// Kernel variants: one translation unit each (obca_variant.cu compiled with -DKV_*; _lib.py lists them), so that they
// build in parallel.  A variant exports the host-side handle of its kernel.
typedef const void* kernel_fn;
extern "C" {
const void* obca_kv_cfg3(void);    // <4,128,3, N=20, 4 obstacles, 16 rows>   headline
const void* obca_kv_cfg5(void);    // <4,192,2, N=20, 6, 24>
const void* obca_kv_cfg2(void);    // <4,128,3, N=10, 2, 8>
const void* obca_kv_cfg4d(void);   // <4,128,3, N=5, 6, 18>   closed loop, obstacle detected
const void* obca_kv_cfg4f(void);   // <4,128,3, N=5, 5, 14>   closed loop, free phase
const void* obca_kv_g4_128(void);  // generic: sizes read from the parameter block
const void* obca_kv_g4_192(void);
const void* obca_kv_g4_416(void);
const void* obca_kv_g8_128(void);
const void* obca_kv_g8_416(void);
const void* obca_kv_r4_128(void);  // recovery kernels (generic sizes): the complete sequence over the failed instances
const void* obca_kv_r4_192(void);
const void* obca_kv_r4_416(void);
const void* obca_kv_r8_128(void);
const void* obca_kv_r8_416(void);
}

#define OBCA_HOST_CHUNKS 4

struct obca_ctx {
  int device;
  int max_batch;
  obca_params P;
  int emax;               // largest edge count seen at the last solve (selects the kernel variant)
  int nwarps, threads, grid;   // per instance: warps, threads; grid = instances resident on the device (blocks x groups)
  int groups, blocks;          // first-pass kernel: instances per block, blocks launched (one per SM)
  size_t smem_bytes;           // per instance
  kernel_fn fn, fn_rec;   // first-pass kernel, recovery kernel
  int32_t* fail_list;     // per launch slot: instances whose first pass failed (max_batch entries each)
  int32_t* order;         // per launch slot: longest-first work order of the first pass (max_batch entries each)
  float* score;           // ... and the difficulty estimates it is sorted by
  int32_t* rank;          // ... and their ranks
  int cfg_emax, cfg_uref; // configuration the launch geometry was computed for
  unsigned int* counter;
  double* wd_buf;         // watchdog checkpoints, one slot per resident block
  size_t wd_bytes;
  int64_t wd_stride;
  int64_t launches;
  cudaEvent_t ev0[OBCA_HOST_CHUNKS], ev1[OBCA_HOST_CHUNKS];   // around the launches of each slot
  int timed_slots;        // slots used by the last solve call (0: nothing timed yet)
  void* stage;            // host-path staging
  size_t stage_bytes;
  int slots, cfg_slots;   // launches that may be in flight at once (own work counter and checkpoint slots each)
  cudaStream_t hs[OBCA_HOST_CHUNKS];   // host path: one stream per chunk of a large batch
  cudaStream_t aux[OBCA_HOST_CHUNKS];  // per launch slot: stream of the recovery block that runs beside the first pass
  cudaEvent_t ev_fork[OBCA_HOST_CHUNKS], ev_join[OBCA_HOST_CHUNKS];
  cudaEvent_t ev_shared;
};

static int sm_count_of(int device) {
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

// kernel variant for (max edges per obstacle, threads per block); the BASELINE configurations have kernels with
// compile-time sizes
// `groups`: instances a first-pass block hosts side by side (the variant's blocks-per-SM figure: one block per SM)
static kernel_fn pick_kernel(int emax, int threads, int N, int no, int R, int* groups) {
  *groups = 3;
  if (emax <= 4 && N == 20 && no == 4 && R == 16) return obca_kv_cfg3();
  if (emax <= 4 && N == 20 && no == 6 && R == 24) { *groups = 2; return obca_kv_cfg5(); }
  if (emax <= 4 && N == 10 && no == 2 && R == 8) return obca_kv_cfg2();
  if (emax <= 4 && N == 5 && no == 6 && R == 18) return obca_kv_cfg4d();
  if (emax <= 4 && N == 5 && no == 5 && R == 14) return obca_kv_cfg4f();
  if (emax <= 4) {
    if (threads <= 128) return obca_kv_g4_128();
    if (threads <= 192) { *groups = 2; return obca_kv_g4_192(); }
    *groups = 1;
    return obca_kv_g4_416();
  }
  if (threads <= 128) { *groups = 2; return obca_kv_g8_128(); }
  *groups = 1;
  return obca_kv_g8_416();
}
static kernel_fn pick_recovery_kernel(int emax, int threads) {
  if (emax <= 4) return threads <= 128 ? obca_kv_r4_128() : (threads <= 192 ? obca_kv_r4_192() : obca_kv_r4_416());
  return threads <= 128 ? obca_kv_r8_128() : obca_kv_r8_416();
}

static int configure(obca_ctx* c, int emax, int has_uref) {
  if (c->fn && c->cfg_emax == emax && c->cfg_uref == has_uref && c->cfg_slots == c->slots) return OBCA_OK;
  const obca_params& P = c->P;
  const int nb = P.n_obs * (P.N + 1);
  c->nwarps = (nb + 31) / 32 + 1;
  c->threads = 32 * c->nwarps;
  obca::Sm sm;
  c->smem_bytes = obca::sm_carve(sm, nullptr, P.N, P.n_obs, P.rows, c->nwarps, has_uref) * sizeof(double);
  c->fn = pick_kernel(emax, c->threads, P.N, P.n_obs, P.rows, &c->groups);
  c->smem_bytes = (c->smem_bytes + 15) & ~(size_t)15;
  if (c->smem_bytes > 227 * 1024) return OBCA_E_SIZE;
  while (c->groups > 1 && c->groups * c->smem_bytes > 227 * 1024) c->groups -= 1;
  if (cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->groups * c->smem_bytes)) != cudaSuccess) {
    cudaGetLastError();
    return OBCA_E_CUDA;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c->fn, c->groups * c->threads, c->groups * c->smem_bytes) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return OBCA_E_CUDA;
  }
  c->blocks = sm_count_of(c->device);
  c->grid = c->blocks * c->groups;
  c->fn_rec = pick_recovery_kernel(emax, c->threads);
  int per_sm_rec = 0;
  if (cudaFuncSetAttribute(c->fn_rec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->groups * c->smem_bytes)) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_rec, c->fn_rec, c->groups * c->threads, c->groups * c->smem_bytes) != cudaSuccess || per_sm_rec < 1) {
    cudaGetLastError();
    return OBCA_E_CUDA;
  }
  c->wd_stride = (emax <= 4) ? obca::Solver<4>::wd_doubles(c->threads, P.N + 1) : obca::Solver<8>::wd_doubles(c->threads, P.N + 1);
  // two checkpoints per resident instance; one block more than the first pass uses: the recovery block beside it
  const size_t need = (size_t)c->slots * (c->grid + c->groups) * 2 * c->wd_stride * sizeof(double);
  if (need > c->wd_bytes) {
    if (c->wd_buf) cudaFree(c->wd_buf);
    c->wd_buf = nullptr; c->wd_bytes = 0;
    if (cudaMalloc(&c->wd_buf, need) != cudaSuccess) { cudaGetLastError(); return OBCA_E_NOMEM; }
    c->wd_bytes = need;
  }
  c->cfg_emax = emax; c->cfg_uref = has_uref; c->cfg_slots = c->slots;
  return OBCA_OK;
}

// fp64 throughput probe: 8 independent DFMA chains per thread, nothing else in the loop.  The solver is bound by the
// fp64 pipe and its latencies, not by HBM, so bench.py reports the solver's fp64 rate against THIS measured ceiling
// beside the HBM roofline the metric asks for.
__global__ void __launch_bounds__(256) obca_dfma_probe(double* out, int iters, double b, double c) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ---------------------------------------------------------------------------------------------------------------
// Longest-first work order.  A launch ends with its last instance: with the queue in the caller's order the makespan is
// ~6 % above the ideal (the longest of the last few hundred instances decides), and an instance whose first pass fails
// near the end of the queue starts its recovery when everything else has finished.  The instances are therefore handed
// out in descending order of a difficulty estimate computed from the inputs alone - how much of the reference window
// is in collision (separating-axis clearance of the ego rectangle at every reference pose against every obstacle), how
// sharply the window turns, how far the start heading is from the first reference heading.  The weights are a
// least-squares fit of the iteration count on the headline workload (correlation 0.6); the order changes nothing but
// the time (every instance is solved independently - tests compare ordered and unordered launches bit for bit).
// Measured: cfg 3, eight batches, profiles/r2_order_ab.log.  OBCA_B200_FIFO=1 keeps the caller's order.
// One warp per instance, lane = stage of the reference window (N + 1 <= 32).
__global__ void __launch_bounds__(128) obca_order_score(const obca::KParams kp, float* score) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, k = threadIdx.x & 31;
  if (b >= kp.batch) return;
  const obca_params& P = kp.P;
  const int N = P.N, no = P.n_obs, R = P.rows;
  const double* x0 = kp.x0 + (size_t)3 * b;
  const double* xr = kp.xref + (size_t)b * (N + 1) * 3;
  const size_t ob = kp.shared_obs ? 0 : (size_t)b * R;
  const double* A = kp.A + 2 * ob;
  const double* b0 = kp.b0 + ob;
  const double* db = (kp.db && kp.stacked) ? kp.db + ob : nullptr;
  const double e0 = P.ego[0], e1 = P.ego[1], e2 = P.ego[2], e3 = P.ego[3];
  const double ax[4] = {e0, e0, -e2, -e2}, ay[4] = {e1, -e3, -e3, e1};
  const bool on = k <= N;
  const int kk = on ? k : N;
  const double px = xr[3 * kk], py = xr[3 * kk + 1], th = xr[3 * kk + 2];
  double sn, cs;
  sincos(th, &sn, &cs);
  double cx[4], cy[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { cx[j] = px + cs * ax[j] - sn * ay[j]; cy[j] = py + sn * ax[j] + cs * ay[j]; }
  double Sk = 10.0;
  for (int i = 0; i < no; ++i) {
    const int r0 = kp.eptr[i], E = kp.eptr[i + 1] - r0;
    double sep = -1e30;
    for (int r = r0; r < r0 + E; ++r) {              // face normals of the obstacle
      const double a0 = A[2 * r], a1 = A[2 * r + 1], bk = b0[r] + (db ? kk * db[r] : 0.0);
      double m = a0 * cx[0] + a1 * cy[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) m = fmin(m, a0 * cx[j] + a1 * cy[j]);
      sep = fmax(sep, (m - bk) * rsqrt(a0 * a0 + a1 * a1));
    }
    if (E >= 3) {                                    // axes of the ego rectangle against the obstacle's vertices
      double ulo = 1e30, uhi = -1e30, vlo = 1e30, vhi = -1e30;
      for (int r = r0; r < r0 + E; ++r) {
        const int q = (r + 1 < r0 + E) ? r + 1 : r0;
        const double a0 = A[2 * r], a1 = A[2 * r + 1], c0_ = A[2 * q], c1_ = A[2 * q + 1];
        const double br = b0[r] + (db ? kk * db[r] : 0.0), bq = b0[q] + (db ? kk * db[q] : 0.0);
        const double det = a0 * c1_ - a1 * c0_;
        if (fabs(det) < 1e-12) continue;
        const double id = 1.0 / det;
        const double vx = (br * c1_ - a1 * bq) * id - px, vy = (a0 * bq - br * c0_) * id - py;
        const double u = vx * cs + vy * sn, v = -vx * sn + vy * cs;
        ulo = fmin(ulo, u); uhi = fmax(uhi, u); vlo = fmin(vlo, v); vhi = fmax(vhi, v);
      }
      if (uhi >= ulo) sep = fmax(sep, fmax(fmax(ulo - e0, -e2 - uhi), fmax(vlo - e1, -e3 - vhi)));
    }
    Sk = fmin(Sk, sep);
  }
  const unsigned full = 0xffffffffu;
  const int c0 = __popc(__ballot_sync(full, on && Sk < 0.0)), c05 = __popc(__ballot_sync(full, on && Sk < 0.5)),
            c15 = __popc(__ballot_sync(full, on && Sk < 1.5));
  const double th_next = __shfl_down_sync(full, th, 1);
  double d = (k < N) ? fabs(th_next - th) : 0.0, dsum = d, dmax = d, smin = on ? Sk : 10.0;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    dsum += __shfl_xor_sync(full, dsum, o);
    dmax = fmax(dmax, __shfl_xor_sync(full, dmax, o));
    smin = fmin(smin, __shfl_xor_sync(full, smin, o));
  }
  if (k == 0) {
    const double hd = fabs(x0[2] - th);
    float sc = (float)(13.41 + 0.37 * smin - 0.17 * c05 + 2.29 * c0 + 0.17 * c15 - 0.38 * dsum + 1.12 * hd + 5.34 * dmax +
                       0.03 * x0[0] + 0.35 * fmin(Sk, 3.0));
    // the ranks must be a permutation whatever the inputs hold (degenerate rows, NaN poses): a key that does not compare
    // would leave its instance out of the work list
    if (!(sc == sc)) sc = 0.0f;
    score[b] = fminf(fmaxf(sc, -1.0e30f), 1.0e30f);
  }
}
// rank by counting (stable: ties in index order), the comparisons of one instance spread over gridDim.y blocks;
// then order[rank] = instance.  n <= OBCA_ORDER_MAX.
#define OBCA_ORDER_MAX 16384
#define OBCA_ORDER_SPLIT 8
__global__ void __launch_bounds__(256) obca_order_rank(const float* score, int n, int32_t* rank) {
  __shared__ float tile[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float si = (i < n) ? score[i] : 0.0f;
  const int per = ((n + OBCA_ORDER_SPLIT - 1) / OBCA_ORDER_SPLIT + 255) & ~255;
  const int jb = blockIdx.y * per, je = (jb + per < n) ? jb + per : n;
  int r = 0;
  for (int j0 = jb; j0 < je; j0 += 256) {
    const int j = j0 + threadIdx.x;
    tile[threadIdx.x] = (j < je) ? score[j] : -3.0e38f;
    __syncthreads();
    const int m = (je - j0 < 256) ? je - j0 : 256;
#pragma unroll 8
    for (int q = 0; q < m; ++q) {
      const float sj = tile[q];
      r += (sj > si) || (sj == si && j0 + q < i);
    }
    __syncthreads();
  }
  if (i < n && r) atomicAdd(&rank[i], r);
}
__global__ void obca_order_scatter(const int32_t* rank, int n, int32_t* order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) order[rank[i]] = i;
}

extern "C" {

int obca_b200_abi_version(void) { return OBCA_B200_ABI_VERSION; }

int obca_b200_fp64_peak(int device, double* tflops) {
  if (!tflops) return OBCA_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return OBCA_E_NODEVICE; }
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return OBCA_E_CUDA;
  if (device >= ndev || cudaSetDevice(device) != cudaSuccess) return OBCA_E_ARG;
  const int blocks = sm_count_of(device) * 8, threads = 256, iters = 1 << 14;
  double* buf = nullptr;
  if (cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return OBCA_E_NOMEM; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {   // first run warms up
    cudaEventRecord(e0, 0);
    obca_dfma_probe<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); cudaFree(buf); return OBCA_E_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
  return OBCA_OK;
}

const char* obca_b200_strerror(int rc) {
  switch (rc) {
    case OBCA_OK: return "ok";
    case OBCA_E_ARG: return "invalid argument";
    case OBCA_E_NODEVICE: return "no CUDA device (this library has no host fallback)";
    case OBCA_E_CUDA: return "CUDA runtime error";
    case OBCA_E_NOMEM: return "out of device memory";
    case OBCA_E_SIZE: return "problem size outside compiled limits (N+1 <= 32, obstacles <= 12, rows <= 48, edges per obstacle <= 8, on-chip state <= 227 KB)";
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
  // per launch slot: work counter of the first pass, of the recovery, length of the list of failed instances, flag
  // 'first pass finished'; last word:
  // bulk-copy prefetches that timed out (diagnostics, obca_b200_bulk_timeouts)
  if (cudaMalloc(&c->counter, (4 * OBCA_HOST_CHUNKS + 1) * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(c->counter, 0, (4 * OBCA_HOST_CHUNKS + 1) * sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); free(c); return OBCA_E_NOMEM; }
  if (cudaMalloc(&c->fail_list, (size_t)OBCA_HOST_CHUNKS * max_batch * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&c->order, (size_t)OBCA_HOST_CHUNKS * max_batch * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&c->score, (size_t)OBCA_HOST_CHUNKS * max_batch * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&c->rank, (size_t)OBCA_HOST_CHUNKS * max_batch * sizeof(int32_t)) != cudaSuccess) {
    cudaGetLastError(); cudaFree(c->counter); if (c->fail_list) cudaFree(c->fail_list); if (c->order) cudaFree(c->order); free(c); return OBCA_E_NOMEM;
  }
  for (int j = 0; j < OBCA_HOST_CHUNKS; ++j) { cudaEventCreate(&c->ev0[j]); cudaEventCreate(&c->ev1[j]); }
  cudaEventCreateWithFlags(&c->ev_shared, cudaEventDisableTiming);
  c->slots = 1;
  *out = c;
  return OBCA_OK;
}

int obca_b200_destroy(obca_ctx* c) {
  if (!c) return OBCA_E_ARG;
  cudaSetDevice(c->device);
  cudaFree(c->counter);
  cudaFree(c->fail_list);
  cudaFree(c->order);
  cudaFree(c->score);
  cudaFree(c->rank);
  if (c->stage) cudaFree(c->stage);
  if (c->wd_buf) cudaFree(c->wd_buf);
  for (int j = 0; j < OBCA_HOST_CHUNKS; ++j) { cudaEventDestroy(c->ev0[j]); cudaEventDestroy(c->ev1[j]); }
  cudaEventDestroy(c->ev_shared);
  for (int j = 0; j < OBCA_HOST_CHUNKS; ++j) {
    if (c->hs[j]) cudaStreamDestroy(c->hs[j]);
    if (c->aux[j]) { cudaStreamDestroy(c->aux[j]); cudaEventDestroy(c->ev_fork[j]); cudaEventDestroy(c->ev_join[j]); }
  }
  free(c);
  return OBCA_OK;
}

#ifdef OBCA_PROFILE
// phase cycle counters (tick i closes phase i): 0 start 1 assemble 2 combine 3 reduce 4 control 5 riccati 6 roll-out
// 7 steps 8 reduce 9 line-search setup 10 trial 11 reduce+test 12 update 13 exit
static unsigned long long* g_prof_dev = nullptr;
static unsigned long long* prof_buffer() {
  if (!g_prof_dev && (cudaMalloc(&g_prof_dev, 48 * sizeof(unsigned long long)) != cudaSuccess ||
                      cudaMemset(g_prof_dev, 0, 48 * sizeof(unsigned long long)) != cudaSuccess)) g_prof_dev = nullptr;
  return g_prof_dev;
}
int obca_b200_prof_read(unsigned long long* out, int reset) {
  if (!prof_buffer()) return OBCA_E_CUDA;
  if (cudaMemcpy(out, g_prof_dev, 48 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return OBCA_E_CUDA;
  if (reset && cudaMemset(g_prof_dev, 0, 48 * sizeof(unsigned long long)) != cudaSuccess) return OBCA_E_CUDA;
  return OBCA_OK;
}
#endif

// device bytes held by the context: the solver keeps its whole working set on-chip, so this is only the work-queue
// counter, the watchdog checkpoint slots (one per resident block) and the staging buffer of the host entry point
int64_t obca_b200_scratch_bytes(const obca_ctx* c) {
  return c ? (int64_t)((4 * OBCA_HOST_CHUNKS + 1) * sizeof(unsigned int) + (size_t)OBCA_HOST_CHUNKS * c->max_batch * (3 * sizeof(int32_t) + sizeof(float)) +
                       c->stage_bytes + c->wd_bytes) : 0;
}
int64_t obca_b200_launch_count(const obca_ctx* c) { return c ? c->launches : 0; }

// diagnostics: bulk-copy input prefetches that did not land within the kernel's bounded wait (the plain loads took
// over).  Expected 0; synchronises the device.
int64_t obca_b200_bulk_timeouts(obca_ctx* c) {
  if (!c) return -1;
  unsigned int v = 0;
  if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(&v, c->counter + 4 * OBCA_HOST_CHUNKS, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
  return (int64_t)v;
}

// Kernel time of the last solve call.  A chunked host solve keeps several launches in flight on separate streams: its
// kernel time is the span from the start of the first chunk's launch to the end of the last one to finish.
float obca_b200_last_kernel_ms(obca_ctx* c) {
  if (!c || c->timed_slots < 1) return -1.0f;
  float best = -1.0f;
  for (int j = 0; j < c->timed_slots; ++j) {
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev0[0], c->ev1[j]) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
    if (ms > best) best = ms;
  }
  return best;
}

static int solve_slot(obca_ctx* c, int slot, int batch, const int32_t* count_dev, const int32_t* index_dev,
                            const double* x0, const double* u0, const double* xref, const double* uref,
                            const double* T_max, const double* term, const double* Ts_inst, const int32_t* edge_ptr,
                            const double* A, const double* b0, const double* db, int obstacles_shared, double* x, double* u,
                            double* lam, double* mu, double* T, double* obj, int32_t* status, int32_t* iters,
                            void* cuda_stream) {
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
  int emax = 0;
  for (int i = 0; i < P.n_obs; ++i) {
    const int E = edge_ptr[i + 1] - edge_ptr[i];
    if (E < 1) return OBCA_E_ARG;
    if (E > emax) emax = E;
  }
  if (emax > 8) return OBCA_E_SIZE;
  if (cudaSetDevice(c->device) != cudaSuccess) return OBCA_E_CUDA;
  int rc = configure(c, emax <= 4 ? 4 : 8, uref != nullptr);
  if (rc != OBCA_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  obca::KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.P = P;
  for (int i = 0; i <= P.n_obs; ++i) kp.eptr[i] = edge_ptr[i];
  kp.batch = batch; kp.shared_obs = obstacles_shared ? 1 : 0;
  kp.free_ = free_; kp.has_term = has_term; kp.stacked = (P.mode != OBCA_MODE_FREE);
  kp.x0 = x0; kp.u0 = u0; kp.xref = xref; kp.uref = uref; kp.Tmax = T_max; kp.term = term; kp.Ts_inst = Ts_inst;
  kp.A = A; kp.b0 = b0; kp.db = db;
  kp.x = x; kp.u = u; kp.lam = lam; kp.mu = mu; kp.T = T; kp.obj = obj; kp.status = status; kp.iters = iters;
  unsigned int* const cnt = c->counter + 4 * slot;   // work counter of the first pass | of the recovery | failures | done flag
  kp.counter = cnt; kp.wd_buf = c->wd_buf + (size_t)slot * (c->grid + c->groups) * 2 * c->wd_stride; kp.wd_stride = c->wd_stride;
  kp.index = index_dev; kp.count_dev = count_dev;
  const bool recover = obca::recovery_follows(P.init, OBCA_ST_LSFAIL);   // do the flags allow anything after a failed pass?
  int32_t* const fail_list = c->fail_list + (size_t)slot * c->max_batch;
  if (recover) { kp.fail_list = fail_list; kp.fail_count = cnt + 2; }
  kp.bulk_timeouts = c->counter + 4 * OBCA_HOST_CHUNKS;
#ifdef OBCA_PROFILE
  kp.prof = prof_buffer();
#endif
  kp.smem_stride = (int64_t)(c->smem_bytes / sizeof(double));
  if (cudaMemsetAsync(cnt, 0, 4 * sizeof(unsigned int), st) != cudaSuccess) return OBCA_E_CUDA;
  int nwarps = c->nwarps, has_uref = uref != nullptr;
  const dim3 block(c->groups * c->threads);
  const size_t smem = c->groups * c->smem_bytes;
  cudaError_t lerr = cudaSuccess;
  // The recovery block that runs BESIDE the first pass, on an SM that launch leaves free: it polls the list of failed
  // instances while it is being written (restoration phase, fresh starts, other start points: 100-300 iterations per
  // instance - as a tail after the launch a single failure holds the batch for 2-15 ms).  Launching it behind a first
  // pass that takes all 148 SMs (it then becomes resident in the tail of the launch) was measured over the batches of
  // eight ranks: 0.5 % faster where no first pass fails, 8-10 % slower where one or two do (mean 21.1 against 20.3 ms,
  // profiles/r2_seeds_ab.log) - the reserved SM is the default, OBCA_B200_RESERVE_SM=0 switches to the other.
  obca::KParams kr = kp;
  const bool beside = recover && c->blocks > 1;
  static const bool reserve_sm = !(getenv("OBCA_B200_RESERVE_SM") && atoi(getenv("OBCA_B200_RESERVE_SM")) == 0);
  if (recover) {
    kr.counter = cnt + 1; kr.index = fail_list; kr.count_dev = (const int32_t*)(cnt + 2);
    kr.fail_list = nullptr; kr.fail_count = nullptr;
    if (cudaMemsetAsync(fail_list, 0xff, (size_t)batch * sizeof(int32_t), st) != cudaSuccess) return OBCA_E_CUDA;
  }
  auto launch_beside = [&]() {
    obca::KParams kb = kr;
    kb.poll = cnt + 3; kb.heartbeat = cnt; kb.wd_block0 = c->blocks;
    void* args_b[3] = {&kb, &nwarps, &has_uref};
    lerr = cudaLaunchKernel(c->fn_rec, dim3(1), block, args_b, smem, c->aux[slot]);
    cudaEventRecord(c->ev_join[slot], c->aux[slot]);
    c->launches += 1;
  };
  if (beside) {
    if (!c->aux[slot]) {
      if (cudaStreamCreateWithFlags(&c->aux[slot], cudaStreamNonBlocking) != cudaSuccess) return OBCA_E_CUDA;
      cudaEventCreateWithFlags(&c->ev_fork[slot], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&c->ev_join[slot], cudaEventDisableTiming);
    }
    cudaEventRecord(c->ev_fork[slot], st);
    cudaStreamWaitEvent(c->aux[slot], c->ev_fork[slot], 0);
    if (reserve_sm) launch_beside();
  }
  const int width = (beside && reserve_sm) ? c->blocks - 1 : c->blocks;
  const int need_blocks = (batch + c->groups - 1) / c->groups;
  const int grid = width < need_blocks ? width : need_blocks;
  cudaEventRecord(c->ev0[slot], st);
  static const bool fifo = getenv("OBCA_B200_FIFO") && atoi(getenv("OBCA_B200_FIFO")) != 0;
  if (!fifo && !index_dev && !count_dev && P.n_obs > 0 && batch >= 4 * c->grid && batch <= OBCA_ORDER_MAX) {
    float* const score = c->score + (size_t)slot * c->max_batch;
    int32_t* const order = c->order + (size_t)slot * c->max_batch;
    int32_t* const rank = c->rank + (size_t)slot * c->max_batch;
    cudaMemsetAsync(rank, 0, (size_t)batch * sizeof(int32_t), st);
    obca_order_score<<<(batch + 3) / 4, 128, 0, st>>>(kp, score);
    obca_order_rank<<<dim3((batch + 255) / 256, OBCA_ORDER_SPLIT), 256, 0, st>>>(score, batch, rank);
    obca_order_scatter<<<(batch + 255) / 256, 256, 0, st>>>(rank, batch, order);
    kp.index = order;
    c->launches += 3;
  }
  void* args[3] = {&kp, &nwarps, &has_uref};
  if (lerr == cudaSuccess) lerr = cudaLaunchKernel(c->fn, dim3(grid), block, args, smem, st);
  if (lerr == cudaSuccess && beside && !reserve_sm) launch_beside();
  if (lerr == cudaSuccess && recover) {
    // first pass finished: tell the block beside it to stop claiming (it finishes the instances it holds), and let the
    // whole device take what is left of the list - the two share the work counter; the stream joins the side block last
    if (beside) cudaMemsetAsync(cnt + 3, 1, 1, st);   // (low byte = 1)
    void* args_r[3] = {&kr, &nwarps, &has_uref};
    const int grid_r = c->blocks < need_blocks ? c->blocks : need_blocks;
    lerr = cudaLaunchKernel(c->fn_rec, dim3(grid_r), block, args_r, smem, st);
    c->launches += 1;
    if (beside) cudaStreamWaitEvent(st, c->ev_join[slot], 0);
  }
  cudaEventRecord(c->ev1[slot], st);
  c->timed_slots = slot + 1 > c->timed_slots || slot == 0 ? slot + 1 : c->timed_slots;
  c->launches += 1;
  if (lerr != cudaSuccess || cudaGetLastError() != cudaSuccess) return OBCA_E_CUDA;
  return OBCA_OK;
}

int obca_b200_solve_indexed(obca_ctx* c, int batch, const int32_t* count_dev, const int32_t* index_dev, const double* x0,
                            const double* u0, const double* xref, const double* uref, const double* T_max, const double* term,
                            const double* Ts_inst, const int32_t* edge_ptr, const double* A, const double* b0, const double* db,
                            int obstacles_shared, double* x, double* u, double* lam, double* mu, double* T, double* obj,
                            int32_t* status, int32_t* iters, void* cuda_stream) {
  return solve_slot(c, 0, batch, count_dev, index_dev, x0, u0, xref, uref, T_max, term, Ts_inst, edge_ptr, A, b0, db,
                    obstacles_shared, x, u, lam, mu, T, obj, status, iters, cuda_stream);
}

int obca_b200_solve(obca_ctx* c, int batch, const double* x0, const double* u0, const double* xref, const double* uref,
                    const double* T_max, const double* term, const double* Ts_inst, const int32_t* edge_ptr, const double* A,
                    const double* b0, const double* db, int obstacles_shared, double* x, double* u, double* lam, double* mu,
                    double* T, double* obj, int32_t* status, int32_t* iters, void* cuda_stream) {
  return obca_b200_solve_indexed(c, batch, nullptr, nullptr, x0, u0, xref, uref, T_max, term, Ts_inst, edge_ptr, A, b0, db,
                                 obstacles_shared, x, u, lam, mu, T, obj, status, iters, cuda_stream);
}

int obca_b200_solve_host(obca_ctx* c, int batch, const double* x0, const double* u0, const double* xref, const double* uref,
                         const double* T_max, const double* term, const double* Ts_inst, const int32_t* edge_ptr,
                         const double* A, const double* b0, const double* db, int obstacles_shared, double* x, double* u,
                         double* lam, double* mu, double* T, double* obj, int32_t* status, int32_t* iters) {
  if (!c || batch < 0 || batch > c->max_batch) return OBCA_E_ARG;
  if (batch == 0) return OBCA_OK;
  const obca_params& P = c->P;
  if (cudaSetDevice(c->device) != cudaSuccess) return OBCA_E_CUDA;
  const size_t B = batch, N = P.N, R = P.rows, no = P.n_obs, Bo = obstacles_shared ? 1 : B;
  // one staging buffer: inputs then outputs, all 8-byte aligned
  const size_t n_in[10] = {B * 3, B * 2, B * (N + 1) * 3, uref ? B * N * 2 : 0, T_max ? B : 0, term ? B * 3 : 0,
                           Bo * R * 2, Bo * R, db ? Bo * R : 0, Ts_inst ? B : 0};
  const double* h_in[10] = {x0, u0, xref, uref, T_max, term, A, b0, db, Ts_inst};
  const size_t n_out[6] = {B * (N + 1) * 3, B * N * 2, B * (N + 1) * R, B * (N + 1) * 4 * no, B, B};
  double* h_out[6] = {x, u, lam, mu, T, obj};
  size_t tot = 0;
  for (int i = 0; i < 10; ++i) tot += n_in[i];
  for (int i = 0; i < 6; ++i) tot += n_out[i];
  const size_t bytes = tot * sizeof(double) + 2 * B * sizeof(int32_t);
  if (bytes > c->stage_bytes) {
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_bytes = 0;
    if (cudaMalloc(&c->stage, bytes) != cudaSuccess) { cudaGetLastError(); return OBCA_E_NOMEM; }
    c->stage_bytes = bytes;
  }
  double* d = (double*)c->stage;
  double* d_in[10];
  double* d_out[6];
  for (int i = 0; i < 10; ++i) {
    d_in[i] = n_in[i] ? d : nullptr;
    if (n_in[i] && !h_in[i]) return OBCA_E_ARG;
    d += n_in[i];
  }
  for (int i = 0; i < 6; ++i) {
    if (!h_out[i]) return OBCA_E_ARG;
    d_out[i] = d; d += n_out[i];
  }
  if (!status || !iters) return OBCA_E_ARG;
  int32_t* d_status = (int32_t*)d;
  int32_t* d_iters = d_status + B;
  // A large batch goes through in chunks on separate streams: the copy-back of one chunk overlaps the solve of the next
  // and the last blocks of one launch overlap the first of the next.  Small batches take one launch on the default stream.
  size_t out_doubles = 0;
  for (int i = 0; i < 6; ++i) out_doubles += n_out[i];
  const int chunks = (B >= 2048 && out_doubles * sizeof(double) >= (8u << 20)) ? OBCA_HOST_CHUNKS : 1;
  if (chunks > 1) {
    c->slots = OBCA_HOST_CHUNKS;
    for (int j = 0; j < chunks; ++j)
      if (!c->hs[j] && cudaStreamCreateWithFlags(&c->hs[j], cudaStreamNonBlocking) != cudaSuccess) return OBCA_E_CUDA;
  }
  // per-instance element counts (0: absent, or shared by the batch)
  const size_t per_in[10] = {3, 2, (N + 1) * 3, uref ? N * 2 : 0, T_max ? 1u : 0u, term ? 3u : 0u,
                             obstacles_shared ? 0 : R * 2, obstacles_shared ? 0 : R, (db && !obstacles_shared) ? R : 0,
                             Ts_inst ? 1u : 0u};
  const size_t per_out[6] = {(N + 1) * 3, N * 2, (N + 1) * R, (N + 1) * 4 * no, 1, 1};
  cudaStream_t s0 = chunks > 1 ? c->hs[0] : (cudaStream_t)0;
  if (obstacles_shared) {
    for (int i = 6; i <= 8; ++i)
      if (n_in[i] && cudaMemcpyAsync(d_in[i], h_in[i], n_in[i] * sizeof(double), cudaMemcpyHostToDevice, s0) != cudaSuccess) return OBCA_E_CUDA;
    if (chunks > 1 && cudaEventRecord(c->ev_shared, s0) != cudaSuccess) return OBCA_E_CUDA;
  }
  for (int j = 0; j < chunks; ++j) {
    const size_t lo = B * j / chunks, hi = B * (j + 1) / chunks, nb = hi - lo;
    if (nb == 0) continue;
    cudaStream_t st = chunks > 1 ? c->hs[j] : (cudaStream_t)0;
    if (chunks > 1 && j > 0 && obstacles_shared && cudaStreamWaitEvent(st, c->ev_shared, 0) != cudaSuccess) return OBCA_E_CUDA;
    const double* di[10];
    for (int i = 0; i < 10; ++i) {
      di[i] = d_in[i] ? d_in[i] + lo * per_in[i] : nullptr;
      if (per_in[i] && cudaMemcpyAsync(d_in[i] + lo * per_in[i], h_in[i] + lo * per_in[i], nb * per_in[i] * sizeof(double),
                                       cudaMemcpyHostToDevice, st) != cudaSuccess) return OBCA_E_CUDA;
    }
    int rc = solve_slot(c, j, (int)nb, nullptr, nullptr, di[0], di[1], di[2], di[3], di[4], di[5], di[9], edge_ptr, di[6], di[7], di[8],
                        obstacles_shared, d_out[0] + lo * per_out[0], d_out[1] + lo * per_out[1], d_out[2] + lo * per_out[2],
                        d_out[3] + lo * per_out[3], d_out[4] + lo, d_out[5] + lo, d_status + lo, d_iters + lo, st);
    if (rc != OBCA_OK) return rc;
    for (int i = 0; i < 6; ++i)
      if (cudaMemcpyAsync(h_out[i] + lo * per_out[i], d_out[i] + lo * per_out[i], nb * per_out[i] * sizeof(double),
                          cudaMemcpyDeviceToHost, st) != cudaSuccess) return OBCA_E_CUDA;
    if (cudaMemcpyAsync(status + lo, d_status + lo, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) return OBCA_E_CUDA;
    if (cudaMemcpyAsync(iters + lo, d_iters + lo, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) return OBCA_E_CUDA;
  }
  for (int j = 0; j < chunks; ++j)
    if (cudaStreamSynchronize(chunks > 1 ? c->hs[j] : (cudaStream_t)0) != cudaSuccess) return OBCA_E_CUDA;
  return OBCA_OK;
}

}  // extern "C"

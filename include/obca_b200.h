/* obca_b200.h - C-ABI of the B200-native batched OBCA-MPC solver.
 *
 * The reference (tg623623nana/Vehicle_Motion_Planning_with_Obstacles_Avoidance_using_MPC) has no FFI of
 * its own: its de-facto boundary is the Python method call  self.obca_solver.<method>(...)  on
 * `class obca` (src/obca.py:10, constructed at src/closed_loop.py:22).  Each entry point below replaces
 * the CasADi/IPOPT work behind those methods:
 *
 *   obca_b200_solve  mode 0  FREE          obca.obca_mpc4   src/obca.py:828-1071  (closed_loop.py:118,382)
 *                    mode 1  FIXED_SET     obca.obca_mpc6   src/obca.py:1361-1562 (closed_loop.py:131,269,389)
 *                    mode 2  FIXED_NOTERM  obca.obca_mpc8   src/obca.py:1564-1758 (closed_loop.py:137,275,395)
 *                    mode 3  FREE_STACKED  obca.obca2 fixtime=0  src/obca.py:338-629 (closed_loop.py:170,263)
 *                    mode 4  FIXED_OBCA2   obca.obca2 fixtime=1  (terminal set optional, obca.py:518-521)
 *
 * One call solves `batch` independent NLPs (a group of threads each, the groups of an SM in one thread block; large
 * batches are handed out longest-first by a difficulty estimate - the order never changes a result).  All arrays are float64, C-contiguous,
 * batch-major; the caller owns every buffer, the library owns only the context.  Functions return 0 on
 * success and a negative code on argument / CUDA errors (obca_b200_strerror); they never throw and never
 * exit.  The per-instance solver outcome is in `status` (>= 0  <=>  the reference's feas == True).
 * There is no host fallback: without a CUDA device obca_b200_create fails with OBCA_E_NODEVICE.
 */
#ifndef OBCA_B200_H
#define OBCA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OBCA_B200_ABI_VERSION 2

enum { OBCA_MODE_FREE = 0, OBCA_MODE_FIXED_SET = 1, OBCA_MODE_FIXED_NOTERM = 2, OBCA_MODE_FREE_STACKED = 3,
       OBCA_MODE_FIXED_OBCA2 = 4 };
/* start point: 0 = the reference's (every Opti variable 0, Topt = 1: obca.py:856), 1 = poses from xref,
 * 2 = A* warm start (poses from xref, T from arc length, inputs by differences, duals from the most
 * separating face) */
enum { OBCA_INIT_ZERO = 0, OBCA_INIT_XREF = 1, OBCA_INIT_WARM = 2 };
/* 4 = the caller's own guess: the poses are read from the OUTPUT array x [B,N+1,3] before it is overwritten and the
 * rest of the start point is built from them as for OBCA_INIT_WARM (a route from a global planner where xref is only
 * start and goal - closed_loop.py:113-120 -, or the previous plan shifted by one step).  With OBCA_INIT_RETRY the
 * other start points follow as for WARM. */
#define OBCA_INIT_GUESS 4
/* An instance whose line search, regularisation or progress fails (status -4, -2, -5) enters the feasibility-
 * restoration phase, as in IPOPT (up to two rounds of: minimise the violation of the state box, terminal and OBCA
 * distance rows from the point reached, then the NLP again from the restored point).
 * OR-ed into `init`: if the attempt still fails (or ends at a local minimiser of the violation, -6 / -7) the instance
 * is restarted from the other start points (WARM -> XREF -> ZERO, XREF -> WARM -> ZERO, ZERO -> WARM -> XREF) before the
 * failure is reported; `iters` is the total over the attempts.  The problems are non-convex: a restart may end in a
 * different local solution. */
#define OBCA_INIT_RETRY 16
/* OR-ed into `init`: OBCA_INIT_SOFT(n), n <= 15.  After a failure of the same kind the solver first keeps the primal
 * point it reached and starts again from there with fresh multipliers (y = 0, z = 1), slacks (max(d(x), bound_push)),
 * barrier parameter and filter - at most n times per start point.  (Round 1's stand-in for the restoration phase;
 * kept for callers that switch the phase off.)
 * OBCA_INIT_KEEP is that start code (internal: contexts are created with ZERO, XREF or WARM). */
#define OBCA_INIT_KEEP 3
/* OR-ed into `init`: switch the feasibility-restoration phase off (a failed line search is then reported at once, or
 * handed to the restart rules above) */
#define OBCA_INIT_NORESTO 32
/* OR-ed into `init`: the iteration budget below counts per start point instead of per instance (single solves, where
 * a long tail does not hold a batch) */
#define OBCA_INIT_PATIENT 64
#define OBCA_INIT_SOFT(n) (((n) & 15) << 8)
/* no further pass is started once the passes of an instance add up to this many iterations (recovered instances of the
 * closed-loop workload need 80 at the median and 216 at most; an instance that fails all twelve passes would run 350-800) */
#define OBCA_RECOVERY_BUDGET 300
#define OBCA_SOFT_RESTARTS(init) (((init) >> 8) & 15)
/* per-instance status */
/* 0: optimality error <= tol.  1: IPOPT's "Solved To Acceptable Level" (error <= acceptable_tol for acceptable_iter
 * iterations, or the run could not progress from such a point).  2: the run ended on the rounding-noise floor of a
 * degenerate vertex of the OBCA dual polytope - primal infeasibility <= 1e-6, barrier parameter <= 1e-6, and only the
 * dual infeasibility (IPOPT's scaled AND the unscaled one) above acceptable_tol, at most 1e3 * acceptable_tol (1e-3 for
 * mpc4, 1e-5 for mpc6/8); the objective is constant to ~10 digits there and a multiplier fit on the active set shows
 * these points stationary to ~1e-7 (profiles/r2_parity_report.json), but the solver's own error estimate is looser
 * than anything IPOPT reports as success, hence its own code.  feas <=> status >= 0. */
enum { OBCA_ST_OK = 0, OBCA_ST_ACCEPTABLE = 1, OBCA_ST_FLOOR = 2, OBCA_ST_MAXITER = -1, OBCA_ST_REGFAIL = -2, OBCA_ST_EMPTYBOX = -3,
       OBCA_ST_LSFAIL = -4, OBCA_ST_STALL = -5,
       OBCA_ST_INFEASIBLE = -6,   /* the restoration phase converged to a local minimiser of the constraint violation
                                     that is infeasible (IPOPT: "Converged to a point of local infeasibility")       */
       OBCA_ST_RESTOFAIL = -7 };  /* the restoration phase itself failed (IPOPT: "Restoration Failed")               */
/* return codes */
enum { OBCA_OK = 0, OBCA_E_ARG = -1, OBCA_E_NODEVICE = -2, OBCA_E_CUDA = -3, OBCA_E_NOMEM = -4, OBCA_E_SIZE = -5 };

#define OBCA_MAX_STAGES 32   /* N + 1 <= 32: one lane per stage */
#define OBCA_MAX_OBS    12
#define OBCA_MAX_ROWS   48   /* sum of edges per time step */

typedef struct {
  int32_t mode;            /* OBCA_MODE_*                                                              */
  int32_t N;               /* horizon (N + 1 <= OBCA_MAX_STAGES)                                       */
  int32_t n_obs;           /* obstacles per time step                                                  */
  int32_t rows;            /* R = sum_i (vObs[i] - 1) half-space rows per time step                    */
  int32_t init;            /* OBCA_INIT_*                                                              */
  int32_t max_iter;        /* 3000 (mpc4, IPOPT default) / 1000 (mpc6/8: obca.py:1538)                 */
  int32_t has_term;        /* terminal set present (always for FIXED_SET)                              */
  int32_t acceptable_iter; /* 15 (IPOPT default)                                                       */
  double  Ts, dmin, ego[4];
  double  Q[9], P[9], R1[4], R2[4];
  double  xL[2], xU[2], uL[2], uU[2];
  double  acc_max[2];      /* {0.6, pi/6}  (obca.py:932-933)                                           */
  double  time_cost[2];    /* {10, 1}      (obca.py:888)                                               */
  double  T_min;           /* 1e-4         (obca.py:963)                                               */
  double  tol;             /* 1e-8  (IPOPT default)                                                    */
  double  acceptable_tol;  /* 1e-6 default; 1e-8 for mpc6/8 (obca.py:1538-1539)                        */
  double  mu_init;         /* initial barrier parameter                                                */
  double  bound_push;      /* slacks start at max(d(x0), bound_push)                                   */
} obca_params;

typedef struct obca_ctx obca_ctx;   /* opaque: device copies of params / obstacle rows + scratch */

int  obca_b200_abi_version(void);
/* device < 0 => current device.  max_batch sizes the per-instance scratch. */
int  obca_b200_create (obca_ctx** out, int device, int max_batch, const obca_params* p);
int  obca_b200_destroy(obca_ctx* ctx);
/* bytes of device scratch held by the context (for memory budgeting) */
int64_t obca_b200_scratch_bytes(const obca_ctx* ctx);

/* All pointers are DEVICE pointers except edge_ptr (host).  Asynchronous on cuda_stream.
 *   x0 [B,3]  u0 [B,2]  xref [B,N+1,3]  uref [B,N,2] or NULL
 *   T_max [B] (free modes; obca.py:961-962) or NULL    term [B,3] = {xmin, ymin, ymax} (terminal set) or NULL
 *   Ts_inst [B] per-instance sampling time or NULL (=> params.Ts).  The receding-horizon loop hands every
 *           scenario its own inherited step (closed_loop.py:586-587), so a lock-step batch needs one per instance
 *   edge_ptr [n_obs+1] prefix sum of (vObs-1), shared by the batch
 *   A [Bo,rows,2]  b0 [Bo,rows]  db [Bo,rows] or NULL (b_k = b0 + k*db; mode FREE ignores db: obca.py:969)
 *   obstacles_shared != 0 => Bo = 1 (one scene broadcast to the batch), else Bo = B
 * outputs
 *   x [B,N+1,3]  u [B,N,2]  lam [B,N+1,rows]  mu [B,N+1,4*n_obs]  T [B] (time scale; 1 in fixed modes)
 *   obj [B]  status [B]  iters [B]
 * One solve per context at a time: a context owns the work counters, the list of failed instances, the work order and
 * the watchdog checkpoints of its launches, so a second obca_b200_solve on the same context must be stream-ordered after
 * the first (same stream, or an event); concurrent solves need one context each (obca_b200_solve_host does that
 * internally with four launch slots for the chunks of a large batch).                                    */
int  obca_b200_solve  (obca_ctx* ctx, int batch,
                       const double* x0, const double* u0, const double* xref, const double* uref,
                       const double* T_max, const double* term, const double* Ts_inst,
                       const int32_t* edge_ptr, const double* A, const double* b0, const double* db,
                       int obstacles_shared,
                       double* x, double* u, double* lam, double* mu, double* T, double* obj,
                       int32_t* status, int32_t* iters, void* cuda_stream);
/* obca_b200_solve over a work list that lives on the device: work item w (0 <= w < *count_dev) solves instance
 * index_dev[w]; every per-instance array is indexed by the INSTANCE number, so callers keep one set of arrays for all
 * instances and hand each solver mode the list of those it should solve (the receding-horizon loop below: FREE /
 * FIXED_SET / FIXED_NOTERM subsets of one scenario batch, closed_loop.py:382-395) without a host round trip.
 * `batch` bounds *count_dev and the instance numbers (<= max_batch).  count_dev / index_dev NULL => obca_b200_solve. */
int  obca_b200_solve_indexed(obca_ctx* ctx, int batch, const int32_t* count_dev, const int32_t* index_dev,
                       const double* x0, const double* u0, const double* xref, const double* uref,
                       const double* T_max, const double* term, const double* Ts_inst,
                       const int32_t* edge_ptr, const double* A, const double* b0, const double* db,
                       int obstacles_shared,
                       double* x, double* u, double* lam, double* mu, double* T, double* obj,
                       int32_t* status, int32_t* iters, void* cuda_stream);
/* Same call with HOST pointers: stages inputs to the device, solves, copies the results back and
 * synchronises.  This is what the single-problem Python methods (obca.obca_mpc4 ...) use. */
int  obca_b200_solve_host(obca_ctx* ctx, int batch,
                       const double* x0, const double* u0, const double* xref, const double* uref,
                       const double* T_max, const double* term, const double* Ts_inst,
                       const int32_t* edge_ptr, const double* A, const double* b0, const double* db,
                       int obstacles_shared,
                       double* x, double* u, double* lam, double* mu, double* T, double* obj,
                       int32_t* status, int32_t* iters);
/* kernel launches issued by this context so far (bench.py's gpu_launches) */
int64_t obca_b200_launch_count(const obca_ctx* ctx);
/* diagnostics: input prefetches (cp.async.bulk) that did not complete within the kernel's bounded wait, so that plain
 * loads took over; expected 0.  Synchronises the device. */
int64_t obca_b200_bulk_timeouts(obca_ctx* ctx);
/* elapsed device time (ms) of the last solve's kernel, measured with CUDA events on its stream;
 * valid after the stream was synchronised */
float obca_b200_last_kernel_ms(obca_ctx* ctx);
const char* obca_b200_strerror(int rc);
/* Measured fp64 throughput of the device (TFLOP/s): a kernel of independent DFMA chains, best of three.  The solver is
 * bound by the fp64 pipe and its latencies, not by HBM; bench.py reports the solver's fp64 rate against this ceiling. */
int  obca_b200_fp64_peak(int device, double* tflops);

/* ---- host planner: the callers' side of the solve, batched on the host cores ----------------
 * obca_b200_astar_batch replaces a_star.solve + rebuild_path + create_reference_path (src/a_star.py:39-102, 137-147,
 * 189-200; called through closedLoop.update_path, src/closed_loop.py:555-563) for n independent queries, with the
 * reference's neighbour order, cost, heap tie-breaking and stale-entry rules, so the routes are the same cell for
 * cell.  HOST pointers.
 *   grids [n_grids,H,W] uint8, 1 = occupied     grid_index [n] or NULL (every query uses grid 0)
 *   start_rc, goal_rc [n,2] int32 (row, col)    ref [n,max_len,3] (x, y, yaw)
 *   ref_len [n]: number of path points (the start cell is excluded, as in the reference); 0 = no route or fewer
 *           than two points; -len if max_len was too small (the call then returns OBCA_E_SIZE)
 *   n_threads <= 0: all hardware threads                                                                          */
int  obca_b200_astar_batch(int n, const uint8_t* grids, int n_grids, int H, int W, const int32_t* grid_index,
                           const int32_t* start_rc, const int32_t* goal_rc, int max_len, double* ref,
                           int32_t* ref_len, int n_threads);
/* closedLoop.update_reference_trajectory (src/closed_loop.py:502-528) for n poses: window of N+1 path points from
 * the first closest one, clamped to the last.  path_index [n] or NULL (pose q uses path q).  HOST pointers.
 *   ref [n_paths,max_len,3]  ref_len [n_paths]  x0 [n,3]  xref [n,N+1,3]                                          */
int  obca_b200_reference_windows(int n, const double* ref, const int32_t* ref_len, int max_len,
                                 const int32_t* path_index, const double* x0, int N, double* xref);

/* ---- device-side input builder ----------------------------------------------------------------
 * Half-space rows from raw vertices: obstacleModel.obstacle_H_Represent (src/model_obstacle.py:37-102; same branch
 * rules, rows not normalised) for the first time block, and the per-step increment db that replaces the N+1
 * translated copies of problemSetting.rebuild_lObs (src/demo_setting.py:457-473):  b_k = b0 + k*db,
 * db = (Ts*speed) * (A . (cos, sin)).  DEVICE pointers except edge_ptr; asynchronous on the stream.
 *   edge_ptr [n_obs+1] host      verts [B, rows + n_obs, 2]: polygon i = its E_i + 1 vertices (clockwise, as the
 *   reference lists them), polygons back to back      vel [B,n_obs,3] = speed, cos(heading), sin(heading) or NULL
 *   Ts_inst [B] or NULL (=> ts)   A [B,rows,2]  b0 [B,rows]  db [B,rows] or NULL                                   */
int  obca_b200_build_rows(int batch, int n_obs, const int32_t* edge_ptr, const double* verts, const double* vel,
                          const double* Ts_inst, double ts, double* A, double* b0, double* db, void* cuda_stream);

/* ---- device-resident receding-horizon loop -----------------------------------------------------
 * closedLoop.closed_loop_mpc4 (src/closed_loop.py:323-441) for n_scenarios Monte-Carlo scenarios in lock-step: they
 * share the static map, the A* path, start and goal; each has one moving rectangle.  Per step, on the device: goal test,
 * update_obstacle (445-486), sensor (591-630), update_reference_trajectory (502-528), terminal set (371), obstacle
 * rows, then obca_mpc4 on the scenarios that see nothing, obca_mpc6 on those that do, obca_mpc8 on its failures
 * (382-395), then the first input is applied (416-419).  No host synchronisation between reset and read. */
typedef struct {
  int32_t N;               /* horizon of both phases (closed_loop.py:84,87 with N_free == N_fix)                      */
  int32_t max_steps;       /* 30 (closed_loop.py:341)                                                                 */
  int32_t n_static;        /* static obstacles; the moving one is appended after them (demo_setting.py:445-452)       */
  int32_t rows_static;     /* their half-space rows                                                                   */
  int32_t path_len;        /* points of the reference path                                                            */
  int32_t terminal_rule;   /* 0: [x0.x+5, inf) x [1, 9] (closed_loop.py:371); 1: [5, inf) x [x0.y+4, 60] (simulation.py:72) */
  double  sense;           /* lidar range (senseDis, demo_setting.py:70)                                              */
  double  goal[2];
  double  goal_tol;        /* stop when the squared distance to the goal is below this: 0.1 (closed_loop.py:343)      */
  double  start[3];
  double  Ts0;             /* sampling time before the first free-time solve: 0.1                                     */
  int32_t speculative;     /* != 0: run the solve without the terminal set beside the one with it on every detected
                              scenario and take its result where the other fails (same results; worth it where the
                              terminal-set solve mostly fails, wasteful where it mostly succeeds)                     */
  int32_t reserved;
} obca_loop_params;
typedef struct obca_loop obca_loop;
/* p_free / p_set / p_noterm: solver parameters of the three modes (n_obs = n_static, n_static+1, n_static+1).
 * HOST pointers: edges_static [n_static], A_static [rows_static,2], b_static [rows_static], path [path_len,3]. */
int  obca_b200_loop_create(obca_loop** out, int device, int n_scenarios, const obca_loop_params* lp,
                           const obca_params* p_free, const obca_params* p_set, const obca_params* p_noterm,
                           const int32_t* edges_static, const double* A_static, const double* b_static,
                           const double* path);
/* HOST: dyn [B,7] = cx, cy, heading, length, width, speed, first step; heading_cs [B,2] = cos, sin of the heading */
int  obca_b200_loop_reset(obca_loop* l, const double* dyn, const double* heading_cs, void* cuda_stream);
/* issue n_steps steps on the stream; asynchronous */
int  obca_b200_loop_run(obca_loop* l, int n_steps, void* cuda_stream);
/* copy the logs to HOST buffers (any may be NULL) and synchronise: traj [B,max_steps+1,3] (NaN after the last step
 * taken), steps [B], failed [B], mode_log [B,max_steps] (OBCA_MODE_* or -1), x [B,3], u [B,2], Ts_opt [B],
 * solves [3] = FREE / FIXED_SET / FIXED_NOTERM solves since reset */
int  obca_b200_loop_read(obca_loop* l, double* traj, int32_t* steps, int32_t* failed, int32_t* mode_log, double* x,
                         double* u, double* Ts_opt, int64_t* solves, void* cuda_stream);
int64_t obca_b200_loop_launch_count(const obca_loop* l);
int  obca_b200_loop_destroy(obca_loop* l);

#ifdef __cplusplus
}
#endif
#endif

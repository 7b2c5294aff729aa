/*
 * trajopt_b200 — C ABI of the B200-native batched DDP/iLQR solver on SO(3)/SE(3).
 *
 * The reference (chenghuailin/trajectory_optimization_matrix_lie_groups) is pure Python and has
 * no FFI; its boundary is the class API of traoptlibrary.  The entry points below are what the
 * Python controllers of the drop-in package bind through ctypes, one per reference call:
 *
 *   trajopt_create / trajopt_destroy      <- controller construction
 *                                            traoptlibrary/traopt_controller.py:532-573 (iLQR_Tracking_SO3),
 *                                            :1036-1128 (SO3_MS), :1837-1878 (SE3), :2359-2415 (SE3_MS),
 *                                            :3145-3200 (AL_iLQR_Tracking_SE3_MS)
 *   trajopt_set_params                    <- dynamics / cost / constraint objects handed to the controller
 *                                            traopt_dynamics.py:279-327, 633-690, 1214-1278;
 *                                            traopt_cost.py:297-329, 587-622, 1182-1209;
 *                                            traopt_constraints.py:70-81;  fit() keyword arguments
 *                                            traopt_controller.py:575-578, 1131-1133, 1880-1881, 2443-2445, 3218-3221
 *   trajopt_set_reference                 <- q_ref / xi_ref of the tracking cost and of the MS initial guess
 *                                            traopt_cost.py:614-616, 322-323; traopt_controller.py:2399-2400, 3123-3136
 *   trajopt_begin                         <- start of fit(): regulariser reset, initial rollout / initial guess
 *                                            traopt_controller.py:1899-1924, 2463-2491, 2015-2028, 3123-3136
 *   trajopt_iterate                       <- the `for iteration in range(n_iterations)` loop body
 *                                            traopt_controller.py:1926-2007 (SS), 2493-2633 (MS), 3231-3264 (AL outer)
 *   trajopt_iterate_inner                 <- AL: the inner fit()'s loop body, one iteration at a time (:3236-3240)
 *   trajopt_export / trajopt_export_hist  <- fit() return values and what the on_iteration callbacks record
 *                                            traopt_controller.py:2013, 2639, 3266-3267; benchmark_SE3_tracking.py:22-42
 *   trajopt_solve / trajopt_solve_host    <- one whole fit() for every problem of the batch; the batch itself
 *                                            replaces the joblib pool of visualization/perturb_all_compute.py:240-250
 *   trajopt_debug_*                       <- per-stage quantities (_linearization :2098-2176 / :2823-2910,
 *                                            _backward_pass :2178-2261 / :2912-3006) exported for parity tests
 *
 * Conventions: FP64 throughout.  Pointers named d_* are device pointers, h_* host pointers.
 * Poses cross the ABI as unit quaternion [x, y, z, w] (+ position [x, y, z] for SE3); twists are
 * [omega, v].  A state row is  SE3/drone: q(4) p(3) xi(6) = 13 doubles,  SO3: q(4) w(3) = 7 doubles.
 * Batched arrays are problem-major, C-contiguous:  x0 [B][NS], us [B][N][NU], xs [B][N+1][NS].
 * Every function returns 0 on success, a negative TRAJOPT_E_* code otherwise; trajopt_last_error()
 * returns a static description.  No exceptions cross the ABI.  A handle belongs to one device and
 * is not thread-safe.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef TRAJOPT_B200_H
#define TRAJOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TRAJOPT_RIGID: RigidBodyDynamics (traopt_dynamics.py:901-1206) = SE3 + gravity, 6 inputs
 * TRAJOPT_PEND:  Pendulum3dDyanmics (traopt_dynamics.py:421-626) = SO3 + gravity torque, pivot-force input */
enum { TRAJOPT_SO3 = 0, TRAJOPT_SE3 = 1, TRAJOPT_DRONE = 2, TRAJOPT_RIGID = 3, TRAJOPT_PEND = 4 };
enum { TRAJOPT_SS = 0, TRAJOPT_MS = 1, TRAJOPT_AL_MS = 2 };

/* per-problem status (low 4 bits) and flags */
enum {
    TRAJOPT_CONVERGED = 0,   /* gradient (and defect) tolerance met                                  */
    TRAJOPT_MAX_ITER = 1,    /* n_iterations exhausted                                               */
    TRAJOPT_NO_DESCENT = 2,  /* "Couldn't find descent direction" (traopt_controller.py:2005-2007)   */
    TRAJOPT_RUNNING = 3,     /* not finished yet (only visible between trajopt_iterate calls)        */
    TRAJOPT_FLAG_REG_EXCEEDED = 16, /* "exceeded max regularization term" fired (:2238-2240)         */
    TRAJOPT_FLAG_NONFINITE = 32     /* a cost or state became NaN/Inf                                */
};

enum {
    TRAJOPT_E_INVALID = -1,  /* bad argument                                  */
    TRAJOPT_E_CUDA = -2,     /* CUDA runtime error (see trajopt_last_error)   */
    TRAJOPT_E_STATE = -3     /* call order violated (e.g. iterate before begin) */
};

typedef struct trajopt_handle trajopt_handle;

/* Everything the reference spreads over Dynamics / Cost / Constraint / fit(**kwargs). Row-major. */
typedef struct trajopt_params {
    double dt;
    double Ib[9];            /* body inertia, J[0:3,0:3] (traopt_dynamics.py:662)                       */
    double mass;             /* J[4,4] (traopt_dynamics.py:663)                                        */
    double gravity;          /* DroneDynamics / RigidBodyDynamics _g = 9.8 (:1245, :936); ignored otherwise */
    double Q[144];           /* stage weight  [NX][NX]; only the two diagonal NPxNP blocks are read    */
    double P[144];           /* terminal weight [NX][NX]                                               */
    double R[36];            /* control weight [NU][NU]                                                */
    double lb[6], ub[6];     /* InputConstraint bounds (AL only)                                       */
    int32_t has_constraints;
    int32_t rollout_linear;  /* rollout='linear' (1) / 'nonlinear' (0)                                 */
    int32_t line_search;     /* MS merit line search (traopt_controller.py:2549-2590)                  */
    int32_t n_alphas;        /* 13 (SS, SO3-MS) or 20 (SE3-MS); 0 = class default                      */
    int32_t max_iters;       /* n_iterations                                                           */
    double tol_grad_norm;
    double tol_d_norm;
    double max_reg;          /* 1e10; <= 0 disables the give-up test like max_reg=None                 */
    double defect_kappa;     /* 1e-12 (SE3-MS) / 1e-14 (SO3-MS); 0 = class default                     */
    /* augmented Lagrangian outer loop (traopt_controller.py:3182-3184, 3218-3221) */
    int32_t n_al_iters;
    double al_mu0, al_mu_scale, al_mu_max, tol_constr;
    double length;           /* Pendulum3dDyanmics length of the stick (traopt_dynamics.py:459); ignored otherwise */
    /* AL only, beyond the reference (which has InputConstraint only, traopt_constraints.py:66-169): box bounds on the
     * body velocity xi = [omega, v], handled by the same BaseConstraint / ALConstrainedCost algebra
     * (g = [lb - xi; xi - ub], g_x = [0 -I; 0 I], traopt_cost.py:1236-1320); active at the terminal stage too */
    double xi_lb[6], xi_ub[6];
    int32_t has_state_bounds;
} trajopt_params;

const char* trajopt_last_error(void);
int trajopt_version(void);

int trajopt_create(int kind, int method, int N, int B, int device, trajopt_handle** out);
int trajopt_destroy(trajopt_handle* h);
int trajopt_set_params(trajopt_handle* h, const trajopt_params* p);
/* h_q_ref [N+1][7] (SE3/drone) or [N+1][4] (SO3), h_xi_ref [N+1][6|3]; host pointers, shared by the batch */
int trajopt_set_reference(trajopt_handle* h, const double* h_q_ref, const double* h_xi_ref);

/* Receding-horizon use (the caller loop of an MPC: the reference's README names it, its scripts solve single horizons only):
 * a shared reference LONGER than the horizon — n_rows >= N + 1 samples, same row formats as trajopt_set_reference — is
 * uploaded once, and trajopt_set_reference_offset slides the (N + 1)-sample window the problems track: stage i tracks
 * sample first_row + i.  No copy per step.  trajopt_set_reference / _batch replace it. */
int trajopt_set_reference_long(trajopt_handle* h, const double* h_q_ref, const double* h_xi_ref, int64_t n_rows);
int trajopt_set_reference_offset(trajopt_handle* h, int64_t first_row);

/* One reference per problem (the batch differs in references as well as in initial states): DEVICE pointers,
 * problem-major d_q_ref [B][N+1][7|4], d_xi_ref [B][N+1][6|3].  Replaces the shared reference until the next
 * trajopt_set_reference.  The multiple-shooting initial guess (traopt_controller.py:3123-3136) and the tracking cost
 * (traopt_cost.py:614-616) of problem b then use reference b. */
int trajopt_set_reference_batch(trajopt_handle* h, const double* d_q_ref, const double* d_xi_ref, void* stream);

/* Continuous batching: n_problems (>= 0, any number) problems through the handle's B slots.  A slot whose problem has
 * finished is given the next x0 of the queue before the following DDP iteration, so every launch works on (nearly) B
 * running problems instead of waiting for the slowest of a batch.  Problem p's result is the one trajopt_solve gives for
 * the same x0 (a problem's arithmetic does not depend on its slot); it is written to row p of the output arrays
 * (DEVICE pointers, sized for n_problems; any may be NULL): d_xs [n][N+1][NS], d_us [n][N][NU], d_J, d_iters,
 * d_status, d_grad, d_defect [n].  d_x0 [n][NS] device; d_us_init NULL (zeros) or ONE [N][NU] initial control sequence
 * shared by all problems.  Single and multiple shooting (with or without line search); not the augmented-Lagrangian
 * method, per-problem references or per-problem horizons (those belong to a batch: trajopt_solve).
 * The reference's counterpart is its joblib pool handing the next job to whichever worker is free
 * (visualization/perturb_all_compute.py:240-250). */
int trajopt_solve_stream(trajopt_handle* h, const double* d_x0, int64_t n_problems, const double* d_us_init, double* d_xs,
                         double* d_us, double* d_J, int32_t* d_iters, int32_t* d_status, double* d_grad, double* d_defect,
                         void* stream);

/* The same through HOST buffers (pinned for full speed): x0 goes up once, and the rows of the problems that are
 * complete (every id below the lowest one still running) leave for the host on a side stream while the solve goes on. */
int trajopt_solve_stream_host(trajopt_handle* h, const double* h_x0, int64_t n_problems, const double* h_us_init, double* h_xs,
                              double* h_us, double* h_J, int32_t* h_iters, int32_t* h_status, double* h_grad, double* h_defect,
                              void* stream);

/* One horizon per problem: DEVICE pointer d_N [B], clamped to [1, N]; NULL restores N for every problem.  Problem b
 * then is the N_b-stage problem on the first N_b + 1 rows of its reference (terminal cost at stage N_b); rows of the
 * exported trajectories beyond N_b hold the initial guess.  Set before trajopt_begin. */
int trajopt_set_horizons(trajopt_handle* h, const int32_t* d_N, void* stream);

/* us_mode: 0 = zeros (d_us_init ignored), 1 = one [N][NU] path shared by the batch, 2 = [B][N][NU] */
int trajopt_begin(trajopt_handle* h, const double* d_x0, const double* d_us_init, int us_mode, void* stream);
/* run up to n_iters more iterations; *n_active_out (may be NULL) = problems still running afterwards */
int trajopt_iterate(trajopt_handle* h, int n_iters, int* n_active_out, void* stream);
/* Augmented-Lagrangian handles: up to n_iters iterations of the INNER solve of the current outer iteration (the
 * `on_iteration_ilqr` granularity of AL_iLQR_Tracking_SE3_MS.fit, traopt_controller.py:3236-3240); starts the next
 * outer iteration's inner solve if the previous one has been closed.  *n_active_out = problems whose inner solve is
 * still running; once it is 0, trajopt_iterate(h, 1, ...) closes the outer iteration (constraint evaluation and
 * multiplier / penalty update, :3242-3264).  On other handles identical to trajopt_iterate. */
int trajopt_iterate_inner(trajopt_handle* h, int n_iters, int* n_active_out, void* stream);
/* any output pointer may be NULL.  d_xs [B][N+1][NS], d_us [B][N][NU], per-problem arrays [B] */
int trajopt_export(trajopt_handle* h, double* d_xs, double* d_us, double* d_J, int32_t* d_iters,
                   int32_t* d_status, double* d_grad, double* d_defect, void* stream);
/* histories, problem-major: d_J_hist [B][max_iters], d_grad_hist [B][max_iters+1],
 * d_defect_hist [B][max_iters+1], d_alpha_hist [B][max_iters] (accepted step index, -1 none) */
int trajopt_export_hist(trajopt_handle* h, double* d_J_hist, double* d_grad_hist, double* d_defect_hist,
                        int32_t* d_alpha_hist, void* stream);
/* AL with velocity bounds: d_lmbd_state [B][N+1][2NV], d_imu_state [B][N+1][2NV] */
int trajopt_export_al_state(trajopt_handle* h, double* d_lmbd_state, double* d_imu_state, void* stream);
/* regulariser state after the last backward pass: d_mu [B], d_delta [B] (traopt_controller.py:1899-1900,
 * 2233-2246; `mu` is what the on_iteration callbacks receive) */
int trajopt_export_reg(trajopt_handle* h, double* d_mu, double* d_delta, void* stream);
/* AL only: d_lmbd [B][N+1][2NU], d_imu [B][N+1][2NU] (diagonal), d_mu [B], d_outer_iters [B], d_violation [B] */
int trajopt_export_al(trajopt_handle* h, double* d_lmbd, double* d_imu, double* d_mu, int32_t* d_outer_iters,
                      double* d_violation, void* stream);

/* begin + iterate(max_iters) + export */
int trajopt_solve(trajopt_handle* h, const double* d_x0, const double* d_us_init, int us_mode,
                  double* d_xs, double* d_us, double* d_J, int32_t* d_iters, int32_t* d_status,
                  double* d_grad, double* d_defect, void* stream);
/* same with HOST buffers (pinned for full speed): copies in, solves, copies out, synchronises */
int trajopt_solve_host(trajopt_handle* h, const double* h_x0, const double* h_us_init, int us_mode,
                       double* h_xs, double* h_us, double* h_J, int32_t* h_iters, int32_t* h_status,
                       double* h_grad, double* h_defect, void* stream);

/* The same in two halves: _begin returns a ticket once the solve is complete and every device->host copy is queued (it does
 * not wait for them, and leaves `stream` free), _wait blocks until the host arrays of that ticket are filled.  The next
 * _begin on the handle may be issued before the wait: its compute overlaps the previous copies (the reference's joblib
 * pool overlaps result pickling with the next job the same way, visualization/perturb_all_compute.py:240-250).  Up to 4
 * tickets may be outstanding per handle; the host arrays of a ticket must stay untouched until its wait returns. */
int trajopt_solve_host_begin(trajopt_handle* h, const double* h_x0, const double* h_us_init, int us_mode,
                             double* h_xs, double* h_us, double* h_J, int32_t* h_iters, int32_t* h_status,
                             double* h_grad, double* h_defect, void* stream, int* ticket_out);
int trajopt_solve_host_wait(trajopt_handle* h, int ticket);

/* ---- parity-test exports (dense, problem-major) ------------------------------------------- */
/* linearise the CURRENT trajectory: d_Fx [B][N][NX][NX], d_Fu [B][N][NX][NU], d_defect [B][N][NX],
 * d_L [B][N+1], d_Lx [B][N+1][NX], d_Lxx [B][N+1][NX][NX], d_Lu [B][N][NU]; any may be NULL */
int trajopt_debug_linearize(trajopt_handle* h, double* d_Fx, double* d_Fu, double* d_defect, double* d_L,
                            double* d_Lx, double* d_Lxx, double* d_Lu, void* stream);
/* gains of the last backward pass: d_k [B][N][NU], d_K [B][N][NU][NX] */
int trajopt_debug_gains(trajopt_handle* h, double* d_k, double* d_K, void* stream);
/* line-search table of the last iteration, d_table [rows][B] (row-major, rows = trajopt_debug_linesearch_rows):
 * single shooting: row a = J_new of step size a (traopt_controller.py:1972-1990);
 * multiple shooting with line_search: rows [0, n_alphas) = J_new, [n_alphas, 2 n_alphas) = ||d_new|| of step size a
 * (:2560-2576), then c1, c2 (expected cost change :2756-2769), defect weight (:2774-2788) and merit (:2556).
 * Rows of step sizes that were not evaluated for a problem keep their previous content. */
int trajopt_debug_linesearch_rows(trajopt_handle* h);
int trajopt_debug_linesearch(trajopt_handle* h, double* d_table, void* stream);
/* the reference's per-stage callbacks on n independent rows against reference row i (0 <= i <= N):
 * d_x [n][NS], d_u [n][NU] (ignored when terminal); outputs, any may be NULL: d_f [n][NS] = f(x,u),
 * d_Fx [n][NX][NX], d_Fu [n][NX][NU], d_l [n], d_lx [n][NX], d_lxx [n][NX][NX], d_lu [n][NU],
 * d_err [n][NX] = [Log(q q_ref_i^-1); xi - xi_ref_i]   (BaseDynamics.f/f_x/f_u traopt_dynamics.py:24-64,
 * BaseCost.l/l_x/l_u/l_xx traopt_cost.py:14-110, cost._err :659-673) */
int trajopt_debug_stage(trajopt_handle* h, int i, int terminal, int n, const double* d_x, const double* d_u,
                        double* d_f, double* d_Fx, double* d_Fu, double* d_l, double* d_lx, double* d_lxx,
                        double* d_lu, double* d_err, void* stream);
/* Lie-group primitives on n independent inputs (problem-major rows), see csrc/api.cu for op codes */
int trajopt_debug_lie(int op, int n, const double* d_in, double* d_out, void* stream);

/* measured FP64 FMA throughput of the current device (the roofline denominator of this path: there is
 * no FP64 figure in MEASURED_PEAKS.json).  Runs a register-only DFMA kernel for about `ms_target`
 * milliseconds, timed with CUDA events; *out_tflops = 2 * DFMA / time. */
int trajopt_debug_fp64_peak(double ms_target, double* out_tflops, void* stream);

/* kernels launched by this library since the last reset (for bench.py's gpu_launches claim) */
int64_t trajopt_launch_count(int reset);
/* device seconds spent in the phases of the last trajopt_iterate calls since reset:
 * out[0]=linearise, out[1]=backward, out[2]=forward, out[3]=other; and their launch counts in cnt[4] */
int trajopt_phase_times(trajopt_handle* h, double* out_ms, int64_t* cnt, int reset);
int trajopt_set_profiling(trajopt_handle* h, int enable);
/* Compaction of the running problems into the leading slots during a solve (changes no result, only which CTAs
 * stay busy): active once the (padded) batch is >= min_batch and running * ratio <= slots in use.  Defaults 1024, 4;
 * min_batch < 0 disables it. */
int trajopt_set_compaction(trajopt_handle* h, int min_batch, int ratio);

/* Backward sweep of the SE3 / quadrotor / rigid-body families (_backward_pass, traopt_controller.py:2178-2261 / 2912-3006):
 * variant 0 = automatic (CTAs of six warps per 32 problems while the slots in use are at most one group per SM, counting
 * `lanes` solver handles that share the device; CTAs of four warps up to two groups per SM; CTAs of two warps above),
 * 2 / 4 / 6 = always that variant.  The variants are bit-identical; this only changes how a launch maps onto the SMs.
 * Defaults 0, 1. */
int trajopt_set_sweep(trajopt_handle* h, int variant, int lanes);

/* Line searches (single shooting, traopt_controller.py:1972-1990; multiple shooting with line_search, :2549-2590) try the
 * step sizes alphas[0], alphas[1], ... in order and accept the first that passes.  Large batches roll out alphas[0] for
 * everybody, then the rest for the problems that rejected it, then the accepted one again to keep its trajectory.  While
 * the batch (padded to 32) is <= max_batch one launch rolls out EVERY step size and keeps every candidate trajectory
 * (n_alphas x the trajectory memory): a solve of a few problems costs its number of dependent rollouts.  Same decisions,
 * bit-identical results.  Default 256; 0 = never. */
int trajopt_set_line_search_batch(trajopt_handle* h, int max_batch);

#ifdef __cplusplus
}
#endif
#endif /* TRAJOPT_B200_H */

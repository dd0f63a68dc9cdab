/* nmpc_b200.h -- C-ABI of the B200-native batched NMPC solver (libnmpc_b200.so).
 *
 * Drop-in boundary: this library sits behind the one call the reference
 * makes on its hot path,
 *     solver = nlpsol('solver','ipopt', nlp_prob, opts)       AllScripts/centralized_six_robots_implementation.py:345-346
 *     sol    = solver(x0=, p=, lbx=, ubx=, lbg=, ubg=)        AllScripts/centralized_six_robots_implementation.py:432
 * for the NLP family those scripts build (multiple-shooting unicycle robots with pairwise
 * collision rows, :207-352; single robot: casadi_test.py:34-109), plus the two warm-start
 * helpers of the MPC loop (shift(), :160-169, and the X0 shift at :465) and the Euler plant of
 * casadi_test.py:17-26.
 *
 * Conventions
 *  - plain C, no torch types; every array is IEEE double unless stated, row-major [B, .].
 *  - decision vector  w = [X_0; ...; X_N; U_0; ...; U_{N-1}]  (n = 3Nr(N+1) + 2Nr N), CasADi's
 *    column-major reshape of the reference's X (ns x N+1) and U (nc x N)          (:240-245,339)
 *  - parameter  p = [x0bar (3Nr); xs (3Nr)]                                        (:419)
 *  - constraint rows g: N+1 blocks of [3Nr equality rows ; M = Nr(Nr-1)/2 distance rows]; block 0 =
 *    X_0 - x0bar and M constant rows (3.5); block k+1 = Euler defect of stage k and the squared
 *    distances on X_k, pairs in lexicographic order                               (:278,282-331)
 *  - multipliers in CasADi's sign convention (L = f + lam_g'g + lam_x'x).
 *  - functions whose names end in _host take HOST pointers (pageable or pinned: they are handed to
 *    cudaMemcpyAsync as they are, so only pinned buffers overlap with device work), copy them into a
 *    device staging area owned by the handle, and synchronise before returning; all others take DEVICE
 *    pointers, are asynchronous on `stream` and never allocate.
 *  - a handle belongs to the CUDA device that was current in nmpc_create; every later call must be made
 *    with the same device current (checked: NMPC_EINVAL otherwise).
 *  - return value: 0 on success, <0 on an API error (nmpc_last_error() has the text).
 *    Non-convergence is NOT an error: it is reported per instance in status[], and the last
 *    iterate is returned, as CasADi does by default (the reference never reads the status).
 */
#ifndef NMPC_B200_H
#define NMPC_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nmpc_desc {
    int Nr;        /* robots          (reference variable m,  ...six...py:199)  1..10 on the CUDA path */
    int N;         /* horizon         (reference variable N,  :198)                                 */
    double T;      /* sampling period (:197)                                                        */
    double Q[3];   /* diag state weights   (1, 5, 0.1)   :252-259                                   */
    double R[2];   /* diag control weights (0.5, 0.05)   :261-266                                   */
} nmpc_desc;

/* Solver options: the ones the reference sets (:345) and the IPOPT defaults that shape the path. */
typedef struct nmpc_opts {
    double tol;                       /* 1e-8                                   */
    int    max_iter;                  /* 2000  (:345)                           */
    double acceptable_tol;            /* 1e-8  (:345)                           */
    int    acceptable_iter;           /* 15                                     */
    double acceptable_obj_change_tol; /* 1e-6  (:345)                           */
    double dual_inf_tol;              /* 1                                      */
    double constr_viol_tol;           /* 1e-4                                   */
    double compl_inf_tol;             /* 1e-4                                   */
    double mu_init;                   /* 0.1                                    */
    double kappa_mu;                  /* 0.2   mu_linear_decrease_factor        */
    double theta_mu;                  /* 1.5   mu_superlinear_decrease_power    */
    double barrier_tol_factor;        /* 10                                     */
    double tau_min;                   /* 0.99                                   */
    double bound_push;                /* 0.01  (and slack_bound_push)           */
    double bound_frac;                /* 0.01  (and slack_bound_frac)           */
    double bound_relax_factor;        /* 1e-8                                   */
    double bound_mult_init_val;       /* 1                                      */
    double constr_mult_init_max;      /* 1e3                                    */
    double kappa_sigma;               /* 1e10                                   */
    double kappa_d;                   /* 1e-5                                   */
    double nlp_scaling_max_gradient;  /* 100                                    */
    int    max_soc;                   /* 4                                      */
    int    max_resto_iter;            /* 100  bound on the restoration fallback */
} nmpc_opts;

enum { NMPC_SOLVED = 0, NMPC_ACCEPTABLE = 1, NMPC_MAX_ITER = 2, NMPC_INFEASIBLE = 3, NMPC_NUMERICAL = 4 };

/* stats[B][NMPC_NSTATS] */
enum { NMPC_ST_KKT_ERR = 0, NMPC_ST_PRIMAL_INF, NMPC_ST_DUAL_INF, NMPC_ST_COMPL, NMPC_ST_MU,
       NMPC_ST_N_REG, NMPC_ST_N_RESTO, NMPC_ST_N_SOC, NMPC_ST_N_FACTOR, NMPC_ST_N_LS,
       NMPC_ST_FILTER_EVICT,   /* filter entries dropped because the fixed-size filter was full (IPOPT's is unbounded): 0 on every parity case */
       NMPC_NSTATS };

enum { NMPC_EINVAL = -1, NMPC_EBOUNDS = -2, NMPC_ENOTSUP = -3, NMPC_ECUDA = -4, NMPC_ENOMEM = -5 };

typedef struct nmpc_handle nmpc_handle;

void nmpc_default_opts(nmpc_opts *o);
const char *nmpc_last_error(void);

/* Replaces the nlpsol(...) factory (:345-346).  The handle owns only immutable problem metadata. */
int nmpc_create(const nmpc_desc *d, const nmpc_opts *o, nmpc_handle **out);
void nmpc_destroy(nmpc_handle *h);

/* Launch tuning (no effect on results).  These used to be environment variables read inside nmpc_create; they are explicit
 * now so that nothing on the product path depends on the process environment.
 *   convoy            0 off, 1 the warps of a CTA meet at the start of every interior-point iteration, 2 (default) also
 *                     before the forward pass (instruction-cache locality of the one-warp-per-instance path)
 *   ctas_per_sm       > 0: cap on resident CTAs per SM (occupancy experiments); 0 = as many as fit
 *   force_block_path  != 0: run any Nr on the CTA-per-instance dense-block solver (test hook)
 *   thread_min_batch  > 0: batches at least this large of the one-robot static-obstacle family use the thread-per-instance
 *                     solver; 0 = never */
typedef struct nmpc_tuning { int convoy, ctas_per_sm, force_block_path, thread_min_batch; } nmpc_tuning;
void nmpc_default_tuning(nmpc_tuning *t);
/* nmpc_create / nmpc_create_obstacles (n_obs > 0) with explicit tuning; t == NULL means the defaults. */
int nmpc_create_tuned(const nmpc_desc *d, const nmpc_opts *o, const nmpc_tuning *t, int n_obs, const double *obs, nmpc_handle **out);

int nmpc_n(const nmpc_handle *h);        /* decision variables                         */
int nmpc_mg(const nmpc_handle *h);       /* constraint rows                            */
int nmpc_np(const nmpc_handle *h);       /* parameters (6 Nr)                          */
int nmpc_nnz_jac(const nmpc_handle *h);  /* 3Nr + N(11Nr+4M)   CasADi's CCS of dg/dw   */
int nmpc_nnz_hess(const nmpc_handle *h); /* N(6Nr+2M)          lower triangle          */

/* Bytes of device scratch nmpc_solve needs for a batch of B (independent of B beyond the
 * number of resident warps, so one workspace serves any batch size <= the B asked for). */
size_t nmpc_workspace_bytes(const nmpc_handle *h, int B);
/* Same when lbx/ubx/lbg/ubg are per-instance (bounds_batched = 1): adds B sets of stage-layout bound rows. */
size_t nmpc_workspace_bytes_batched_bounds(const nmpc_handle *h, int B);

/* Replaces  sol = solver(x0=,p=,lbx=,ubx=,lbg=,ubg=)  (:432) for B independent instances.
 * bounds_batched = 0: lbx/ubx [n], lbg/ubg [mg] shared by the batch; 1: [B,n] / [B,mg].
 * Outputs x [B,n], f [B], g [B,mg], lam_x [B,n], lam_g [B,mg], status/iters [B] (int32),
 * stats [B,NMPC_NSTATS]; any output except x may be NULL. */
int nmpc_solve(nmpc_handle *h, int B, const double *x0, const double *p,
               const double *lbx, const double *ubx, const double *lbg, const double *ubg,
               int bounds_batched, double *x, double *f, double *g, double *lam_x, double *lam_g,
               int32_t *status, int32_t *iters, double *stats,
               void *workspace, size_t workspace_bytes, void *stream);

/* Static circular obstacles (the reference's obstacle-avoidance scripts, first_/third_scenario_mpc_obstacle_avoidance.py:96-152):
 * same variables, cost and dynamics, plus n_obs rows per robot and stage
 *     sqrt((x_i - ox)^2 + (y_i - oy)^2) - clearance          on X_k, k = 0..N-1
 * with clearance = rob_dim + r_obs (:125), bounded below by the margin the scripts put in lbg (0.05 / 0.1) and above by +inf.
 * obs: HOST array [n_obs][3] = (ox, oy, clearance).  Row layout of g / lbg / ubg / lam_g (as the scripts build it, :109-125):
 * [X_0 - x0bar (3 Nr)], then for k = 0..N-1: [defect rows (3 Nr); pair rows (M, none for one robot); obstacle rows, robot-major
 * (Nr n_obs)]; mg = 3 Nr + N (3 Nr + M + Nr n_obs).  Runs on the warp-per-instance path for 1..4 robots while M + Nr n_obs <= 32
 * (one robot with up to 32 obstacles; the reference's scripts have one robot and 1, 4 or 6), else on the CTA-per-instance
 * dense-block path; nmpc_eval and the CCS patterns are not available for this family. */
int nmpc_create_obstacles(const nmpc_desc *d, const nmpc_opts *o, int n_obs, const double *obs, nmpc_handle **out);

/* Small generic optimal-control problems, one GPU thread per instance.  model NMPC_OCP_VAN_DER_POL is the direct-multiple-shooting
 * demo of mpc_pose_control_casadi.py:22-114: 2 states, 1 control, N intervals over the horizon T, each integrated by rk_steps RK4
 * steps together with the cost quadrature; decision vector INTERLEAVED [X_0, U_0, X_1, ..., U_{N-1}, X_N] (n = 3N + 2, :77-106),
 * g = F(X_k, U_k) - X_{k+1} (mg = 2N, :104), no parameter vector (pass p = NULL).  The initial state must be fixed through
 * lbx == ubx on X_0 (:79-80; IPOPT's make_parameter treatment); other variables must have lb < ub.  o == NULL: IPOPT defaults
 * (max_iter 3000).  nmpc_solve / nmpc_solve_host / nmpc_n / nmpc_mg work on the handle; shift, plant and eval do not apply. */
enum { NMPC_OCP_VAN_DER_POL = 1 };
int nmpc_create_ocp(int model, int N, double T, int rk_steps, const nmpc_opts *o, nmpc_handle **out);

/* Scheduling hint for the following nmpc_solve* calls on this handle: order [len] (DEVICE int32, caller owned, a permutation
 * of 0..len-1) is the sequence in which the persistent teams pull instances from the work queue; NULL (len ignored) restores
 * index order.  Instances differ in iteration count (17 on average, up to 70, one MPC step after a solve), so a closed loop
 * that passes the previous step's iteration counts sorted in descending order (longest first) shortens the tail of the launch.
 * Results do not depend on the order.  The array must stay valid, and its contents must be complete on the stream the solve
 * is launched on (nmpc_solve_host uses the handle's own stream: synchronise the producer first), until the hint is replaced.
 * A solve whose batch size differs from len fails with NMPC_EINVAL; entries outside 0..len-1 are skipped by the kernel
 * (the outputs of the instances a broken permutation leaves out are not written), so a stale or corrupt order can never
 * cause an out-of-bounds access. */
int nmpc_set_order(nmpc_handle *h, const int32_t *order, int len);

/* nmpc_solve plus a per-iteration trace [B][max_trace][8] = (mu, scaled KKT error, theta, f, alpha_primal,
 * alpha_dual, delta_w, line-search trials) -- the parity tests compare it with the oracle's trace. */
int nmpc_solve_trace(nmpc_handle *h, int B, const double *x0, const double *p,
                     const double *lbx, const double *ubx, const double *lbg, const double *ubg,
                     int bounds_batched, double *x, double *f, double *g, double *lam_x, double *lam_g,
                     int32_t *status, int32_t *iters, double *stats, double *trace, int max_trace,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Same call with HOST buffers (what a CasADi-style caller has): H2D of the inputs, solve, D2H of
 * the requested outputs, synchronous.  Staging buffers live in the handle and grow on demand. */
int nmpc_solve_host(nmpc_handle *h, int B, const double *x0, const double *p,
                    const double *lbx, const double *ubx, const double *lbg, const double *ubg,
                    int bounds_batched, double *x, double *f, double *g, double *lam_x, double *lam_g,
                    int32_t *status, int32_t *iters, double *stats);

/* Warm-start shift between MPC steps: u0 = [u[1:]; u[-1]] (shift(), :160-169) and
 * X0 = [X[1:]; X[N-1]] (:465 -- the reference appends row N-1, not row N).  [B,n] -> [B,n]. */
int nmpc_shift(nmpc_handle *h, int B, const double *x_prev, double *x0_next, void *stream);

/* Euler plant of casadi_test.py:17-26: state <- state + T f(state, u), u = first control of x_opt
 * (x_opt [B,n]); state, state_next [B,3Nr] (may alias). */
int nmpc_plant(nmpc_handle *h, int B, const double *state, const double *x_opt, double *state_next,
               void *stream);

/* Stand-alone derivative evaluation (what CasADi's AD hands IPOPT): for B points
 * w [B,n], p [B,6Nr], lam_g [B,mg]:  f [B], grad [B,n], g [B,mg], jac [B,nnz_jac] and
 * hess [B,nnz_hess] (CCS value order of nmpc_jac_pattern / nmpc_hess_pattern).  Outputs may be NULL. */
int nmpc_eval(nmpc_handle *h, int B, const double *w, const double *p, const double *lam_g,
              double *f, double *grad, double *g, double *jac, double *hess, void *stream);

/* CCS patterns (HOST int32 arrays: colptr [n+1], rowidx [nnz]). */
int nmpc_jac_pattern(const nmpc_handle *h, int32_t *colptr, int32_t *rowidx);
int nmpc_hess_pattern(const nmpc_handle *h, int32_t *colptr, int32_t *rowidx);

/* Measurement helper: sustained FP64 FMA throughput of the device (TFLOP/s) from a register-only DFMA
 * kernel -- the roofline denominator for the factorisation (MEASURED_PEAKS.json has no FP64 figure). */
int nmpc_probe_fp64(double *tflops_out);
/* Same for the FP64 tensor-core instruction (mma.sync.m8n8k4.f64, DMMA; tcgen05 has no f64 kind): the number behind the
 * dense-block path's choice of register tiles on the FP64 FMA pipe (csrc/block_solver.cuh). */
int nmpc_probe_dmma(double *tflops_out);

/* Debug aid: cycle counters of the phases of the dense-block factorisation (Nr > 10), CTA 0 only:
 * [0] stage-parallel pre-pass, [1] p + P r mat-vec, [2] stage-matrix assembly, [3] control-block elimination,
 * [4] first contraction, [5] second contraction.  reset != 0 clears the counters after the read. */
int nmpc_debug_block_profile(long long *out16, int reset);

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
long long nmpc_launch_count(const nmpc_handle *h);

#ifdef __cplusplus
}
#endif
#endif

/* CPU oracle for the multi-robot unicycle NMPC hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libnmpc_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference (the .py scripts under /root/reference/AllScripts) delegates the whole
 * solve to CasADi -> IPOPT -> MUMPS, none of which is vendored, pinned or installable
 * here, and it ships no tests or golden vectors.  This oracle restates
 *   (1) the NLP exactly as the scripts build it
 *       (centralized_six_robots_implementation.py:207-352, casadi_test.py:34-109), and
 *   (2) the published IPOPT algorithm (Waechter & Biegler, Math. Prog. 106, 2006) with
 *       IPOPT's documented option defaults and the options the scripts set
 *       (centralized_six_robots_implementation.py:345).
 * It is checked against finite differences, a dense NumPy KKT solve, SciPy SLSQP /
 * trust-constr KKT points and the solver-independent known answers recorded in
 * SURVEY.md App. D (tests/); it is NOT checked against real IPOPT output.
 */
#ifndef NMPC_ORACLE_H
#define NMPC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int Nr;       /* robots                      (reference variable m)            */
    int N;        /* horizon                     (reference variable N)            */
    double T;     /* sampling period                                              */
    double Q[3];  /* diag state weights  (1,5,0.1)  ...six...py:252-259            */
    double R[2];  /* diag control weights (0.5,0.05) ...six...py:261-266           */
    /* static circular obstacles (first_/third_scenario_mpc_obstacle_avoidance.py:96-152); 0 / NULL = the centralized family.
     * With nobs > 0 every robot gets nobs rows per stage  sqrt((x-ox)^2+(y-oy)^2) - clearance  after the pair rows, block 0 of
     * g holds the initial condition only (no dummy rows), mg = 3Nr + N(3Nr + M + Nr nobs); orc_eval and the CCS patterns do
     * not cover this family (orc_solve / orc_solve_batch do). */
    int nobs;
    const double *obs;  /* [nobs][3] = centre x, y, clearance (rob_dim + r_obs) */
} orc_desc;

typedef struct {
    double tol;                 /* 1e-8   IPOPT default                            */
    int    max_iter;            /* 2000   ...six...py:345                          */
    double acceptable_tol;      /* 1e-8   ...six...py:345                          */
    int    acceptable_iter;     /* 15                                              */
    double acceptable_obj_change_tol; /* 1e-6 ...six...py:345                      */
    double dual_inf_tol;        /* 1                                               */
    double constr_viol_tol;     /* 1e-4                                            */
    double compl_inf_tol;       /* 1e-4                                            */
    double mu_init;             /* 0.1                                             */
    double kappa_mu;            /* 0.2   mu_linear_decrease_factor                 */
    double theta_mu;            /* 1.5   mu_superlinear_decrease_power             */
    double barrier_tol_factor;  /* 10                                              */
    double tau_min;             /* 0.99                                            */
    double bound_push;          /* 0.01  (also slack_bound_push)                   */
    double bound_frac;          /* 0.01  (also slack_bound_frac)                   */
    double bound_relax_factor;  /* 1e-8                                            */
    double bound_mult_init_val; /* 1                                               */
    double constr_mult_init_max;/* 1e3                                             */
    double kappa_sigma;         /* 1e10                                            */
    double kappa_d;             /* 1e-5                                            */
    double nlp_scaling_max_gradient; /* 100                                        */
    int    max_soc;             /* 4                                               */
    int    max_resto_iter;      /* bound on the feasibility-restoration fallback   */
} orc_opts;

enum { ORC_SOLVED = 0, ORC_ACCEPTABLE = 1, ORC_MAX_ITER = 2, ORC_INFEASIBLE = 3, ORC_NUMERICAL = 4 };

/* stats[] slots written by orc_solve (length ORC_NSTATS) */
enum { ORC_ST_KKT_ERR = 0, ORC_ST_PRIMAL_INF, ORC_ST_DUAL_INF, ORC_ST_COMPL, ORC_ST_MU,
       ORC_ST_N_REG, ORC_ST_N_RESTO, ORC_ST_N_SOC, ORC_ST_N_FACTOR, ORC_ST_N_LS,
       ORC_ST_FILTER_EVICT /* filter overflows (0) */, ORC_ST_FILTER_ADDS /* most entries added within one barrier subproblem */, ORC_NSTATS };

/* per-iteration trace row (length ORC_NTRACE) */
enum { ORC_TR_MU = 0, ORC_TR_ERR, ORC_TR_THETA, ORC_TR_OBJ, ORC_TR_ALPHA_PR, ORC_TR_ALPHA_DU,
       ORC_TR_DELTA_W, ORC_TR_N_LS, ORC_NTRACE };

void orc_default_opts(orc_opts *o);

int orc_n(const orc_desc *d);        /* decision variables      ns(N+1)+nc N        */
int orc_mg(const orc_desc *d);       /* constraint rows         (N+1)(ns+M)         */
int orc_nnz_jac(const orc_desc *d);  /* 3Nr + N(11Nr+4M)                            */
int orc_nnz_hess(const orc_desc *d); /* N(6Nr+2M)  (lower triangle)                 */
void orc_jac_pattern(const orc_desc *d, int *colptr, int *rowidx);
void orc_hess_pattern(const orc_desc *d, int *colptr, int *rowidx);

/* f, grad f, g, CCS values of dg/dw and of the lower triangle of
 * hess(f + lam_g' g).  Any output may be NULL; lam_g may be NULL iff hess_vals is. */
void orc_eval(const orc_desc *d, const double *w, const double *p, const double *lam_g,
              double *f, double *grad, double *g, double *jac_vals, double *hess_vals);

/* One nlpsol-style solve.  Outputs other than x may be NULL.  trace (max_trace x ORC_NTRACE)
 * may be NULL.  Returns 0, or <0 on API error. */
int orc_solve(const orc_desc *d, const orc_opts *o, const double *x0, const double *p,
              const double *lbx, const double *ubx, const double *lbg, const double *ubg,
              double *x, double *f, double *g, double *lam_x, double *lam_g,
              int *status, int *iters, double *stats, double *trace, int max_trace);

/* B independent solves, OpenMP over instances.  bounds_batched: 0 = bounds shared, 1 = [B,.] */
int orc_solve_batch(const orc_desc *d, const orc_opts *o, int B, const double *x0, const double *p,
                    const double *lbx, const double *ubx, const double *lbg, const double *ubg,
                    int bounds_batched, double *x, double *f, double *g, double *lam_x,
                    double *lam_g, int *status, int *iters, double *stats, int nthreads);

/* Warm-start shift (...six...py:160-169,465) and Euler plant (casadi_test.py:17-26). */
void orc_shift(const orc_desc *d, const double *x_prev, double *x0_next);
void orc_plant(const orc_desc *d, const double *state, const double *u0, double *state_next);

/* One Newton/KKT step of the interior-point system at a given primal-dual point, for
 * step-level parity tests (all vectors in the flat reference layout: w[n], g-rows[mg]):
 *   (W + Sx + dw I) dx + J' ylam = -gx ;  J dx - ds = -rg  with ds free on equality rows
 *   forced to 0, and  (Ss + dw) ds - ylam = -gs  on inequality rows.
 * in : w, lam_g (for W), sig_x[n], sig_s[mg] (ignored on equality rows), gx[n], gs[mg], rg[mg]
 * out: dx[n], ds[mg], ylam[mg];  returns 0 ok, 1 = wrong inertia (a pivot <= 0). */
int orc_kkt_step(const orc_desc *d, const double *p, const double *lbg, const double *ubg,
                 const double *w, const double *lam_g, double obj_scale,
                 const double *sig_x, const double *sig_s, double delta_w,
                 const double *gx, const double *gs, const double *rg,
                 double *dx, double *ds, double *ylam);

#ifdef __cplusplus
}
#endif
#endif

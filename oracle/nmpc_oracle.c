/* CPU oracle for the multi-robot unicycle NMPC hot path -- TEST INFRASTRUCTURE ONLY.
 * See nmpc_oracle.h for scope and the PARITY UNPINNED statement.
 *
 * Part 1 restates the NLP of /root/reference/AllScripts/centralized_six_robots_implementation.py
 *   :207-245 (symbols, layout)  :252-266 (Q,R)  :278 (initial block + dummy rows)
 *   :282-331 (cost, Euler defects, pairwise squared distances)  :339 (decision vector)
 *   :349-352 (bounds)  :160-169,465 (warm-start shift)   casadi_test.py:17-26 (Euler plant)
 * Part 2 restates what the third-party nlpsol('ipopt') call at :345-346,432 does, from the
 *   published algorithm (Waechter & Biegler 2006) and IPOPT's documented defaults; the KKT
 *   system is solved stage-wise (Riccati) instead of with MUMPS.
 */
#include "nmpc_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DUMMY_ROW_VALUE 3.5 /* ...six...py:278 */
#define FILTER_CAP 256   /* IPOPT's filter is unbounded; an overflow is counted in ORC_ST_FILTER_EVICT (never seen) */

/* ------------------------------------------------------------------------------------ */
/* dimensions                                                                            */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int Nr, N, ns, nc, nz, M, n, mg, blk, nX, nI, nE;
    int Mp, nobs;      /* M = Mp pair rows + Nr * nobs obstacle rows per block */
    const double *obs;
    double T, Q[3], R[2];
    int *pi, *pj; /* lexicographic pairs i<j  (...six...py:288-306) */
} dims_t;

/* offset of block b in the flat g / lbg / ubg / lam_g vectors: the obstacle family has no inequality rows in block 0 */
static inline int goff(const dims_t *D, int b) { return D->nobs ? (b == 0 ? 0 : D->ns + (b - 1) * D->blk) : b * D->blk; }

/* inequality row q of a block on the stage vector zk: value, gradient w.r.t. (x_i, y_i) (negated for robot j of a pair;
 * j = -1 for an obstacle row, first_scenario_mpc_obstacle_avoidance.py:125) and its second derivatives */
typedef struct { double dv, gx, gy, hxx, hyy, hxy; int i, j; } rowg_t;
static inline rowg_t row_geom(const dims_t *D, const double *zk, int q)
{
    rowg_t r;
    if (q < D->Mp) {
        r.i = D->pi[q]; r.j = D->pj[q];
        double dx = zk[3 * r.i] - zk[3 * r.j], dy = zk[3 * r.i + 1] - zk[3 * r.j + 1];
        r.dv = dx * dx + dy * dy; r.gx = 2 * dx; r.gy = 2 * dy; r.hxx = 2; r.hyy = 2; r.hxy = 0;
    } else {
        int e = q - D->Mp, o = e % D->nobs;
        r.i = e / D->nobs; r.j = -1;
        double dx = zk[3 * r.i] - D->obs[3 * o], dy = zk[3 * r.i + 1] - D->obs[3 * o + 1];
        double rho = sqrt(dx * dx + dy * dy);
        r.gx = dx / rho; r.gy = dy / rho; r.dv = rho - D->obs[3 * o + 2];
        r.hxx = (1 - r.gx * r.gx) / rho; r.hyy = (1 - r.gy * r.gy) / rho; r.hxy = -r.gx * r.gy / rho;
    }
    return r;
}

static void dims_init(dims_t *D, const orc_desc *d)
{
    D->Nr = d->Nr; D->N = d->N; D->T = d->T;
    memcpy(D->Q, d->Q, sizeof D->Q); memcpy(D->R, d->R, sizeof D->R);
    D->ns = 3 * d->Nr; D->nc = 2 * d->Nr; D->nz = 5 * d->Nr;
    D->Mp = d->Nr * (d->Nr - 1) / 2; D->nobs = d->obs ? d->nobs : 0; D->obs = d->obs;
    D->M = D->Mp + d->Nr * D->nobs;
    D->nX = D->ns * (D->N + 1);
    D->n = D->nX + D->nc * D->N;
    D->blk = D->ns + D->M;
    D->mg = D->nobs ? D->ns + D->N * D->blk : (D->N + 1) * D->blk;
    D->nI = (D->N + 1) * D->M;
    D->nE = (D->N + 1) * D->ns;
    D->pi = (int *)malloc(sizeof(int) * (D->Mp + 1));
    D->pj = (int *)malloc(sizeof(int) * (D->Mp + 1));
    int q = 0;
    for (int i = 0; i < D->Nr; i++)
        for (int j = i + 1; j < D->Nr; j++) { D->pi[q] = i; D->pj[q] = j; q++; }
}
static void dims_free(dims_t *D) { free(D->pi); free(D->pj); }

int orc_n(const orc_desc *d) { return 3 * d->Nr * (d->N + 1) + 2 * d->Nr * d->N; }
int orc_mg(const orc_desc *d)
{
    int M = d->Nr * (d->Nr - 1) / 2;
    if (d->obs && d->nobs > 0) return 3 * d->Nr + d->N * (3 * d->Nr + M + d->Nr * d->nobs);
    return (d->N + 1) * (3 * d->Nr + M);
}
int orc_nnz_jac(const orc_desc *d)
{ int M = d->Nr * (d->Nr - 1) / 2; return 3 * d->Nr + d->N * (11 * d->Nr + 4 * M); }
int orc_nnz_hess(const orc_desc *d)
{ int M = d->Nr * (d->Nr - 1) / 2; return d->N * (6 * d->Nr + 2 * M); }

/* flat index of state c of robot i at stage k, control c of robot i at stage k */
static inline int IX(const dims_t *D, int k, int i, int c) { return k * D->ns + 3 * i + c; }
static inline int IU(const dims_t *D, int k, int i, int c) { return D->nX + k * D->nc + 2 * i + c; }

/* ------------------------------------------------------------------------------------ */
/* Part 1a: flat-layout evaluation with CCS derivative values (mirrors what CasADi's AD    */
/*          hands to IPOPT: grad f, jac g, hess of  sigma f + lam' g)                      */
/* ------------------------------------------------------------------------------------ */
typedef struct { int r, c; double v; } trip_t;

static int trip_cmp(const void *a, const void *b)
{
    const trip_t *x = (const trip_t *)a, *y = (const trip_t *)b;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    return 0;
}

/* Jacobian triplets in a fixed emission order (values optional: w may be NULL -> v = 1) */
static int jac_trips(const dims_t *D, const double *w, trip_t *t)
{
    int nn = 0; double T = D->T;
    for (int r = 0; r < D->ns; r++) { t[nn].r = r; t[nn].c = r; t[nn].v = 1.0; nn++; }
    for (int k = 0; k < D->N; k++) {
        int b = (k + 1) * D->blk;
        for (int i = 0; i < D->Nr; i++) {
            double th = w ? w[IX(D, k, i, 2)] : 0.3, v = w ? w[IU(D, k, i, 0)] : 1.0;
            double c = cos(th), s = sin(th);
            int rx = b + 3 * i, ry = rx + 1, rt = rx + 2;
            t[nn++] = (trip_t){rx, IX(D, k + 1, i, 0), 1.0};
            t[nn++] = (trip_t){rx, IX(D, k, i, 0), -1.0};
            t[nn++] = (trip_t){rx, IX(D, k, i, 2), T * v * s};
            t[nn++] = (trip_t){rx, IU(D, k, i, 0), -T * c};
            t[nn++] = (trip_t){ry, IX(D, k + 1, i, 1), 1.0};
            t[nn++] = (trip_t){ry, IX(D, k, i, 1), -1.0};
            t[nn++] = (trip_t){ry, IX(D, k, i, 2), -T * v * c};
            t[nn++] = (trip_t){ry, IU(D, k, i, 0), -T * s};
            t[nn++] = (trip_t){rt, IX(D, k + 1, i, 2), 1.0};
            t[nn++] = (trip_t){rt, IX(D, k, i, 2), -1.0};
            t[nn++] = (trip_t){rt, IU(D, k, i, 1), -T};
        }
        for (int q = 0; q < D->M; q++) {
            int i = D->pi[q], j = D->pj[q], r = b + D->ns + q;
            double dx = w ? w[IX(D, k, i, 0)] - w[IX(D, k, j, 0)] : 1.0;
            double dy = w ? w[IX(D, k, i, 1)] - w[IX(D, k, j, 1)] : 1.0;
            t[nn++] = (trip_t){r, IX(D, k, i, 0), 2 * dx};
            t[nn++] = (trip_t){r, IX(D, k, j, 0), -2 * dx};
            t[nn++] = (trip_t){r, IX(D, k, i, 1), 2 * dy};
            t[nn++] = (trip_t){r, IX(D, k, j, 1), -2 * dy};
        }
    }
    return nn;
}

/* lower-triangle Hessian triplets (row >= col) */
static int hess_trips(const dims_t *D, const double *w, const double *lam, trip_t *t)
{
    int nn = 0; double T = D->T;
    for (int k = 0; k < D->N; k++) {
        int b = (k + 1) * D->blk;
        for (int i = 0; i < D->Nr; i++) {
            double th = w ? w[IX(D, k, i, 2)] : 0.3, v = w ? w[IU(D, k, i, 0)] : 1.0;
            double lx = lam ? lam[b + 3 * i] : 1.0, ly = lam ? lam[b + 3 * i + 1] : 1.0;
            double c = cos(th), s = sin(th);
            double sx = 0, sy = 0; /* collision curvature on the diagonal */
            for (int q = 0; q < D->M; q++)
                if (D->pi[q] == i || D->pj[q] == i) {
                    double mu = lam ? lam[b + D->ns + q] : 1.0; sx += 2 * mu; sy += 2 * mu;
                }
            t[nn++] = (trip_t){IX(D, k, i, 0), IX(D, k, i, 0), 2 * D->Q[0] + sx};
            t[nn++] = (trip_t){IX(D, k, i, 1), IX(D, k, i, 1), 2 * D->Q[1] + sy};
            t[nn++] = (trip_t){IX(D, k, i, 2), IX(D, k, i, 2), 2 * D->Q[2] + T * v * (lx * c + ly * s)};
            t[nn++] = (trip_t){IU(D, k, i, 0), IX(D, k, i, 2), T * (lx * s - ly * c)};
            t[nn++] = (trip_t){IU(D, k, i, 0), IU(D, k, i, 0), 2 * D->R[0]};
            t[nn++] = (trip_t){IU(D, k, i, 1), IU(D, k, i, 1), 2 * D->R[1]};
        }
        for (int q = 0; q < D->M; q++) {
            int i = D->pi[q], j = D->pj[q];
            double mu = lam ? lam[b + D->ns + q] : 1.0;
            t[nn++] = (trip_t){IX(D, k, j, 0), IX(D, k, i, 0), -2 * mu};
            t[nn++] = (trip_t){IX(D, k, j, 1), IX(D, k, i, 1), -2 * mu};
        }
    }
    return nn;
}

static void trips_to_ccs(trip_t *t, int nn, int ncol, int *colptr, int *rowidx)
{
    qsort(t, nn, sizeof(trip_t), trip_cmp);
    int c = 0; colptr[0] = 0;
    for (int e = 0; e < nn; e++) {
        while (c < t[e].c) colptr[++c] = e;
        rowidx[e] = t[e].r;
    }
    while (c < ncol) colptr[++c] = nn;
}

void orc_jac_pattern(const orc_desc *d, int *colptr, int *rowidx)
{
    dims_t D; dims_init(&D, d);
    trip_t *t = (trip_t *)malloc(sizeof(trip_t) * orc_nnz_jac(d));
    int nn = jac_trips(&D, NULL, t);
    trips_to_ccs(t, nn, D.n, colptr, rowidx);
    free(t); dims_free(&D);
}

void orc_hess_pattern(const orc_desc *d, int *colptr, int *rowidx)
{
    dims_t D; dims_init(&D, d);
    trip_t *t = (trip_t *)malloc(sizeof(trip_t) * (orc_nnz_hess(d) + 1));
    int nn = hess_trips(&D, NULL, NULL, t);
    trips_to_ccs(t, nn, D.n, colptr, rowidx);
    free(t); dims_free(&D);
}

void orc_eval(const orc_desc *d, const double *w, const double *p, const double *lam_g,
              double *f, double *grad, double *g, double *jac_vals, double *hess_vals)
{
    dims_t D; dims_init(&D, d);
    const double *xs = p + D.ns; double T = D.T;
    if (f || grad) {
        double acc = 0;
        if (grad) memset(grad, 0, sizeof(double) * D.n);
        for (int k = 0; k < D.N; k++)
            for (int i = 0; i < D.Nr; i++) {
                for (int c = 0; c < 3; c++) {
                    double e = w[IX(&D, k, i, c)] - xs[3 * i + c];
                    acc += D.Q[c] * e * e;
                    if (grad) grad[IX(&D, k, i, c)] = 2 * D.Q[c] * e;
                }
                for (int c = 0; c < 2; c++) {
                    double u = w[IU(&D, k, i, c)];
                    acc += D.R[c] * u * u;
                    if (grad) grad[IU(&D, k, i, c)] = 2 * D.R[c] * u;
                }
            }
        if (f) *f = acc;
    }
    if (g) {
        for (int r = 0; r < D.ns; r++) g[r] = w[r] - p[r];
        for (int q = 0; q < D.M; q++) g[D.ns + q] = DUMMY_ROW_VALUE;
        for (int k = 0; k < D.N; k++) {
            int b = (k + 1) * D.blk;
            for (int i = 0; i < D.Nr; i++) {
                double th = w[IX(&D, k, i, 2)], v = w[IU(&D, k, i, 0)], om = w[IU(&D, k, i, 1)];
                g[b + 3 * i] = w[IX(&D, k + 1, i, 0)] - (w[IX(&D, k, i, 0)] + T * v * cos(th));
                g[b + 3 * i + 1] = w[IX(&D, k + 1, i, 1)] - (w[IX(&D, k, i, 1)] + T * v * sin(th));
                g[b + 3 * i + 2] = w[IX(&D, k + 1, i, 2)] - (th + T * om);
            }
            for (int q = 0; q < D.M; q++) {
                int i = D.pi[q], j = D.pj[q];
                double dx = w[IX(&D, k, i, 0)] - w[IX(&D, k, j, 0)];
                double dy = w[IX(&D, k, i, 1)] - w[IX(&D, k, j, 1)];
                g[b + D.ns + q] = dx * dx + dy * dy;
            }
        }
    }
    if (jac_vals) {
        int nn = orc_nnz_jac(d);
        trip_t *t = (trip_t *)malloc(sizeof(trip_t) * nn);
        jac_trips(&D, w, t);
        qsort(t, nn, sizeof(trip_t), trip_cmp);
        for (int e = 0; e < nn; e++) jac_vals[e] = t[e].v;
        free(t);
    }
    if (hess_vals) {
        int nn = orc_nnz_hess(d);
        trip_t *t = (trip_t *)malloc(sizeof(trip_t) * (nn + 1));
        hess_trips(&D, w, lam_g, t);
        qsort(t, nn, sizeof(trip_t), trip_cmp);
        for (int e = 0; e < nn; e++) hess_vals[e] = t[e].v;
        free(t);
    }
    dims_free(&D);
}

/* ...six...py:160-169 (u0 = [u[1:]; u[-1]]) and :465 (X0 = [X[1:]; X[N-1]], row N-1 not N) */
void orc_shift(const orc_desc *d, const double *xp, double *xn)
{
    dims_t D; dims_init(&D, d);
    double *tmp = (double *)malloc(sizeof(double) * D.n);
    for (int k = 0; k < D.N; k++) memcpy(tmp + k * D.ns, xp + (k + 1) * D.ns, sizeof(double) * D.ns);
    memcpy(tmp + D.N * D.ns, xp + (D.N - 1) * D.ns, sizeof(double) * D.ns);
    for (int k = 0; k < D.N - 1; k++)
        memcpy(tmp + D.nX + k * D.nc, xp + D.nX + (k + 1) * D.nc, sizeof(double) * D.nc);
    memcpy(tmp + D.nX + (D.N - 1) * D.nc, xp + D.nX + (D.N - 1) * D.nc, sizeof(double) * D.nc);
    memcpy(xn, tmp, sizeof(double) * D.n);
    free(tmp); dims_free(&D);
}

/* casadi_test.py:17-26  x <- x + T f(x,u) */
void orc_plant(const orc_desc *d, const double *st, const double *u0, double *out)
{
    for (int i = 0; i < d->Nr; i++) {
        double th = st[3 * i + 2], v = u0[2 * i], om = u0[2 * i + 1];
        out[3 * i] = st[3 * i] + d->T * v * cos(th);
        out[3 * i + 1] = st[3 * i + 1] + d->T * v * sin(th);
        out[3 * i + 2] = th + d->T * om;
    }
}

/* ------------------------------------------------------------------------------------ */
/* Part 1b: stage-structured evaluation used by the solver                                */
/*   stage vector z_k = [X_k (ns) ; U_k (nc)], k = 0..N (U_N unused, kept 0)              */
/*   equality blocks  e = 0..N : block 0 = X_0 - xbar0, block k+1 = defect of stage k      */
/*   inequality blocks b = 0..N: block 0 = dummy rows (3.5), block k+1 = distances on X_k  */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    dims_t D;
    const orc_opts *o;
    const double *p;
    double df;                       /* objective scaling (nlp_scaling_max_gradient)      */
    double *zl, *zu;                 /* relaxed variable bounds, stage layout, +-inf      */
    double *ceq;                     /* rhs of equality rows                              */
    double *dl, *du;                 /* relaxed inequality bounds                         */
    /* Riccati storage */
    double *Pall, *pall, *Kall, *kall, *Mw, *W1, *mw;
} ctx_t;

static void trig_stage(const ctx_t *C, const double *zk, double *cs, double *sn)
{
    for (int i = 0; i < C->D.Nr; i++) { cs[i] = cos(zk[3 * i + 2]); sn[i] = sin(zk[3 * i + 2]); }
}

/* c[(N+1)*ns] = equality residuals, dv[(N+1)*M] = inequality row values */
static void eval_cons(const ctx_t *C, const double *z, double *c, double *dv)
{
    const dims_t *D = &C->D; int ns = D->ns, nz = D->nz, M = D->M; double T = D->T;
    for (int j = 0; j < ns; j++) c[j] = z[j] - C->p[j] - C->ceq[j];
    for (int q = 0; q < M; q++) dv[q] = DUMMY_ROW_VALUE;
    for (int k = 0; k < D->N; k++) {
        const double *zk = z + k * nz, *zn = z + (k + 1) * nz;
        double *ck = c + (k + 1) * ns; const double *ce = C->ceq + (k + 1) * ns;
        for (int i = 0; i < D->Nr; i++) {
            double th = zk[3 * i + 2], v = zk[ns + 2 * i], om = zk[ns + 2 * i + 1];
            ck[3 * i] = zn[3 * i] - (zk[3 * i] + T * v * cos(th)) - ce[3 * i];
            ck[3 * i + 1] = zn[3 * i + 1] - (zk[3 * i + 1] + T * v * sin(th)) - ce[3 * i + 1];
            ck[3 * i + 2] = zn[3 * i + 2] - (th + T * om) - ce[3 * i + 2];
        }
        double *dk = dv + (k + 1) * M;
        /* (pair rows keep their original expression: exactly symmetric scenarios decide their branch by round-off) */
        for (int q = 0; q < D->Mp; q++) {
            int i = D->pi[q], j = D->pj[q];
            double dx = zk[3 * i] - zk[3 * j], dy = zk[3 * i + 1] - zk[3 * j + 1];
            dk[q] = dx * dx + dy * dy;
        }
        for (int q = D->Mp; q < M; q++) dk[q] = row_geom(D, zk, q).dv;
    }
}

static double eval_obj(const ctx_t *C, const double *z)
{
    const dims_t *D = &C->D; const double *xs = C->p + D->ns; double acc = 0;
    for (int k = 0; k < D->N; k++) {
        const double *zk = z + k * D->nz;
        for (int i = 0; i < D->Nr; i++) {
            for (int c = 0; c < 3; c++) { double e = zk[3 * i + c] - xs[3 * i + c]; acc += D->Q[c] * e * e; }
            for (int c = 0; c < 2; c++) { double u = zk[D->ns + 2 * i + c]; acc += D->R[c] * u * u; }
        }
    }
    return acc;
}

static void eval_grad(const ctx_t *C, const double *z, double *gf)
{
    const dims_t *D = &C->D; const double *xs = C->p + D->ns;
    memset(gf, 0, sizeof(double) * (D->N + 1) * D->nz);
    for (int k = 0; k < D->N; k++) {
        const double *zk = z + k * D->nz; double *gk = gf + k * D->nz;
        for (int i = 0; i < D->Nr; i++) {
            for (int c = 0; c < 3; c++) gk[3 * i + c] = 2 * D->Q[c] * (zk[3 * i + c] - xs[3 * i + c]);
            for (int c = 0; c < 2; c++) gk[D->ns + 2 * i + c] = 2 * D->R[c] * zk[D->ns + 2 * i + c];
        }
    }
}

/* out = Jc' yc + Jd' yd  (stage layout) */
static void eval_jtv(const ctx_t *C, const double *z, const double *yc, const double *yd, double *out)
{
    const dims_t *D = &C->D; int ns = D->ns, nz = D->nz, M = D->M; double T = D->T;
    memset(out, 0, sizeof(double) * (D->N + 1) * nz);
    for (int k = 0; k <= D->N; k++)
        for (int j = 0; j < ns; j++) out[k * nz + j] = yc[k * ns + j];
    for (int k = 0; k < D->N; k++) {
        const double *zk = z + k * nz; double *ok = out + k * nz;
        const double *lam = yc + (k + 1) * ns, *mu = yd + (k + 1) * M;
        for (int i = 0; i < D->Nr; i++) {
            double th = zk[3 * i + 2], v = zk[ns + 2 * i], c = cos(th), s = sin(th);
            double lx = lam[3 * i], ly = lam[3 * i + 1], lt = lam[3 * i + 2];
            ok[3 * i] -= lx; ok[3 * i + 1] -= ly;
            ok[3 * i + 2] -= lt + (-T * v * s) * lx + (T * v * c) * ly;
            ok[ns + 2 * i] -= T * (c * lx + s * ly);
            ok[ns + 2 * i + 1] -= T * lt;
        }
        for (int q = 0; q < D->Mp; q++) {
            int i = D->pi[q], j = D->pj[q];
            double dx = zk[3 * i] - zk[3 * j], dy = zk[3 * i + 1] - zk[3 * j + 1];
            ok[3 * i] += 2 * dx * mu[q]; ok[3 * j] -= 2 * dx * mu[q];
            ok[3 * i + 1] += 2 * dy * mu[q]; ok[3 * j + 1] -= 2 * dy * mu[q];
        }
        for (int q = D->Mp; q < M; q++) {
            rowg_t g = row_geom(D, zk, q);
            ok[3 * g.i] += g.gx * mu[q]; ok[3 * g.i + 1] += g.gy * mu[q];
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Part 2a: stage-wise (Riccati) solve of the condensed primal-dual system                 */
/*   (W + Sx + dw) dz + Jc' ytc + Jd' ytd = -gx                                            */
/*   (Ss + dw) ds - ytd = -gs ;  Jc dz = -rc ;  Jd dz - ds = -rd                            */
/* W = dfW * hess f + sum ycW hess c + sum ydW hess d ; act[r]=0 rows (no bound) are skipped */
/* Returns 0, or 1 when a control-block pivot is <= 0 (wrong inertia).                     */
/* ------------------------------------------------------------------------------------ */
static int kkt_solve(ctx_t *C, const double *z, const double *ycW, const double *ydW, double dfW,
                     double zeta, const double *sigx, const double *sigs, const unsigned char *act,
                     double delta, const double *gx, const double *gs, const double *rc,
                     const double *rd, double *dz, double *ds, double *ytc, double *ytd)
{
    const dims_t *D = &C->D;
    const int ns = D->ns, nc = D->nc, nz = D->nz, M = D->M, N = D->N, Nr = D->Nr;
    const double T = D->T;
    double *P = C->Pall + (size_t)N * ns * ns, *pv = C->pall + (size_t)N * ns;
    double cs[64], sn[64], a[64], b[64], pr[192];
    /* terminal stage: no cost, no distance rows (X_N is unconstrained but boxed) */
    memset(P, 0, sizeof(double) * ns * ns);
    for (int j = 0; j < ns; j++) { P[j * ns + j] = sigx[N * nz + j] + delta + zeta; pv[j] = gx[N * nz + j]; }
    for (int k = N - 1; k >= 0; k--) {
        const double *zk = z + k * nz;
        const double *Pn = C->Pall + (size_t)(k + 1) * ns * ns, *pn = C->pall + (size_t)(k + 1) * ns;
        double *Mw = C->Mw, *W1 = C->W1, *m = C->mw;
        trig_stage(C, zk, cs, sn);
        for (int i = 0; i < Nr; i++) { double v = zk[ns + 2 * i]; a[i] = -T * v * sn[i]; b[i] = T * v * cs[i]; }
        /* pr = pn + Pn r,  r = -rc[k+1] */
        for (int r = 0; r < ns; r++) {
            double acc = pn[r];
            for (int j = 0; j < ns; j++) acc -= Pn[r * ns + j] * rc[(k + 1) * ns + j];
            pr[r] = acc;
        }
        /* W1 = Pn [A B]  (ns x nz) */
        for (int r = 0; r < ns; r++)
            for (int i = 0; i < Nr; i++) {
                double Px = Pn[r * ns + 3 * i], Py = Pn[r * ns + 3 * i + 1], Pt = Pn[r * ns + 3 * i + 2];
                W1[r * nz + 3 * i] = Px; W1[r * nz + 3 * i + 1] = Py;
                W1[r * nz + 3 * i + 2] = Pt + a[i] * Px + b[i] * Py;
                W1[r * nz + ns + 2 * i] = T * (cs[i] * Px + sn[i] * Py);
                W1[r * nz + ns + 2 * i + 1] = T * Pt;
            }
        /* M = [A B]' W1 ; m = [A B]' pr */
        for (int i = 0; i < Nr; i++) {
            for (int col = 0; col < nz; col++) {
                double Wx = W1[(3 * i) * nz + col], Wy = W1[(3 * i + 1) * nz + col], Wt = W1[(3 * i + 2) * nz + col];
                Mw[(3 * i) * nz + col] = Wx; Mw[(3 * i + 1) * nz + col] = Wy;
                Mw[(3 * i + 2) * nz + col] = Wt + a[i] * Wx + b[i] * Wy;
                Mw[(ns + 2 * i) * nz + col] = T * (cs[i] * Wx + sn[i] * Wy);
                Mw[(ns + 2 * i + 1) * nz + col] = T * Wt;
            }
            double px = pr[3 * i], py = pr[3 * i + 1], pt = pr[3 * i + 2];
            m[3 * i] = px; m[3 * i + 1] = py; m[3 * i + 2] = pt + a[i] * px + b[i] * py;
            m[ns + 2 * i] = T * (cs[i] * px + sn[i] * py); m[ns + 2 * i + 1] = T * pt;
        }
        /* + stage Hessian and gradient */
        for (int j = 0; j < nz; j++) { Mw[j * nz + j] += sigx[k * nz + j] + delta + zeta; m[j] += gx[k * nz + j]; }
        const double *lam = ycW + (k + 1) * ns;
        for (int i = 0; i < Nr; i++) {
            double v = zk[ns + 2 * i], lx = lam[3 * i], ly = lam[3 * i + 1];
            for (int c = 0; c < 3; c++) Mw[(3 * i + c) * nz + 3 * i + c] += dfW * 2 * D->Q[c];
            for (int c = 0; c < 2; c++) Mw[(ns + 2 * i + c) * nz + ns + 2 * i + c] += dfW * 2 * D->R[c];
            Mw[(3 * i + 2) * nz + 3 * i + 2] += T * v * (lx * cs[i] + ly * sn[i]);
            double cr = T * (lx * sn[i] - ly * cs[i]);
            Mw[(3 * i + 2) * nz + ns + 2 * i] += cr; Mw[(ns + 2 * i) * nz + 3 * i + 2] += cr;
        }
        for (int q = 0; q < M; q++) {
            int r = (k + 1) * M + q;
            if (!act[r]) continue;
            if (q >= D->Mp) { /* obstacle row: only robot i's own block, curvature (I - n n')/rho */
                rowg_t rg = row_geom(D, zk, q);
                double Dq = sigs[r] + delta, hq = Dq * rd[r] + gs[r];
                double oxx = Dq * rg.gx * rg.gx + ydW[r] * rg.hxx, oyy = Dq * rg.gy * rg.gy + ydW[r] * rg.hyy, oxy = Dq * rg.gx * rg.gy + ydW[r] * rg.hxy;
                int xi = 3 * rg.i, yi = 3 * rg.i + 1;
                Mw[xi * nz + xi] += oxx; Mw[yi * nz + yi] += oyy; Mw[xi * nz + yi] += oxy; Mw[yi * nz + xi] += oxy;
                m[xi] += rg.gx * hq; m[yi] += rg.gy * hq;
                continue;
            }
            int i = D->pi[q], j = D->pj[q];
            double gxq = 2 * (zk[3 * i] - zk[3 * j]), gyq = 2 * (zk[3 * i + 1] - zk[3 * j + 1]);
            double Dq = sigs[r] + delta, mu2 = 2 * ydW[r], hq = Dq * rd[r] + gs[r];
            double xx = Dq * gxq * gxq + mu2, yy = Dq * gyq * gyq + mu2, xy = Dq * gxq * gyq;
            int xi = 3 * i, yi = 3 * i + 1, xj = 3 * j, yj = 3 * j + 1;
            Mw[xi * nz + xi] += xx; Mw[xj * nz + xj] += xx; Mw[xi * nz + xj] -= xx; Mw[xj * nz + xi] -= xx;
            Mw[yi * nz + yi] += yy; Mw[yj * nz + yj] += yy; Mw[yi * nz + yj] -= yy; Mw[yj * nz + yi] -= yy;
            Mw[xi * nz + yi] += xy; Mw[yi * nz + xi] += xy; Mw[xj * nz + yj] += xy; Mw[yj * nz + xj] += xy;
            Mw[xi * nz + yj] -= xy; Mw[yj * nz + xi] -= xy; Mw[xj * nz + yi] -= xy; Mw[yi * nz + xj] -= xy;
            m[xi] += gxq * hq; m[xj] -= gxq * hq; m[yi] += gyq * hq; m[yj] -= gyq * hq;
        }
        /* symmetric sweep of the control pivots */
        for (int j = ns; j < nz; j++) {
            double d = Mw[j * nz + j];
            if (!(d > 0.0) || !isfinite(d)) return 1;
            double inv = 1.0 / d;
            double colj[320];
            for (int i = 0; i < nz; i++) colj[i] = Mw[i * nz + j];
            double mj = m[j] * inv;
            for (int l = 0; l < nz; l++) {
                if (l == j) continue;
                double t = Mw[j * nz + l] * inv;
                for (int i = 0; i < nz; i++) if (i != j) Mw[i * nz + l] -= colj[i] * t;
                Mw[j * nz + l] = t;
            }
            for (int i = 0; i < nz; i++) if (i != j) { m[i] -= colj[i] * mj; Mw[i * nz + j] = colj[i] * inv; }
            m[j] = mj; Mw[j * nz + j] = -inv;
        }
        double *Pk = C->Pall + (size_t)k * ns * ns, *pk = C->pall + (size_t)k * ns;
        double *Kk = C->Kall + (size_t)k * nc * ns, *kk = C->kall + (size_t)k * nc;
        for (int r = 0; r < ns; r++) {
            for (int j = 0; j < ns; j++) Pk[r * ns + j] = 0.5 * (Mw[r * nz + j] + Mw[j * nz + r]);
            pk[r] = m[r];
        }
        for (int u = 0; u < nc; u++) {
            for (int j = 0; j < ns; j++) Kk[u * ns + j] = Mw[(ns + u) * nz + j];
            kk[u] = m[ns + u];
        }
    }
    /* forward pass */
    for (int j = 0; j < ns; j++) dz[j] = -rc[j];
    for (int k = 0; k <= N; k++) {
        double *dk = dz + k * nz;
        const double *Pk = C->Pall + (size_t)k * ns * ns, *pk = C->pall + (size_t)k * ns;
        for (int r = 0; r < ns; r++) {
            double acc = pk[r];
            for (int j = 0; j < ns; j++) acc += Pk[r * ns + j] * dk[j];
            ytc[k * ns + r] = -acc;
        }
        if (k == N) { for (int u = 0; u < nc; u++) dk[ns + u] = 0; break; }
        const double *zk = z + k * nz;
        const double *Kk = C->Kall + (size_t)k * nc * ns, *kk = C->kall + (size_t)k * nc;
        for (int u = 0; u < nc; u++) {
            double acc = kk[u];
            for (int j = 0; j < ns; j++) acc += Kk[u * ns + j] * dk[j];
            dk[ns + u] = -acc;
        }
        double *dn = dz + (k + 1) * nz;
        for (int i = 0; i < Nr; i++) {
            double th = zk[3 * i + 2], v = zk[ns + 2 * i], c = cos(th), s = sin(th);
            double dv = dk[ns + 2 * i], dw = dk[ns + 2 * i + 1], dth = dk[3 * i + 2];
            dn[3 * i] = dk[3 * i] + (-T * v * s) * dth + T * c * dv - rc[(k + 1) * ns + 3 * i];
            dn[3 * i + 1] = dk[3 * i + 1] + (T * v * c) * dth + T * s * dv - rc[(k + 1) * ns + 3 * i + 1];
            dn[3 * i + 2] = dth + T * dw - rc[(k + 1) * ns + 3 * i + 2];
        }
    }
    /* slack steps and inequality multipliers */
    for (int q = 0; q < M; q++) {
        ds[q] = act[q] ? rd[q] : 0.0;
        ytd[q] = act[q] ? (sigs[q] + delta) * ds[q] + gs[q] : 0.0;
    }
    for (int k = 0; k < N; k++) {
        const double *zk = z + k * nz, *dk = dz + k * nz;
        for (int q = 0; q < M; q++) {
            int r = (k + 1) * M + q;
            if (!act[r]) { ds[r] = 0; ytd[r] = 0; continue; }
            if (q >= D->Mp) {
                rowg_t rg = row_geom(D, zk, q);
                ds[r] = rg.gx * dk[3 * rg.i] + rg.gy * dk[3 * rg.i + 1] + rd[r];
            } else {
                int i = D->pi[q], j = D->pj[q];
                double gxq = 2 * (zk[3 * i] - zk[3 * j]), gyq = 2 * (zk[3 * i + 1] - zk[3 * j + 1]);
                ds[r] = gxq * (dk[3 * i] - dk[3 * j]) + gyq * (dk[3 * i + 1] - dk[3 * j + 1]) + rd[r];
            }
            ytd[r] = (sigs[r] + delta) * ds[r] + gs[r];
        }
    }
    return 0;
}

static void ctx_alloc_riccati(ctx_t *C)
{
    const dims_t *D = &C->D; int ns = D->ns, nc = D->nc, nz = D->nz, N = D->N;
    C->Pall = (double *)malloc(sizeof(double) * (size_t)(N + 1) * ns * ns);
    C->pall = (double *)malloc(sizeof(double) * (size_t)(N + 1) * ns);
    C->Kall = (double *)malloc(sizeof(double) * (size_t)(N ? N : 1) * nc * ns);
    C->kall = (double *)malloc(sizeof(double) * (size_t)(N ? N : 1) * nc);
    C->Mw = (double *)malloc(sizeof(double) * (size_t)nz * nz);
    C->W1 = (double *)malloc(sizeof(double) * (size_t)ns * nz);
    C->mw = (double *)malloc(sizeof(double) * (size_t)nz);
}
static void ctx_free_riccati(ctx_t *C)
{ free(C->Pall); free(C->pall); free(C->Kall); free(C->kall); free(C->Mw); free(C->W1); free(C->mw); }

/* flat w[n] <-> stage layout z[(N+1)*nz] */
static void flat_to_stage(const dims_t *D, const double *w, double *z, double fill)
{
    for (int k = 0; k <= D->N; k++) {
        for (int j = 0; j < D->ns; j++) z[k * D->nz + j] = w[k * D->ns + j];
        for (int u = 0; u < D->nc; u++) z[k * D->nz + D->ns + u] = k < D->N ? w[D->nX + k * D->nc + u] : fill;
    }
}
static void stage_to_flat(const dims_t *D, const double *z, double *w)
{
    for (int k = 0; k <= D->N; k++) {
        for (int j = 0; j < D->ns; j++) w[k * D->ns + j] = z[k * D->nz + j];
        if (k < D->N) for (int u = 0; u < D->nc; u++) w[D->nX + k * D->nc + u] = z[k * D->nz + D->ns + u];
    }
}

int orc_kkt_step(const orc_desc *d, const double *p, const double *lbg, const double *ubg,
                 const double *w, const double *lam_g, double obj_scale,
                 const double *sig_x, const double *sig_s, double delta_w,
                 const double *gx, const double *gs, const double *rg,
                 double *dx, double *ds, double *ylam)
{
    ctx_t C; memset(&C, 0, sizeof C); dims_init(&C.D, d); C.p = p;
    const dims_t *D = &C.D; int ns = D->ns, nz = D->nz, M = D->M, N = D->N;
    ctx_alloc_riccati(&C);
    size_t nZ = (size_t)(N + 1) * nz;
    double *z = (double *)calloc(nZ * 4 + (size_t)D->nE * 3 + (size_t)D->nI * 6, sizeof(double));
    double *sx = z + nZ, *gxs = sx + nZ, *dzs = gxs + nZ;
    double *yc = dzs + nZ, *rc = yc + D->nE, *ytc = rc + D->nE;
    double *yd = ytc + D->nE, *ss = yd + D->nI, *gss = ss + D->nI, *rd = gss + D->nI, *dss = rd + D->nI, *ytd = dss + D->nI;
    unsigned char *act = (unsigned char *)malloc(D->nI);
    flat_to_stage(D, w, z, 0.0); flat_to_stage(D, sig_x, sx, 1.0); flat_to_stage(D, gx, gxs, 0.0);
    for (int b = 0; b <= N; b++) {
        for (int j = 0; j < ns; j++) { yc[b * ns + j] = lam_g[b * D->blk + j]; rc[b * ns + j] = rg[b * D->blk + j]; }
        for (int q = 0; q < M; q++) {
            int r = b * D->blk + ns + q, e = b * M + q;
            yd[e] = lam_g[r]; ss[e] = sig_s[r]; gss[e] = gs[r]; rd[e] = rg[r];
            act[e] = (lbg[r] > -INFINITY || ubg[r] < INFINITY) ? 1 : 0;
        }
    }
    int rc_ = kkt_solve(&C, z, yc, yd, obj_scale, 0.0, sx, ss, act, delta_w, gxs, gss, rc, rd, dzs, dss, ytc, ytd);
    if (!rc_) {
        stage_to_flat(D, dzs, dx);
        for (int b = 0; b <= N; b++) {
            for (int j = 0; j < ns; j++) { ylam[b * D->blk + j] = ytc[b * ns + j]; ds[b * D->blk + j] = 0; }
            for (int q = 0; q < M; q++) { ylam[b * D->blk + ns + q] = ytd[b * M + q]; ds[b * D->blk + ns + q] = dss[b * M + q]; }
        }
    }
    free(act); free(z); ctx_free_riccati(&C); dims_free(&C.D);
    return rc_;
}

/* ------------------------------------------------------------------------------------ */
/* Part 2b: primal-dual interior-point filter line-search method                          */
/* ------------------------------------------------------------------------------------ */
void orc_default_opts(orc_opts *o)
{
    o->tol = 1e-8; o->max_iter = 2000; o->acceptable_tol = 1e-8; o->acceptable_iter = 15;
    o->acceptable_obj_change_tol = 1e-6;
    o->dual_inf_tol = 1.0; o->constr_viol_tol = 1e-4; o->compl_inf_tol = 1e-4;
    o->mu_init = 0.1; o->kappa_mu = 0.2; o->theta_mu = 1.5; o->barrier_tol_factor = 10.0;
    o->tau_min = 0.99; o->bound_push = 0.01; o->bound_frac = 0.01; o->bound_relax_factor = 1e-8;
    o->bound_mult_init_val = 1.0; o->constr_mult_init_max = 1e3; o->kappa_sigma = 1e10;
    o->kappa_d = 1e-5; o->nlp_scaling_max_gradient = 100.0; o->max_soc = 4; o->max_resto_iter = 100;
}

static inline int cmp_le(double lhs, double rhs, double bas)
{ return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(bas); }

typedef struct {
    /* iterate */
    double *z, *zL, *zU, *yc, *s, *vL, *vU, *yd;
    /* step */
    double *dz, *dzL, *dzU, *ytc, *ds, *dvL, *dvU, *ytd;
    /* work */
    double *c, *dv, *dms, *gf, *jty, *sigx, *sigs, *gx, *gs, *zt, *st, *ct, *dvt, *csoc, *dsoc,
        *dz2, *ds2, *ytc2, *ytd2, *zero_e, *zero_i, *one_z, *one_i;
    unsigned char *act;
} vecs_t;

static double push_in(double x, double l, double u, double k1, double k2)
{
    int hl = l > -INFINITY, hu = u < INFINITY;
    if (hl && hu) {
        double pl = fmin(k1 * fmax(1.0, fabs(l)), k2 * (u - l));
        double pu = fmin(k1 * fmax(1.0, fabs(u)), k2 * (u - l));
        x = fmax(x, l + pl); x = fmin(x, u - pu);
    } else if (hl) x = fmax(x, l + k1 * fmax(1.0, fabs(l)));
    else if (hu) x = fmin(x, u - k1 * fmax(1.0, fabs(u)));
    return x;
}

/* theta = ||c||_1 + ||d - s||_1 */
static double theta_of(const ctx_t *C, const vecs_t *V, const double *c, const double *dv, const double *s)
{
    double t = 0;
    for (int r = 0; r < C->D.nE; r++) t += fabs(c[r]);
    for (int r = 0; r < C->D.nI; r++) if (V->act[r]) t += fabs(dv[r] - s[r]);
    return t;
}

/* barrier objective  df f - mu sum ln(slack) + kappa_d mu sum(one-sided slack) */
static double barrier_of(const ctx_t *C, const vecs_t *V, const double *z, const double *s, double mu)
{
    const dims_t *D = &C->D; double kd = C->o->kappa_d;
    double lg = 0, damp = 0;
    for (int k = 0; k <= D->N; k++)
        for (int j = 0; j < (k < D->N ? D->nz : D->ns); j++) {
            int e = k * D->nz + j; int hl = C->zl[e] > -INFINITY, hu = C->zu[e] < INFINITY;
            if (hl) lg += log(z[e] - C->zl[e]);
            if (hu) lg += log(C->zu[e] - z[e]);
            if (hl && !hu) damp += z[e] - C->zl[e];
            if (hu && !hl) damp += C->zu[e] - z[e];
        }
    for (int r = 0; r < D->nI; r++) {
        if (!V->act[r]) continue;
        int hl = C->dl[r] > -INFINITY, hu = C->du[r] < INFINITY;
        if (hl) lg += log(s[r] - C->dl[r]);
        if (hu) lg += log(C->du[r] - s[r]);
        if (hl && !hu) damp += s[r] - C->dl[r];
        if (hu && !hl) damp += C->du[r] - s[r];
    }
    return C->df * eval_obj(C, z) - mu * lg + kd * mu * damp;
}

typedef struct { double th[FILTER_CAP], ph[FILTER_CAP]; int n, evict, added, added_max; } filter_t;

static int filter_ok(const filter_t *F, double th, double ph)
{
    for (int i = 0; i < F->n; i++)
        if (!(th < F->th[i] || ph < F->ph[i])) return 0;
    return 1;
}
static void filter_add(filter_t *F, double th, double ph)
{
    int m = 0;
    for (int i = 0; i < F->n; i++)
        if (!(F->th[i] >= th && F->ph[i] >= ph)) { F->th[m] = F->th[i]; F->ph[m] = F->ph[i]; m++; }
    F->n = m;
    if (F->n == FILTER_CAP) { /* drop the oldest */
        for (int i = 1; i < F->n; i++) { F->th[i - 1] = F->th[i]; F->ph[i - 1] = F->ph[i]; }
        F->n--; F->evict++;
    }
    F->th[F->n] = th; F->ph[F->n] = ph; F->n++;
    F->added++;      /* entries added since the last reset: what an append-only filter (the CUDA kernels') has to hold */
    if (F->added > F->added_max) F->added_max = F->added;
}

/* largest alpha in (0,1] keeping primal slacks >= (1-tau) of their current value */
static double alpha_primal_max(const ctx_t *C, const vecs_t *V, const double *z, const double *s,
                               const double *dz, const double *ds, double tau)
{
    const dims_t *D = &C->D; double a = 1.0;
    for (int k = 0; k <= D->N; k++)
        for (int j = 0; j < (k < D->N ? D->nz : D->ns); j++) {
            int e = k * D->nz + j;
            if (C->zl[e] > -INFINITY && dz[e] < 0) a = fmin(a, -tau * (z[e] - C->zl[e]) / dz[e]);
            if (C->zu[e] < INFINITY && dz[e] > 0) a = fmin(a, tau * (C->zu[e] - z[e]) / dz[e]);
        }
    for (int r = 0; r < D->nI; r++) {
        if (!V->act[r]) continue;
        if (C->dl[r] > -INFINITY && ds[r] < 0) a = fmin(a, -tau * (s[r] - C->dl[r]) / ds[r]);
        if (C->du[r] < INFINITY && ds[r] > 0) a = fmin(a, tau * (C->du[r] - s[r]) / ds[r]);
    }
    return a;
}

static int solve_one(const orc_desc *d, const orc_opts *o, const double *x0, const double *p,
                     const double *lbx, const double *ubx, const double *lbg, const double *ubg,
                     double *x, double *f, double *g, double *lam_x, double *lam_g,
                     int *status, int *iters, double *stats, double *trace, int max_trace)
{
    ctx_t C; memset(&C, 0, sizeof C); dims_init(&C.D, d); C.o = o; C.p = p;
    const dims_t *D = &C.D;
    const int ns = D->ns, nz = D->nz, M = D->M, N = D->N, nE = D->nE, nI = D->nI;
    const size_t nZ = (size_t)(N + 1) * nz;
    int rc = 0;
    /* ---- allocate ---- */
    size_t tot = nZ * 19 + (size_t)nE * 10 + (size_t)nI * 26;
    double *pool = (double *)calloc(tot, sizeof(double)), *q = pool;
    vecs_t V;
#define TAKE(name, cnt) name = q; q += (cnt)
    TAKE(C.zl, nZ); TAKE(C.zu, nZ); TAKE(V.z, nZ); TAKE(V.zL, nZ); TAKE(V.zU, nZ); TAKE(V.dz, nZ);
    TAKE(V.dzL, nZ); TAKE(V.dzU, nZ); TAKE(V.gf, nZ); TAKE(V.jty, nZ); TAKE(V.sigx, nZ); TAKE(V.gx, nZ);
    TAKE(V.zt, nZ); TAKE(V.dz2, nZ); TAKE(V.one_z, nZ);
    TAKE(C.ceq, nE); TAKE(V.yc, nE); TAKE(V.ytc, nE); TAKE(V.c, nE); TAKE(V.ct, nE); TAKE(V.csoc, nE);
    TAKE(V.ytc2, nE); TAKE(V.zero_e, nE);
    TAKE(C.dl, nI); TAKE(C.du, nI); TAKE(V.s, nI); TAKE(V.vL, nI); TAKE(V.vU, nI); TAKE(V.yd, nI);
    TAKE(V.ds, nI); TAKE(V.dvL, nI); TAKE(V.dvU, nI); TAKE(V.ytd, nI); TAKE(V.dv, nI); TAKE(V.dms, nI);
    TAKE(V.sigs, nI); TAKE(V.gs, nI); TAKE(V.st, nI); TAKE(V.dvt, nI); TAKE(V.dsoc, nI); TAKE(V.ds2, nI);
    TAKE(V.ytd2, nI); TAKE(V.zero_i, nI); TAKE(V.one_i, nI);
#undef TAKE
    V.act = (unsigned char *)calloc(nI + 1, 1);
    ctx_alloc_riccati(&C);
    for (size_t e = 0; e < nZ; e++) V.one_z[e] = 1.0;
    for (int r = 0; r < nI; r++) V.one_i[r] = 1.0;

    /* ---- bounds (...six...py:349-352), relaxed by bound_relax_factor ---- */
    {
        double *lz = V.zt, *uz = V.dz2; /* temporaries */
        flat_to_stage(D, lbx, lz, -INFINITY); flat_to_stage(D, ubx, uz, INFINITY);
        for (size_t e = 0; e < nZ; e++) {
            double l = lz[e], u = uz[e];
            if (!(l <= u)) { rc = -2; goto done; }
            if (l == u) { rc = -3; goto done; } /* fixed variables are not part of this path */
            C.zl[e] = l > -INFINITY ? l - o->bound_relax_factor * fmax(1.0, fabs(l)) : -INFINITY;
            C.zu[e] = u < INFINITY ? u + o->bound_relax_factor * fmax(1.0, fabs(u)) : INFINITY;
        }
        for (int k = 0; k <= N; k++) { /* U_N does not exist */
            if (k < N) continue;
            for (int u = ns; u < nz; u++) { C.zl[k * nz + u] = -INFINITY; C.zu[k * nz + u] = INFINITY; }
        }
        for (int b = 0; b <= N; b++) {
            for (int j = 0; j < ns; j++) {
                double l = lbg[goff(D, b) + j], u = ubg[goff(D, b) + j];
                if (!(l == u) || !isfinite(l)) { rc = -4; goto done; } /* dynamics rows must be equalities */
                C.ceq[b * ns + j] = l;
            }
            for (int qq = 0; qq < M; qq++) {
                int e = b * M + qq;
                if (D->nobs && b == 0) { C.dl[e] = -INFINITY; C.du[e] = INFINITY; V.act[e] = 0; continue; } /* no such rows */
                double l = lbg[goff(D, b) + ns + qq], u = ubg[goff(D, b) + ns + qq];
                if (!(l <= u)) { rc = -2; goto done; }
                if (l == u) { rc = -5; goto done; } /* equality on a distance row: unsupported */
                C.dl[e] = l > -INFINITY ? l - o->bound_relax_factor * fmax(1.0, fabs(l)) : -INFINITY;
                C.du[e] = u < INFINITY ? u + o->bound_relax_factor * fmax(1.0, fabs(u)) : INFINITY;
                V.act[e] = (l > -INFINITY || u < INFINITY) ? 1 : 0;
            }
        }
    }
    /* ---- objective scaling from the gradient at the user's starting point ---- */
    flat_to_stage(D, x0, V.z, 0.0);
    eval_grad(&C, V.z, V.gf);
    {
        double gmax = 0;
        for (size_t e = 0; e < nZ; e++) gmax = fmax(gmax, fabs(V.gf[e]));
        C.df = gmax > o->nlp_scaling_max_gradient ? fmax(o->nlp_scaling_max_gradient / gmax, 1e-8) : 1.0;
    }
    /* ---- initial point: push into bounds, slacks, multipliers ---- */
    for (size_t e = 0; e < nZ; e++) V.z[e] = push_in(V.z[e], C.zl[e], C.zu[e], o->bound_push, o->bound_frac);
    eval_cons(&C, V.z, V.c, V.dv);
    for (int r = 0; r < nI; r++) V.s[r] = V.act[r] ? push_in(V.dv[r], C.dl[r], C.du[r], o->bound_push, o->bound_frac) : V.dv[r];
    for (size_t e = 0; e < nZ; e++) {
        V.zL[e] = C.zl[e] > -INFINITY ? o->bound_mult_init_val : 0.0;
        V.zU[e] = C.zu[e] < INFINITY ? o->bound_mult_init_val : 0.0;
    }
    for (int r = 0; r < nI; r++) {
        V.vL[r] = (V.act[r] && C.dl[r] > -INFINITY) ? o->bound_mult_init_val : 0.0;
        V.vU[r] = (V.act[r] && C.du[r] < INFINITY) ? o->bound_mult_init_val : 0.0;
    }
    /* least-squares equality multipliers: W = 0, Sx = I, Ss = I, zero constraint rhs */
    eval_grad(&C, V.z, V.gf);
    for (size_t e = 0; e < nZ; e++) V.gx[e] = C.df * V.gf[e] - V.zL[e] + V.zU[e];
    for (int r = 0; r < nI; r++) V.gs[r] = -V.vL[r] + V.vU[r];
    {
        int bad = kkt_solve(&C, V.z, V.zero_e, V.zero_i, 0.0, 0.0, V.one_z, V.one_i, V.act, 0.0, V.gx, V.gs,
                            V.zero_e, V.zero_i, V.dz, V.ds, V.yc, V.yd);
        double ymax = 0;
        for (int r = 0; r < nE; r++) ymax = fmax(ymax, fabs(V.yc[r]));
        for (int r = 0; r < nI; r++) ymax = fmax(ymax, fabs(V.yd[r]));
        if (bad || !(ymax <= o->constr_mult_init_max)) {
            memset(V.yc, 0, sizeof(double) * nE); memset(V.yd, 0, sizeof(double) * nI);
        }
    }
    /* ---- main loop ---- */
    double mu = o->mu_init, tau = fmax(o->tau_min, 1.0 - mu);
    filter_t F; F.n = 0; F.evict = 0; F.added = 0; F.added_max = 0;
    double theta_max = -1, theta_min = -1, delta_last = 0.0, f_prev = 0.0;
    int iter = 0, st = ORC_MAX_ITER, n_acc = 0, n_reg = 0, n_resto = 0, n_soc = 0, n_fact = 1, n_ls = 0;
    double E0 = 0, dual_inf = 0, primal_inf = 0, compl0 = 0;
    const double kd = o->kappa_d;
    const double mu_floor = fmin(o->tol, o->compl_inf_tol) / (o->barrier_tol_factor + 1.0);
    int tiny_prev = 0;
    for (;;) {
        /* --- evaluate residuals at the current iterate --- */
        eval_cons(&C, V.z, V.c, V.dv);
        eval_grad(&C, V.z, V.gf);
        eval_jtv(&C, V.z, V.yc, V.yd, V.jty);
        double fcur = eval_obj(&C, V.z);
        primal_inf = 0;
        for (int r = 0; r < nE; r++) primal_inf = fmax(primal_inf, fabs(V.c[r]));
        double nlp_viol = primal_inf;
        for (int r = 0; r < nI; r++) {
            if (!V.act[r]) { V.dms[r] = 0; V.s[r] = V.dv[r]; continue; }
            V.dms[r] = V.dv[r] - V.s[r];
            primal_inf = fmax(primal_inf, fabs(V.dms[r]));
            nlp_viol = fmax(nlp_viol, fmax(C.dl[r] - V.dv[r], V.dv[r] - C.du[r]));
        }
        double ysum = 0, zsum = 0; int ny = nE, nzb = 0;
        for (int r = 0; r < nE; r++) ysum += fabs(V.yc[r]);
        for (int r = 0; r < nI; r++) if (V.act[r]) { ysum += fabs(V.yd[r]); ny++; }
        for (int k = 0; k <= N; k++)
            for (int j = 0; j < (k < N ? nz : ns); j++) {
                int e = k * nz + j;
                if (C.zl[e] > -INFINITY) { zsum += fabs(V.zL[e]); nzb++; }
                if (C.zu[e] < INFINITY) { zsum += fabs(V.zU[e]); nzb++; }
            }
        for (int r = 0; r < nI; r++) if (V.act[r]) {
            if (C.dl[r] > -INFINITY) { zsum += fabs(V.vL[r]); nzb++; }
            if (C.du[r] < INFINITY) { zsum += fabs(V.vU[r]); nzb++; }
        }
        const double smax = 100.0;
        double sd = fmax(smax, (ysum + zsum) / fmax(1, ny + nzb)) / smax;
        double sc = fmax(smax, zsum / fmax(1, nzb)) / smax;
        double Emu;
        for (int pass = 0;; pass++) {
            /* dual infeasibility (with the kappa_d damping of one-sided bounds) and complementarity */
            dual_inf = 0; compl0 = 0; double complmu = 0;
            for (int k = 0; k <= N; k++)
                for (int j = 0; j < (k < N ? nz : ns); j++) {
                    int e = k * nz + j; int hl = C.zl[e] > -INFINITY, hu = C.zu[e] < INFINITY;
                    double r = C.df * V.gf[e] + V.jty[e] - V.zL[e] + V.zU[e];
                    if (hl && !hu) r += kd * mu;
                    if (hu && !hl) r -= kd * mu;
                    dual_inf = fmax(dual_inf, fabs(r));
                    if (hl) { double t = (V.z[e] - C.zl[e]) * V.zL[e]; compl0 = fmax(compl0, fabs(t)); complmu = fmax(complmu, fabs(t - mu)); }
                    if (hu) { double t = (C.zu[e] - V.z[e]) * V.zU[e]; compl0 = fmax(compl0, fabs(t)); complmu = fmax(complmu, fabs(t - mu)); }
                }
            for (int r = 0; r < nI; r++) {
                if (!V.act[r]) continue;
                int hl = C.dl[r] > -INFINITY, hu = C.du[r] < INFINITY;
                double t = -V.yd[r] - V.vL[r] + V.vU[r];
                if (hl && !hu) t += kd * mu;
                if (hu && !hl) t -= kd * mu;
                dual_inf = fmax(dual_inf, fabs(t));
                if (hl) { double u = (V.s[r] - C.dl[r]) * V.vL[r]; compl0 = fmax(compl0, fabs(u)); complmu = fmax(complmu, fabs(u - mu)); }
                if (hu) { double u = (C.du[r] - V.s[r]) * V.vU[r]; compl0 = fmax(compl0, fabs(u)); complmu = fmax(complmu, fabs(u - mu)); }
            }
            E0 = fmax(fmax(dual_inf / sd, primal_inf), compl0 / sc);
            Emu = fmax(fmax(dual_inf / sd, primal_inf), complmu / sc);
            if (pass == 0) {
                /* --- termination tests (at the top of the iteration, as IPOPT does) --- */
                if (E0 <= o->tol && dual_inf / C.df <= o->dual_inf_tol && nlp_viol <= o->constr_viol_tol &&
                    compl0 / C.df <= o->compl_inf_tol) { st = ORC_SOLVED; goto finished; }
                int acc = E0 <= o->acceptable_tol && dual_inf / C.df <= 1e10 && nlp_viol <= 1e-2 &&
                          compl0 / C.df <= 1e-2 &&
                          (iter == 0 || fabs(fcur - f_prev) / fmax(1.0, fabs(fcur)) <= o->acceptable_obj_change_tol);
                n_acc = acc ? n_acc + 1 : 0;
                if (n_acc >= o->acceptable_iter) { st = ORC_ACCEPTABLE; goto finished; }
                if (iter >= o->max_iter) { st = ORC_MAX_ITER; goto finished; }
            }
            /* --- monotone barrier update --- */
            if (!(Emu <= o->barrier_tol_factor * mu) && !(tiny_prev && pass == 0)) break;
            double nm = fmax(fmin(o->kappa_mu * mu, pow(mu, o->theta_mu)), mu_floor);
            if (nm >= mu) break;
            mu = nm; tau = fmax(o->tau_min, 1.0 - mu); F.n = 0; F.added = 0; tiny_prev = 0;
        }
        f_prev = fcur;
        if (trace && iter < max_trace) {
            double *tr = trace + (size_t)iter * ORC_NTRACE;
            tr[ORC_TR_MU] = mu; tr[ORC_TR_ERR] = E0; tr[ORC_TR_THETA] = theta_of(&C, &V, V.c, V.dv, V.s);
            tr[ORC_TR_OBJ] = fcur; tr[ORC_TR_ALPHA_PR] = tr[ORC_TR_ALPHA_DU] = tr[ORC_TR_DELTA_W] = tr[ORC_TR_N_LS] = 0;
        }
        /* --- primal-dual search direction with inertia correction --- */
        for (int k = 0; k <= N; k++)
            for (int j = 0; j < nz; j++) {
                int e = k * nz + j; double sg = 0, gg = C.df * V.gf[e];
                if (k == N && j >= ns) { V.sigx[e] = 1.0; V.gx[e] = 0; continue; }
                int hl = C.zl[e] > -INFINITY, hu = C.zu[e] < INFINITY;
                if (hl) { double sl = V.z[e] - C.zl[e]; sg += V.zL[e] / sl; gg -= mu / sl; }
                if (hu) { double sl = C.zu[e] - V.z[e]; sg += V.zU[e] / sl; gg += mu / sl; }
                if (hl && !hu) gg += kd * mu;
                if (hu && !hl) gg -= kd * mu;
                V.sigx[e] = sg; V.gx[e] = gg;
            }
        for (int r = 0; r < nI; r++) {
            double sg = 0, gg = 0;
            if (V.act[r]) {
                int hl = C.dl[r] > -INFINITY, hu = C.du[r] < INFINITY;
                if (hl) { double sl = V.s[r] - C.dl[r]; sg += V.vL[r] / sl; gg -= mu / sl; }
                if (hu) { double sl = C.du[r] - V.s[r]; sg += V.vU[r] / sl; gg += mu / sl; }
                if (hl && !hu) gg += kd * mu;
                if (hu && !hl) gg -= kd * mu;
            }
            V.sigs[r] = sg; V.gs[r] = gg;
        }
        double delta = 0.0; int need_resto = 0;
        for (;;) {
            n_fact++;
            int bad = kkt_solve(&C, V.z, V.yc, V.yd, C.df, 0.0, V.sigx, V.sigs, V.act, delta, V.gx, V.gs,
                                V.c, V.dms, V.dz, V.ds, V.ytc, V.ytd);
            if (!bad) break;
            if (delta == 0.0) delta = delta_last == 0.0 ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else delta *= (delta_last == 0.0 || 1e5 * delta_last < delta) ? 100.0 : 8.0;
            if (delta > 1e20) { need_resto = 1; break; }
        }
        if (delta > 0 && !need_resto) { delta_last = delta; n_reg++; }
        double alpha = 0, alpha_z = 0; int ls_count = 0;
        double theta = theta_of(&C, &V, V.c, V.dv, V.s);
        if (theta_max < 0) { theta_max = 1e4 * fmax(1.0, theta); theta_min = 1e-4 * fmax(1.0, theta); }
        double phi = barrier_of(&C, &V, V.z, V.s, mu);
        double *dzA = V.dz, *dsA = V.ds, *ytcA = V.ytc, *ytdA = V.ytd; /* accepted direction */
        if (!need_resto) {
            /* directional derivative of the barrier function */
            double gbd = 0;
            for (int k = 0; k <= N; k++)
                for (int j = 0; j < (k < N ? nz : ns); j++) { int e = k * nz + j; gbd += V.gx[e] * V.dz[e]; }
            for (int r = 0; r < nI; r++) if (V.act[r]) gbd += V.gs[r] * V.ds[r];
            double amax = alpha_primal_max(&C, &V, V.z, V.s, V.dz, V.ds, tau);
            /* tiny step: accept without line search */
            double rel = 0;
            for (size_t e = 0; e < nZ; e++) rel = fmax(rel, fabs(V.dz[e]) / (1.0 + fabs(V.z[e])));
            for (int r = 0; r < nI; r++) if (V.act[r]) rel = fmax(rel, fabs(V.ds[r]) / (1.0 + fabs(V.s[r])));
            int tiny = rel < 10.0 * DBL_EPSILON && theta < 1e-4;
            double amin = 1e-5; /* gamma_theta */
            if (gbd < 0) {
                amin = fmin(1e-5, 1e-8 * theta / (-gbd));
                if (theta <= theta_min) amin = fmin(amin, pow(theta, 1.1) / pow(-gbd, 2.3));
            }
            amin *= 0.05;
            int accepted = 0, armijo_step = 0;
            alpha = amax;
            if (tiny) { accepted = 1; tiny_prev = 1; }
            while (!accepted) {
                ls_count++;
                for (size_t e = 0; e < nZ; e++) V.zt[e] = V.z[e] + alpha * V.dz[e];
                for (int r = 0; r < nI; r++) V.st[r] = V.s[r] + alpha * V.ds[r];
                eval_cons(&C, V.zt, V.ct, V.dvt);
                double th_t = theta_of(&C, &V, V.ct, V.dvt, V.st), ph_t = barrier_of(&C, &V, V.zt, V.st, mu);
                int ok = 0, ftype = gbd < 0 && alpha * pow(-gbd, 2.3) > pow(theta, 1.1);
                if (isfinite(th_t) && isfinite(ph_t) && cmp_le(th_t, theta_max, theta)) {
                    if (ftype && theta <= theta_min) ok = cmp_le(ph_t - phi, 1e-8 * alpha * gbd, phi);
                    else ok = cmp_le(th_t, (1 - 1e-5) * theta, theta) || cmp_le(ph_t - phi, -1e-8 * theta, phi);
                    if (ok) ok = filter_ok(&F, th_t, ph_t);
                }
                if (ok) { accepted = 1; armijo_step = ftype && theta <= theta_min; break; }
                /* second-order correction on the first trial point */
                if (ls_count == 1 && th_t >= theta && o->max_soc > 0) {
                    double th_old = 0, th_tr = th_t, a_soc = alpha; int cnt = 0;
                    memcpy(V.csoc, V.c, sizeof(double) * nE); memcpy(V.dsoc, V.dms, sizeof(double) * nI);
                    const double *ctrial = V.ct, *dvtrial = V.dvt, *strial = V.st;
                    while (cnt < o->max_soc && !accepted && (cnt == 0 || th_tr <= 0.99 * th_old)) {
                        th_old = th_tr;
                        for (int r = 0; r < nE; r++) V.csoc[r] = a_soc * V.csoc[r] + ctrial[r];
                        for (int r = 0; r < nI; r++) V.dsoc[r] = V.act[r] ? a_soc * V.dsoc[r] + (dvtrial[r] - strial[r]) : 0;
                        n_fact++; n_soc++;
                        if (kkt_solve(&C, V.z, V.yc, V.yd, C.df, 0.0, V.sigx, V.sigs, V.act, delta, V.gx, V.gs,
                                      V.csoc, V.dsoc, V.dz2, V.ds2, V.ytc2, V.ytd2)) break;
                        a_soc = alpha_primal_max(&C, &V, V.z, V.s, V.dz2, V.ds2, tau);
                        for (size_t e = 0; e < nZ; e++) V.zt[e] = V.z[e] + a_soc * V.dz2[e];
                        for (int r = 0; r < nI; r++) V.st[r] = V.s[r] + a_soc * V.ds2[r];
                        eval_cons(&C, V.zt, V.ct, V.dvt);
                        double th2 = theta_of(&C, &V, V.ct, V.dvt, V.st), ph2 = barrier_of(&C, &V, V.zt, V.st, mu);
                        int ok2 = 0;
                        if (isfinite(th2) && isfinite(ph2) && cmp_le(th2, theta_max, theta)) {
                            if (ftype && theta <= theta_min) ok2 = cmp_le(ph2 - phi, 1e-8 * alpha * gbd, phi);
                            else ok2 = cmp_le(th2, (1 - 1e-5) * theta, theta) || cmp_le(ph2 - phi, -1e-8 * theta, phi);
                            if (ok2) ok2 = filter_ok(&F, th2, ph2);
                        }
                        if (ok2) {
                            accepted = 1; armijo_step = ftype && theta <= theta_min;
                            dzA = V.dz2; dsA = V.ds2; ytcA = V.ytc2; ytdA = V.ytd2; alpha = a_soc;
                        } else { cnt++; th_tr = th2; }
                    }
                    if (accepted) break;
                }
                alpha *= 0.5;
                if (alpha < amin) break;
            }
            if (!accepted) need_resto = 1;
            else {
                if (!tiny && !armijo_step) filter_add(&F, (1 - 1e-5) * theta, phi - 1e-8 * theta);
                if (!tiny) tiny_prev = 0;
                /* bound-multiplier steps from the accepted direction */
                alpha_z = 1.0;
                for (int k = 0; k <= N; k++)
                    for (int j = 0; j < (k < N ? nz : ns); j++) {
                        int e = k * nz + j;
                        if (C.zl[e] > -INFINITY) {
                            double sl = V.z[e] - C.zl[e];
                            V.dzL[e] = mu / sl - V.zL[e] - V.zL[e] / sl * dzA[e];
                            if (V.dzL[e] < 0) alpha_z = fmin(alpha_z, -tau * V.zL[e] / V.dzL[e]);
                        }
                        if (C.zu[e] < INFINITY) {
                            double sl = C.zu[e] - V.z[e];
                            V.dzU[e] = mu / sl - V.zU[e] + V.zU[e] / sl * dzA[e];
                            if (V.dzU[e] < 0) alpha_z = fmin(alpha_z, -tau * V.zU[e] / V.dzU[e]);
                        }
                    }
                for (int r = 0; r < nI; r++) {
                    if (!V.act[r]) continue;
                    if (C.dl[r] > -INFINITY) {
                        double sl = V.s[r] - C.dl[r];
                        V.dvL[r] = mu / sl - V.vL[r] - V.vL[r] / sl * dsA[r];
                        if (V.dvL[r] < 0) alpha_z = fmin(alpha_z, -tau * V.vL[r] / V.dvL[r]);
                    }
                    if (C.du[r] < INFINITY) {
                        double sl = C.du[r] - V.s[r];
                        V.dvU[r] = mu / sl - V.vU[r] + V.vU[r] / sl * dsA[r];
                        if (V.dvU[r] < 0) alpha_z = fmin(alpha_z, -tau * V.vU[r] / V.dvU[r]);
                    }
                }
                /* accept */
                const double ks = o->kappa_sigma;
                for (int k = 0; k <= N; k++)
                    for (int j = 0; j < (k < N ? nz : ns); j++) {
                        int e = k * nz + j;
                        V.z[e] += alpha * dzA[e];
                        if (C.zl[e] > -INFINITY) {
                            double sl = V.z[e] - C.zl[e], zn = V.zL[e] + alpha_z * V.dzL[e];
                            V.zL[e] = fmax(fmin(zn, ks * mu / sl), mu / (ks * sl));
                        }
                        if (C.zu[e] < INFINITY) {
                            double sl = C.zu[e] - V.z[e], zn = V.zU[e] + alpha_z * V.dzU[e];
                            V.zU[e] = fmax(fmin(zn, ks * mu / sl), mu / (ks * sl));
                        }
                    }
                for (int r = 0; r < nE; r++) V.yc[r] += alpha * (ytcA[r] - V.yc[r]);
                for (int r = 0; r < nI; r++) {
                    if (!V.act[r]) continue;
                    V.s[r] += alpha * dsA[r];
                    V.yd[r] += alpha * (ytdA[r] - V.yd[r]);
                    if (C.dl[r] > -INFINITY) {
                        double sl = V.s[r] - C.dl[r], zn = V.vL[r] + alpha_z * V.dvL[r];
                        V.vL[r] = fmax(fmin(zn, ks * mu / sl), mu / (ks * sl));
                    }
                    if (C.du[r] < INFINITY) {
                        double sl = C.du[r] - V.s[r], zn = V.vU[r] + alpha_z * V.dvU[r];
                        V.vU[r] = fmax(fmin(zn, ks * mu / sl), mu / (ks * sl));
                    }
                }
            }
        }
        if (need_resto) {
            /* Bounded substitute for IPOPT's restoration phase: damped Newton steps on the barrier
             * feasibility problem  min -mu sum ln(slack) + sqrt(mu)/2 |dx|^2  s.t. linearised rows,
             * accepted on an Armijo test of theta alone, until theta <= 0.9 theta_R and the point
             * is acceptable to the filter; exits INFEASIBLE when no progress is possible. */
            n_resto++;
            if (getenv("ORC_DEBUG")) fprintf(stderr, "enter resto iter %d theta %.3e mu %.3e delta %.3e ls %d alpha %.3e\n", iter, theta, mu, delta, ls_count, alpha);
            filter_add(&F, (1 - 1e-5) * theta, phi - 1e-8 * theta);
            double thR = theta, zeta = sqrt(mu); int ok = 0, r_it;
            /* first candidate: project onto the dynamics by a forward rollout of the current controls
             * (multiple shooting: X_0 = xbar0, X_{k+1} = X_k + T f(X_k,U_k) makes every equality row zero),
             * slacks reset from the distances; only a colliding rollout leaves infeasibility behind */
            {
                memcpy(V.zt, V.z, sizeof(double) * nZ);
                for (int j = 0; j < ns; j++) V.zt[j] = push_in(p[j] + C.ceq[j], C.zl[j], C.zu[j], o->bound_push, o->bound_frac);
                for (int k = 0; k < N; k++) {
                    const double *zk = V.zt + k * nz; double *zn = V.zt + (k + 1) * nz; const double *ce = C.ceq + (k + 1) * ns;
                    for (int i = 0; i < D->Nr; i++) {
                        double th = zk[3 * i + 2], v = zk[ns + 2 * i], om = zk[ns + 2 * i + 1];
                        zn[3 * i] = zk[3 * i] + D->T * v * cos(th) + ce[3 * i];
                        zn[3 * i + 1] = zk[3 * i + 1] + D->T * v * sin(th) + ce[3 * i + 1];
                        zn[3 * i + 2] = th + D->T * om + ce[3 * i + 2];
                        for (int cc = 0; cc < 3; cc++) {
                            int e = (k + 1) * nz + 3 * i + cc;
                            zn[3 * i + cc] = push_in(zn[3 * i + cc], C.zl[e], C.zu[e], o->bound_push, o->bound_frac);
                        }
                    }
                }
                eval_cons(&C, V.zt, V.ct, V.dvt);
                for (int r = 0; r < nI; r++) V.st[r] = V.act[r] ? push_in(V.dvt[r], C.dl[r], C.du[r], o->bound_push, o->bound_frac) : V.dvt[r];
                double th_p = theta_of(&C, &V, V.ct, V.dvt, V.st);
                if (th_p < thR) {   /* keep the projected point as the start of the Newton restoration */
                    memcpy(V.z, V.zt, sizeof(double) * nZ);
                    for (int r = 0; r < nI; r++) if (V.act[r]) V.s[r] = V.st[r];
                }
            }
            for (r_it = 0; r_it < o->max_resto_iter; r_it++) {
                eval_cons(&C, V.z, V.c, V.dv);
                for (int r = 0; r < nI; r++) V.dms[r] = V.act[r] ? V.dv[r] - V.s[r] : 0;
                double th = theta_of(&C, &V, V.c, V.dv, V.s);
                if ((th <= 0.9 * thR || th <= 1e-9) && filter_ok(&F, th, barrier_of(&C, &V, V.z, V.s, mu))) { ok = 1; break; }
                for (int k = 0; k <= N; k++)
                    for (int j = 0; j < nz; j++) {
                        int e = k * nz + j; double sg = 0, gg = 0;
                        if (!(k == N && j >= ns)) {
                            if (C.zl[e] > -INFINITY) { double sl = V.z[e] - C.zl[e]; sg += mu / (sl * sl); gg -= mu / sl; }
                            if (C.zu[e] < INFINITY) { double sl = C.zu[e] - V.z[e]; sg += mu / (sl * sl); gg += mu / sl; }
                        }
                        V.sigx[e] = sg; V.gx[e] = gg;
                    }
                for (int r = 0; r < nI; r++) {
                    double sg = 0, gg = 0;
                    if (V.act[r]) {
                        if (C.dl[r] > -INFINITY) { double sl = V.s[r] - C.dl[r]; sg += mu / (sl * sl); gg -= mu / sl; }
                        if (C.du[r] < INFINITY) { double sl = C.du[r] - V.s[r]; sg += mu / (sl * sl); gg += mu / sl; }
                    }
                    V.sigs[r] = sg; V.gs[r] = gg;
                }
                n_fact++;
                if (kkt_solve(&C, V.z, V.zero_e, V.zero_i, 0.0, zeta, V.sigx, V.sigs, V.act, 0.0, V.gx, V.gs,
                              V.c, V.dms, V.dz, V.ds, V.ytc, V.ytd)) break;
                double a = alpha_primal_max(&C, &V, V.z, V.s, V.dz, V.ds, tau); int got = 0;
                while (a > 1e-12) {
                    for (size_t e = 0; e < nZ; e++) V.zt[e] = V.z[e] + a * V.dz[e];
                    for (int r = 0; r < nI; r++) V.st[r] = V.s[r] + a * V.ds[r];
                    eval_cons(&C, V.zt, V.ct, V.dvt);
                    double th_t = theta_of(&C, &V, V.ct, V.dvt, V.st);
                    if (th_t <= (1 - 1e-4 * a) * th) { got = 1; break; }
                    a *= 0.5;
                }
                if (getenv("ORC_DEBUG")) fprintf(stderr, "resto it %d th %.3e a %.3e got %d amax %.3e\n", r_it, th, a, got, alpha_primal_max(&C, &V, V.z, V.s, V.dz, V.ds, tau));
                if (!got) break;
                memcpy(V.z, V.zt, sizeof(double) * nZ);
                for (int r = 0; r < nI; r++) if (V.act[r]) V.s[r] = V.st[r];
                if (th - theta_of(&C, &V, V.ct, V.dvt, V.st) < 1e-14 * fmax(1.0, th)) break; /* stalled */
            }
            if (!ok) { st = ORC_INFEASIBLE; iter++; goto finished; }
            /* multipliers after restoration: equality multipliers reset, bound multipliers kept in the
             * kappa_sigma corridor of the new slacks */
            memset(V.yc, 0, sizeof(double) * nE); memset(V.yd, 0, sizeof(double) * nI);
            const double ks = o->kappa_sigma;
            for (int k = 0; k <= N; k++)
                for (int j = 0; j < (k < N ? nz : ns); j++) {
                    int e = k * nz + j;
                    if (C.zl[e] > -INFINITY) { double sl = V.z[e] - C.zl[e]; V.zL[e] = fmax(fmin(V.zL[e], ks * mu / sl), mu / (ks * sl)); }
                    if (C.zu[e] < INFINITY) { double sl = C.zu[e] - V.z[e]; V.zU[e] = fmax(fmin(V.zU[e], ks * mu / sl), mu / (ks * sl)); }
                }
            for (int r = 0; r < nI; r++) {
                if (!V.act[r]) continue;
                if (C.dl[r] > -INFINITY) { double sl = V.s[r] - C.dl[r]; V.vL[r] = fmax(fmin(V.vL[r], ks * mu / sl), mu / (ks * sl)); }
                if (C.du[r] < INFINITY) { double sl = C.du[r] - V.s[r]; V.vU[r] = fmax(fmin(V.vU[r], ks * mu / sl), mu / (ks * sl)); }
            }
            alpha = 0; alpha_z = 0;
        }
        n_ls += ls_count;
        if (trace && iter < max_trace) {
            double *tr = trace + (size_t)iter * ORC_NTRACE;
            tr[ORC_TR_ALPHA_PR] = alpha; tr[ORC_TR_ALPHA_DU] = alpha_z; tr[ORC_TR_DELTA_W] = delta; tr[ORC_TR_N_LS] = ls_count;
        }
        iter++;
    }
finished:
    /* ---- outputs in the reference layout, multipliers in CasADi's sign convention ---- */
    stage_to_flat(D, V.z, x);
    if (f) *f = eval_obj(&C, V.z);
    if (g && !D->nobs) orc_eval(d, x, p, NULL, NULL, NULL, g, NULL, NULL);
    if (g && D->nobs) { /* the obstacle family is outside orc_eval: rows from the stage-layout evaluation */
        double *cc = (double *)malloc(sizeof(double) * (nE + nI + 1)), *dd = cc + nE;
        const double *ce0 = C.ceq; double *zero = (double *)calloc(nE, sizeof(double));
        C.ceq = zero; eval_cons(&C, V.z, cc, dd); C.ceq = (double *)ce0;
        for (int b = 0; b <= N; b++) {
            for (int j = 0; j < ns; j++) g[goff(D, b) + j] = cc[b * ns + j];
            if (b > 0) for (int qq = 0; qq < M; qq++) g[goff(D, b) + ns + qq] = dd[b * M + qq];
        }
        free(cc); free(zero);
    }
    if (lam_x) {
        for (size_t e = 0; e < nZ; e++) V.zt[e] = (V.zU[e] - V.zL[e]) / C.df;
        stage_to_flat(D, V.zt, lam_x);
    }
    if (lam_g)
        for (int b = 0; b <= N; b++) {
            for (int j = 0; j < ns; j++) lam_g[goff(D, b) + j] = V.yc[b * ns + j] / C.df;
            if (!(D->nobs && b == 0)) for (int qq = 0; qq < M; qq++) lam_g[goff(D, b) + ns + qq] = V.yd[b * M + qq] / C.df;
        }
    if (status) *status = st;
    if (iters) *iters = iter;
    if (stats) {
        stats[ORC_ST_KKT_ERR] = E0; stats[ORC_ST_PRIMAL_INF] = primal_inf; stats[ORC_ST_DUAL_INF] = dual_inf;
        stats[ORC_ST_COMPL] = compl0; stats[ORC_ST_MU] = mu; stats[ORC_ST_N_REG] = n_reg;
        stats[ORC_ST_N_RESTO] = n_resto; stats[ORC_ST_N_SOC] = n_soc; stats[ORC_ST_N_FACTOR] = n_fact;
        stats[ORC_ST_N_LS] = n_ls; stats[ORC_ST_FILTER_EVICT] = F.evict; stats[ORC_ST_FILTER_ADDS] = F.added_max;
    }
done:
    ctx_free_riccati(&C); free(V.act); free(pool); dims_free(&C.D);
    return rc;
}

int orc_solve(const orc_desc *d, const orc_opts *o, const double *x0, const double *p,
              const double *lbx, const double *ubx, const double *lbg, const double *ubg,
              double *x, double *f, double *g, double *lam_x, double *lam_g,
              int *status, int *iters, double *stats, double *trace, int max_trace)
{
    if (!d || !o || !x0 || !p || !lbx || !ubx || !lbg || !ubg || !x) return -1;
    if (d->Nr < 1 || d->Nr > 64 || d->N < 1) return -1;
    return solve_one(d, o, x0, p, lbx, ubx, lbg, ubg, x, f, g, lam_x, lam_g, status, iters, stats, trace, max_trace);
}

int orc_solve_batch(const orc_desc *d, const orc_opts *o, int B, const double *x0, const double *p,
                    const double *lbx, const double *ubx, const double *lbg, const double *ubg,
                    int bounds_batched, double *x, double *f, double *g, double *lam_x,
                    double *lam_g, int *status, int *iters, double *stats, int nthreads)
{
    if (!d || !o || B < 0) return -1;
    const int n = orc_n(d), mg = orc_mg(d), np = 6 * d->Nr;
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < B; b++) {
        size_t ob = bounds_batched ? (size_t)b : 0;
        int r = orc_solve(d, o, x0 + (size_t)b * n, p + (size_t)b * np, lbx + ob * n, ubx + ob * n,
                          lbg + ob * mg, ubg + ob * mg, x + (size_t)b * n, f ? f + b : NULL,
                          g ? g + (size_t)b * mg : NULL, lam_x ? lam_x + (size_t)b * n : NULL,
                          lam_g ? lam_g + (size_t)b * mg : NULL, status ? status + b : NULL,
                          iters ? iters + b : NULL, stats ? stats + (size_t)b * ORC_NSTATS : NULL, NULL, 0);
        if (r) {
#ifdef _OPENMP
#pragma omp critical
#endif
            err = r;
        }
    }
    return err;
}

// Thread-per-instance solver for small generic optimal-control problems.
//
// The reference's mpc_pose_control_casadi.py:22-114 (BASELINE.json configs[0]) is CasADi's direct-multiple-shooting demo: the
// Van der Pol oscillator  x1' = (1 - x2^2) x1 - x2 + u,  x2' = x1,  running cost x1^2 + x2^2 + u^2, T = 10, N = 20 intervals, each
// integrated by 4 RK4 steps together with its cost quadrature (:45-59); decision vector INTERLEAVED [X_0, U_0, X_1, U_1, ..., X_N]
// (:77-106); the initial state is fixed through lbw = ubw (:79-80), x1 >= -0.25 (:98), |u| <= 1 (:88-89); g = F(X_k, U_k) - X_{k+1}
// (:104); one solver(x0=, lbx=, ubx=, lbg=, ubg=) call with IPOPT's defaults (:110-113).
//
// Same interior-point driver (ipm_driver.cuh) as the unicycle solvers; what differs is the stage: dense dynamics Jacobians and an
// exact Hessian of the Lagrangian THROUGH the RK4 integrator, obtained with second-order forward-mode AD (hyper-dual numbers over
// the nx + nu stage inputs) in place of CasADi's AD.  A stage is 3 variables, so one thread runs a whole instance (all matrices
// are 3 x 3 or smaller and live in registers / local memory); a batch is one thread per instance.  The model is a functor
// (VanDerPol below): another small OCP only needs its own f(x, u) -> (xdot, L).
#pragma once
#include "nmpc_internal.h"
#include "ipm_driver.cuh"
#include "solver_body.cuh"
#include "bounds_prep.cuh"

// value, gradient and symmetric Hessian w.r.t. NV inputs (packed upper triangle, row-major)
template <int NV>
struct Jet2 {
    static constexpr int NH = NV * (NV + 1) / 2;
    double v, g[NV], h[NH];
    static __host__ __device__ constexpr int hi(int i, int j) { return i * NV - i * (i - 1) / 2 + (j - i); }   // i <= j
    static __host__ __device__ Jet2 constant(double c)
    {
        Jet2 r; r.v = c;
        for (int i = 0; i < NV; i++) r.g[i] = 0.0;
        for (int i = 0; i < NH; i++) r.h[i] = 0.0;
        return r;
    }
    static __host__ __device__ Jet2 variable(double c, int k) { Jet2 r = constant(c); r.g[k] = 1.0; return r; }
};
template <int NV> __host__ __device__ inline Jet2<NV> operator+(const Jet2<NV> &a, const Jet2<NV> &b)
{
    Jet2<NV> r; r.v = a.v + b.v;
    for (int i = 0; i < NV; i++) r.g[i] = a.g[i] + b.g[i];
    for (int i = 0; i < Jet2<NV>::NH; i++) r.h[i] = a.h[i] + b.h[i];
    return r;
}
template <int NV> __host__ __device__ inline Jet2<NV> operator-(const Jet2<NV> &a, const Jet2<NV> &b)
{
    Jet2<NV> r; r.v = a.v - b.v;
    for (int i = 0; i < NV; i++) r.g[i] = a.g[i] - b.g[i];
    for (int i = 0; i < Jet2<NV>::NH; i++) r.h[i] = a.h[i] - b.h[i];
    return r;
}
template <int NV> __host__ __device__ inline Jet2<NV> operator*(double c, const Jet2<NV> &a)
{
    Jet2<NV> r; r.v = c * a.v;
    for (int i = 0; i < NV; i++) r.g[i] = c * a.g[i];
    for (int i = 0; i < Jet2<NV>::NH; i++) r.h[i] = c * a.h[i];
    return r;
}
template <int NV> __host__ __device__ inline Jet2<NV> operator*(const Jet2<NV> &a, const Jet2<NV> &b)
{
    Jet2<NV> r; r.v = a.v * b.v;
    for (int i = 0; i < NV; i++) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
    for (int i = 0; i < NV; i++)
        for (int j = i; j < NV; j++)
            r.h[Jet2<NV>::hi(i, j)] = a.h[Jet2<NV>::hi(i, j)] * b.v + a.v * b.h[Jet2<NV>::hi(i, j)] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
    return r;
}

// mpc_pose_control_casadi.py:25-35
struct VanDerPol {
    static constexpr int NX = 2, NU = 1;
    template <class S> static __host__ __device__ void f(const S *x, const S *u, S *xdot, S &L)
    {
        const S one = S::constant(1.0);
        xdot[0] = (one - x[1] * x[1]) * x[0] - x[1] + u[0];
        xdot[1] = x[0];
        L = x[0] * x[0] + x[1] * x[1] + u[0] * u[0];
    }
};

// one shooting interval: M fixed RK4 steps of size DT on the state and the cost quadrature (:45-59)
template <class Model, class S>
__host__ __device__ inline void rk4_interval(const S *x0, const S *u, int M, double DT, S *xf, S &qf)
{
    constexpr int NX = Model::NX;
    S X[NX], k1[NX], k2[NX], k3[NX], k4[NX], t[NX], q1, q2, q3, q4;
    for (int i = 0; i < NX; i++) X[i] = x0[i];
    qf = S::constant(0.0);
    for (int j = 0; j < M; j++) {
        Model::f(X, u, k1, q1);
        for (int i = 0; i < NX; i++) t[i] = X[i] + (DT / 2) * k1[i];
        Model::f(t, u, k2, q2);
        for (int i = 0; i < NX; i++) t[i] = X[i] + (DT / 2) * k2[i];
        Model::f(t, u, k3, q3);
        for (int i = 0; i < NX; i++) t[i] = X[i] + DT * k3[i];
        Model::f(t, u, k4, q4);
        for (int i = 0; i < NX; i++) X[i] = X[i] + (DT / 6) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
        qf = qf + (DT / 6) * (q1 + 2.0 * q2 + 2.0 * q3 + q4);
    }
    for (int i = 0; i < NX; i++) xf[i] = X[i];
}

template <class Model>
struct ThreadSolver {
    static constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU, LD = 4;   // LD: padded row length (NZ <= 4)
    static_assert(NZ <= LD, "stage vector must fit the padded row");
    typedef Jet2<NZ> J2;
    typedef WarpSolver<1> WS;
    enum Row {
        R_Z, R_ZL, R_ZU, R_BL, R_BU, R_DZ, R_DZ2, R_GX, R_YC, R_YTC, R_YTC2, R_RC, R_CSOC, R_LIN, R_DGV, R_ZT,
        R_DS, R_DS2, R_YD, R_YTD, R_YTD2,   // no inequality rows in this family: selectors only
        R_COUNT = R_DS
    };
    struct EvalOut { double pinf, viol, dinf, c0, cmu, ysum, zsum, theta, f, slog, sdamp; };
    struct StepInfo { double ap, az, gbd, tiny; };
    // per-stage cache of the discretised model (doubles per stage)
    enum { C_F = 0, C_J = C_F + NX, C_HF = C_J + NX * NZ, C_Q = C_HF + NX * J2::NH, C_GQ = C_Q + 1, C_HQ = C_GQ + NZ,
           C_P = C_HQ + J2::NH, C_K = C_P + NX * NX, C_KF = C_K + NU * NX, C_H = C_KF + NU, C_COUNT = C_H + J2::NH };

    const NmpcSolveParams &P;
    double *ws, *cache;
    int N, S, inst, fn, rk_steps;
    bool fixed0;
    double DT, df, ny_nzb, nzb_cnt;
    int n_reg, n_resto, n_soc, n_fact, n_ls;
    double fth[NMPC_FILTER_CAP], fph[NMPC_FILTER_CAP];
    int bad;

    static NMPC_HD long long ws_doubles(int N) { return (long long)(R_COUNT * LD + C_COUNT) * (N + 1); }

    __device__ ThreadSolver(const NmpcSolveParams &p, double *wsp) : P(p), ws(wsp) {}
    static __device__ __forceinline__ void tsync() {}
    __device__ __forceinline__ bool is_lead() const { return true; }
    __device__ __forceinline__ double *row(int r, int k) const { return ws + ((long long)r * S + k) * LD; }
    __device__ __forceinline__ double *cst(int k) const { return cache + (long long)k * C_COUNT; }
    // variable l of stage k takes part in the optimisation (the fixed initial state and U_N do not)
    __device__ __forceinline__ bool valid(int k, int l) const { return l < (k < N ? NZ : NX) && !(fixed0 && k == 0 && l < NX); }

    __device__ void setup(int instance)
    {
        inst = instance; N = P.N; S = N + 1; rk_steps = P.rk_steps; DT = P.T / N / rk_steps;
        cache = ws + (long long)R_COUNT * LD * S;
        df = 1.0; fn = 0; bad = 0; fixed0 = false;
        n_reg = n_resto = n_soc = n_fact = n_ls = 0;
    }
    __device__ bool bounds_rejected()
    {
        // flat interleaved bounds (:79-99); a fixed INITIAL STATE (lbw == ubw on X_0) becomes a parameter, as IPOPT's default
        // fixed_variable_treatment does; any other fixed variable or lb > ub is rejected
        const long long n = (long long)NZ * N + NX;
        const double *lb = P.lbx + (P.bounds_batched ? inst * n : 0), *ub = P.ubx + (P.bounds_batched ? inst * n : 0);
        int nfix0 = 0;
        for (int l = 0; l < NX; l++) nfix0 += lb[l] == ub[l];
        fixed0 = nfix0 == NX;
        int err = fixed0 ? 0 : NMPC_ENOTSUP;   // the initial state must be given (fixed through its bounds, :79-80)
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < (k < N ? NZ : NX); l++) {
                const double lo = lb[k * NZ + l], hi = ub[k * NZ + l];
                if (!(lo <= hi)) err = err ? err : NMPC_EBOUNDS;
                else if (lo == hi && !(k == 0 && l < NX)) err = err ? err : NMPC_ENOTSUP;
                const bool par = fixed0 && k == 0 && l < NX;
                row(R_BL, k)[l] = par ? -NMPC_INF : nmpc_relax_lo(lo, P.o.bound_relax_factor);
                row(R_BU, k)[l] = par ? NMPC_INF : nmpc_relax_hi(hi, P.o.bound_relax_factor);
            }
        const long long mg = (long long)NX * N;
        const double *lg = P.lbg + (P.bounds_batched ? inst * mg : 0), *ug = P.ubg + (P.bounds_batched ? inst * mg : 0);
        for (long long r = 0; r < mg; r++)
            if (!(lg[r] == ug[r]) || !(lg[r] > -NMPC_INF && lg[r] < NMPC_INF)) err = err ? err : NMPC_ENOTSUP;   // shooting rows are equalities
        if (err) {
            if (P.status) P.status[inst] = err;
            if (P.iters) P.iters[inst] = 0;
            return true;
        }
        return false;
    }

    // discretised model of stage k at the stage vector z: values, Jacobian, Hessians (into the stage cache)
    __device__ void model(int k, const double *z)
    {
        J2 x[NX], u[NU], xf[NX], qf;
        for (int i = 0; i < NX; i++) x[i] = J2::variable(z[i], i);
        for (int i = 0; i < NU; i++) u[i] = J2::variable(z[NX + i], NX + i);
        rk4_interval<Model, J2>(x, u, rk_steps, DT, xf, qf);
        double *c = cst(k);
        for (int i = 0; i < NX; i++) {
            c[C_F + i] = xf[i].v;
            for (int l = 0; l < NZ; l++) c[C_J + i * NZ + l] = xf[i].g[l];
            for (int e = 0; e < J2::NH; e++) c[C_HF + i * J2::NH + e] = xf[i].h[e];
        }
        c[C_Q] = qf.v;
        for (int l = 0; l < NZ; l++) c[C_GQ + l] = qf.g[l];
        for (int e = 0; e < J2::NH; e++) c[C_HQ + e] = qf.h[e];
    }

    __device__ __noinline__ void init_point()
    {
        const long long n = (long long)NZ * N + NX;
        const double *x0 = P.x0 + inst * n;
        const nmpc_opts &o = P.o;
        const double *lb = P.lbx + (P.bounds_batched ? inst * n : 0);
        double gmax = 0.0, cnt_z = 0.0;
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < LD; l++) {
                double z = (l < (k < N ? NZ : NX)) ? x0[k * NZ + l] : 0.0;
                if (fixed0 && k == 0 && l < NX) z = lb[l];
                row(R_Z, k)[l] = z;
            }
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < LD; l++) {
                double zl = 0.0, zu = 0.0;
                if (valid(k, l)) {
                    const double lo = row(R_BL, k)[l], hi = row(R_BU, k)[l];
                    row(R_Z, k)[l] = WS::push_in(row(R_Z, k)[l], lo, hi, o.bound_push, o.bound_frac);
                    if (lo > -NMPC_INF) { zl = o.bound_mult_init_val; cnt_z += 1.0; }
                    if (hi < NMPC_INF) { zu = o.bound_mult_init_val; cnt_z += 1.0; }
                }
                row(R_ZL, k)[l] = zl; row(R_ZU, k)[l] = zu; row(R_YC, k)[l] = 0.0; row(R_CSOC, k)[l] = 0.0;
            }
        for (int k = 0; k < N; k++) {   // objective scaling from the gradient at the starting point
            model(k, row(R_Z, k));
            for (int l = 0; l < NZ; l++)
                if (valid(k, l)) gmax = fmax(gmax, fabs(cst(k)[C_GQ + l]));
        }
        df = gmax > o.nlp_scaling_max_gradient ? fmax(o.nlp_scaling_max_gradient / gmax, 1e-8) : 1.0;
        nzb_cnt = cnt_z; ny_nzb = (double)(NX * N) + cnt_z;
    }

    __device__ __noinline__ void eval(bool full, double mu, double alpha, int rdz, int rds, bool trial, bool socacc, double asoc, EvalOut &E)
    {
        const double kd = P.o.kappa_d;
        const int rz = trial ? R_ZT : R_Z;
        if (trial)
            for (int k = 0; k <= N; k++)
                for (int l = 0; l < LD; l++) row(R_ZT, k)[l] = row(R_Z, k)[l] + (valid(k, l) ? alpha * row(rdz, k)[l] : 0.0);
        double pinf = 0, viol = 0, dinf = 0, c0 = 0, cmu = 0, ysum = 0, zsum = 0, th = 0, fo = 0, sdamp = 0, slog = 0;
        for (int k = 0; k < N; k++) model(k, row(rz, k));
        for (int k = 0; k < N; k++) {   // shooting rows  F(X_k, U_k) - X_{k+1}
            for (int i = 0; i < NX; i++) {
                const double c = cst(k)[C_F + i] - row(rz, k + 1)[i];
                pinf = fmax(pinf, fabs(c)); th += fabs(c); viol = fmax(viol, fabs(c));
                if (socacc) row(R_CSOC, k + 1)[i] = asoc * row(R_CSOC, k + 1)[i] + c;
                if (full) ysum += fabs(row(R_YC, k + 1)[i]);
            }
            fo += cst(k)[C_Q];
        }
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < NZ; l++) {
                if (!valid(k, l)) continue;
                const double zk = row(rz, k)[l], lo = row(R_BL, k)[l], hi = row(R_BU, k)[l];
                const bool hl = lo > -NMPC_INF, hu = hi < NMPC_INF;
                if (hl) slog += log(zk - lo);
                if (hu) slog += log(hi - zk);
                if (hl && !hu) sdamp += zk - lo;
                if (hu && !hl) sdamp += hi - zk;
                if (full) {
                    const double zl = row(R_ZL, k)[l], zu = row(R_ZU, k)[l];
                    double r = -zl + zu;
                    if (hl && !hu) r += kd * mu;
                    if (hu && !hl) r -= kd * mu;
                    if (k < N) {
                        r += df * cst(k)[C_GQ + l];
                        for (int i = 0; i < NX; i++) r += cst(k)[C_J + i * NZ + l] * row(R_YC, k + 1)[i];
                    }
                    if (l < NX && k >= 1) r -= row(R_YC, k)[l];
                    dinf = fmax(dinf, fabs(r));
                    if (hl) { const double u = (zk - lo) * zl; c0 = fmax(c0, fabs(u)); cmu = fmax(cmu, fabs(u - mu)); zsum += fabs(zl); }
                    if (hu) { const double u = (hi - zk) * zu; c0 = fmax(c0, fabs(u)); cmu = fmax(cmu, fabs(u - mu)); zsum += fabs(zu); }
                }
            }
        E.pinf = pinf; E.viol = viol; E.theta = th; E.f = fo; E.slog = slog; E.sdamp = sdamp;
        if (full) { E.dinf = dinf; E.c0 = c0; E.cmu = cmu; E.ysum = ysum; E.zsum = zsum; }
    }

    __device__ bool factor_m(int mode, double mu, double delta, bool soc) { return factor(mode, mu, delta, soc); }
    // backward Riccati sweep with dense 3 x 3 stage matrices; false = a control pivot <= 0 (wrong inertia)
    __device__ __noinline__ bool factor(int mode, double mu, double delta, bool soc)
    {
        const double zeta = mode == 2 ? sqrt(mu) : 0.0, kd = P.o.kappa_d;
        n_fact++;
        for (int k = 0; k <= N; k++) {
            if (k < N) model(k, row(R_Z, k));
            double *c = cst(k < N ? k : N);
            for (int l = 0; l < LD; l++) {
                double sig = 0.0, gx = 0.0, dg = 0.0;
                if (valid(k, l)) {
                    const double g0 = k < N ? df * cst(k)[C_GQ + l] : 0.0;
                    BlockSolver::sig_g(mode, kd, row(R_Z, k)[l], row(R_BL, k)[l], row(R_BU, k)[l], row(R_ZL, k)[l], row(R_ZU, k)[l], mu, g0, sig, gx);
                    dg = sig + delta + zeta;
                }
                row(R_GX, k)[l] = gx; row(R_DGV, k)[l] = dg;
            }
            if (k < N) {
                for (int e = 0; e < J2::NH; e++) {   // Hessian of the Lagrangian of stage k (exact, through the RK4 steps)
                    double h = 0.0;
                    if (mode == 0) {
                        h = df * c[C_HQ + e];
                        for (int i = 0; i < NX; i++) h += row(R_YC, k + 1)[i] * c[C_HF + i * J2::NH + e];
                    }
                    c[C_H + e] = h;
                }
                for (int i = 0; i < NX; i++)
                    row(R_RC, k + 1)[i] = mode == 1 ? 0.0 : (soc ? row(R_CSOC, k + 1)[i] : c[C_F + i] - row(R_Z, k + 1)[i]);
            }
        }
        double Pn[NX][NX], pn[NX];
        for (int i = 0; i < NX; i++) {
            for (int j = 0; j < NX; j++) Pn[i][j] = i == j ? row(R_DGV, N)[i] : 0.0;
            pn[i] = row(R_GX, N)[i];
            row(R_LIN, N)[i] = pn[i];
            for (int j = 0; j < NX; j++) cst(N)[C_P + i * NX + j] = Pn[i][j];
        }
        for (int k = N - 1; k >= 0; k--) {
            const double *c = cst(k);
            double AB[NX][NZ], Mm[NZ][NZ], mv[NZ], pr[NX], W[NX][NZ];
            for (int i = 0; i < NX; i++) for (int l = 0; l < NZ; l++) AB[i][l] = c[C_J + i * NZ + l];
            for (int i = 0; i < NX; i++) {   // dx_{k+1} = A dx + B du + rc_{k+1}
                double a = pn[i];
                for (int j = 0; j < NX; j++) a += Pn[i][j] * row(R_RC, k + 1)[j];
                pr[i] = a;
            }
            for (int i = 0; i < NX; i++) for (int l = 0; l < NZ; l++) { double a = 0.0; for (int j = 0; j < NX; j++) a += Pn[i][j] * AB[j][l]; W[i][l] = a; }
            for (int a = 0; a < NZ; a++) {
                for (int b = 0; b < NZ; b++) {
                    double s = c[C_H + (a <= b ? J2::hi(a, b) : J2::hi(b, a))];
                    for (int i = 0; i < NX; i++) s += AB[i][a] * W[i][b];
                    Mm[a][b] = s;
                }
                Mm[a][a] += row(R_DGV, k)[a];
                double s = row(R_GX, k)[a];
                for (int i = 0; i < NX; i++) s += AB[i][a] * pr[i];
                mv[a] = s;
            }
            // eliminate the controls: Cholesky of M_uu (pivot <= 0: wrong inertia), K = M_uu^-1 M_ux, kf = M_uu^-1 m_u
            double L[NU][NU], Kk[NU][NX], kf[NU];
            for (int a = 0; a < NU; a++)
                for (int b = 0; b <= a; b++) {
                    double sacc = Mm[NX + a][NX + b];
                    for (int t = 0; t < b; t++) sacc -= L[a][t] * L[b][t];
                    if (a == b) {
                        if (!(sacc > 0.0) || !(sacc < NMPC_INF)) return false;
                        L[a][a] = sqrt(sacc);
                    } else L[a][b] = sacc / L[b][b];
                }
            for (int j = 0; j <= NX; j++) {   // columns of M_ux, then m_u
                double y[NU];
                for (int a = 0; a < NU; a++) {
                    double sacc = j < NX ? Mm[NX + a][j] : mv[NX + a];
                    for (int t = 0; t < a; t++) sacc -= L[a][t] * y[t];
                    y[a] = sacc / L[a][a];
                }
                for (int a = NU - 1; a >= 0; a--) {
                    double sacc = y[a];
                    for (int t = a + 1; t < NU; t++) sacc -= L[t][a] * y[t];
                    y[a] = sacc / L[a][a];
                }
                for (int a = 0; a < NU; a++) { if (j < NX) Kk[a][j] = y[a]; else kf[a] = y[a]; }
            }
            for (int i = 0; i < NX; i++) {   // P = M_xx - M_xu K,  p = m_x - M_xu kf
                for (int j = 0; j < NX; j++) { double sacc = Mm[i][j]; for (int a = 0; a < NU; a++) sacc -= Mm[i][NX + a] * Kk[a][j]; Mm[i][j] = sacc; }
                double sacc = mv[i];
                for (int a = 0; a < NU; a++) sacc -= Mm[i][NX + a] * kf[a];
                mv[i] = sacc;
            }
            double *cw = cst(k);
            for (int i = 0; i < NX; i++) {
                for (int j = 0; j < NX; j++) { Pn[i][j] = 0.5 * (Mm[i][j] + Mm[j][i]); cw[C_P + i * NX + j] = Pn[i][j]; }
                pn[i] = mv[i]; row(R_LIN, k)[i] = pn[i];
            }
            for (int u = 0; u < NU; u++) { for (int j = 0; j < NX; j++) cw[C_K + u * NX + j] = Kk[u][j]; cw[C_KF + u] = kf[u]; }
        }
        return true;
    }

    __device__ __noinline__ void forward(double mu, double tau, int rdz, int rds, int rytc, int rytd, StepInfo &si)
    {
        double ap = 0.0, az = 0.0, gbd = 0.0, tiny = 0.0, dx[NX], du[NU];
        for (int i = 0; i < NX; i++) dx[i] = 0.0;   // the initial state is a parameter
        for (int k = 0; k <= N; k++) {
            const double *c = cst(k);
            for (int i = 0; i < NX; i++) {   // multiplier of row F(z_{k-1}) - X_k: the co-state P_k dx + p_k
                double a = row(R_LIN, k)[i];
                for (int j = 0; j < NX; j++) a += c[C_P + i * NX + j] * dx[j];
                row(rytc, k)[i] = a;
            }
            for (int u = 0; u < NU; u++) {
                double a = 0.0;
                if (k < N) { a = c[C_KF + u]; for (int j = 0; j < NX; j++) a += c[C_K + u * NX + j] * dx[j]; }
                du[u] = -a;
            }
            for (int l = 0; l < LD; l++) {
                const double dzl = l < NX ? dx[l] : (l < NZ ? du[l - NX] : 0.0);
                row(rdz, k)[l] = dzl;
                if (valid(k, l)) {
                    const double z = row(R_Z, k)[l];
                    WS::slack_step_terms(z, dzl, row(R_BL, k)[l], row(R_BU, k)[l], row(R_ZL, k)[l], row(R_ZU, k)[l], mu, ap, az);
                    gbd += row(R_GX, k)[l] * dzl; tiny = fmax(tiny, fabs(dzl) / (1.0 + fabs(z)));
                }
            }
            if (k < N) {
                double dn[NX];
                for (int i = 0; i < NX; i++) {
                    double a = row(R_RC, k + 1)[i];
                    for (int j = 0; j < NX; j++) a += c[C_J + i * NZ + j] * dx[j];
                    for (int u = 0; u < NU; u++) a += c[C_J + i * NZ + NX + u] * du[u];
                    dn[i] = a;
                }
                for (int i = 0; i < NX; i++) dx[i] = dn[i];
            }
        }
        si.ap = ap > tau ? tau / ap : 1.0; si.az = az > tau ? tau / az : 1.0; si.gbd = gbd; si.tiny = tiny;
    }

    __device__ __noinline__ void accept(double alpha, double az, double mu, int rdz, int rds, int rytc, int rytd)
    {
        const double ks = P.o.kappa_sigma, iks = 1.0 / ks;
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < LD; l++) {
                if (valid(k, l)) {
                    const double z = row(R_Z, k)[l], dz = row(rdz, k)[l], lo = row(R_BL, k)[l], hi = row(R_BU, k)[l], zn = z + alpha * dz;
                    if (lo > -NMPC_INF) {
                        const double r = 1.0 / (z - lo), cc = mu / (zn - lo), m = row(R_ZL, k)[l];
                        row(R_ZL, k)[l] = fmax(fmin(m + az * (mu * r - m - m * r * dz), ks * cc), iks * cc);
                    }
                    if (hi < NMPC_INF) {
                        const double r = 1.0 / (hi - z), cc = mu / (hi - zn), m = row(R_ZU, k)[l];
                        row(R_ZU, k)[l] = fmax(fmin(m + az * (mu * r - m + m * r * dz), ks * cc), iks * cc);
                    }
                    row(R_Z, k)[l] = zn;
                }
                if (l < NX && k >= 1) { const double yc = row(R_YC, k)[l]; row(R_YC, k)[l] = yc + alpha * (row(rytc, k)[l] - yc); }
            }
    }
    __device__ __noinline__ void accept_primal(double alpha, int rdz, int rds)
    {
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < NZ; l++) if (valid(k, l)) row(R_Z, k)[l] += alpha * row(rdz, k)[l];
    }
    __device__ __noinline__ void resto_reset(double mu)
    {
        const double ks = P.o.kappa_sigma;
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < LD; l++) {
                if (valid(k, l)) {
                    const double z = row(R_Z, k)[l], lo = row(R_BL, k)[l], hi = row(R_BU, k)[l];
                    if (lo > -NMPC_INF) { const double s2 = z - lo; row(R_ZL, k)[l] = fmax(fmin(row(R_ZL, k)[l], ks * mu / s2), mu / (ks * s2)); }
                    if (hi < NMPC_INF) { const double s2 = hi - z; row(R_ZU, k)[l] = fmax(fmin(row(R_ZU, k)[l], ks * mu / s2), mu / (ks * s2)); }
                }
                row(R_YC, k)[l] = 0.0;
            }
    }
    // restoration candidate: integrate the current controls forward from X_0 (every shooting row becomes zero)
    __device__ __noinline__ void rollout_project()
    {
        const nmpc_opts &o = P.o;
        double zt[LD];
        for (int l = 0; l < LD; l++) zt[l] = row(R_Z, 0)[l];
        for (int k = 0; k <= N; k++) {
            for (int l = 0; l < LD; l++) row(R_DZ, k)[l] = valid(k, l) ? zt[l] - row(R_Z, k)[l] : 0.0;
            if (k < N) {
                model(k, zt);
                for (int i = 0; i < NX; i++) zt[i] = WS::push_in(cst(k)[C_F + i], row(R_BL, k + 1)[i], row(R_BU, k + 1)[i], o.bound_push, o.bound_frac);
                for (int u = 0; u < NU; u++) zt[NX + u] = k + 1 < N ? row(R_Z, k + 1)[NX + u] : 0.0;
            }
        }
    }
    __device__ __noinline__ void soc_begin()
    {
        for (int k = 0; k <= N; k++) for (int l = 0; l < LD; l++) row(R_CSOC, k)[l] = l < NX ? row(R_RC, k)[l] : 0.0;
    }
    __device__ double mult_absmax()
    {
        double m = 0.0;
        for (int k = 1; k <= N; k++) for (int i = 0; i < NX; i++) m = fmax(m, fabs(row(R_YC, k)[i]));
        return m;
    }
    __device__ void mult_zero() { for (int k = 0; k <= N; k++) for (int l = 0; l < LD; l++) row(R_YC, k)[l] = 0.0; }

    __device__ bool filter_ok(double th, double ph) const
    {
        for (int i = 0; i < fn; i++)
            if (!(th < fth[i] || ph < fph[i])) return false;
        return true;
    }
    __device__ __noinline__ void filter_add(double th, double ph)
    {
        int m = 0;
        for (int i = 0; i < fn; i++)
            if (!(fth[i] >= th && fph[i] >= ph)) { fth[m] = fth[i]; fph[m] = fph[i]; m++; }
        if (m == NMPC_FILTER_CAP) { for (int i = 1; i < m; i++) { fth[i - 1] = fth[i]; fph[i - 1] = fph[i]; } m--; }
        fth[m] = th; fph[m] = ph; fn = m + 1;
    }
    __device__ bool trial_ok(double th_t, double ph_t, double theta, double phi, double theta_max, double theta_min, double gbd,
                             double alpha_test, bool ftype) const
    {
        if (!(WS::fin(th_t) && WS::fin(ph_t)) || !WS::cmp_le(th_t, theta_max, theta)) return false;
        bool ok;
        if (ftype && theta <= theta_min) ok = WS::cmp_le(ph_t - phi, 1e-8 * alpha_test * gbd, phi);
        else ok = WS::cmp_le(th_t, (1.0 - 1e-5) * theta, theta) || WS::cmp_le(ph_t - phi, -1e-8 * theta, phi);
        return ok && filter_ok(th_t, ph_t);
    }

    __device__ __noinline__ void write_outputs(int st, int iter, double E0, double pinf, double dinf, double c0, double mu)
    {
        const long long n = (long long)NZ * N + NX, mg = (long long)NX * N;
        double *x = P.x + inst * n;
        double fo = 0.0;
        for (int k = 0; k <= N; k++)
            for (int l = 0; l < (k < N ? NZ : NX); l++) {
                x[k * NZ + l] = row(R_Z, k)[l];
                if (P.lam_x) P.lam_x[inst * n + k * NZ + l] = valid(k, l) ? (row(R_ZU, k)[l] - row(R_ZL, k)[l]) / df : 0.0;
            }
        for (int k = 0; k < N; k++) {
            model(k, row(R_Z, k));
            fo += cst(k)[C_Q];
            for (int i = 0; i < NX; i++) {
                if (P.g) P.g[inst * mg + k * NX + i] = cst(k)[C_F + i] - row(R_Z, k + 1)[i];
                if (P.lam_g) P.lam_g[inst * mg + k * NX + i] = row(R_YC, k + 1)[i] / df;
            }
        }
        if (P.f) P.f[inst] = fo;
        if (P.status) P.status[inst] = st;
        if (P.iters) P.iters[inst] = iter;
        if (P.stats) {
            double *sp = P.stats + (long long)inst * NMPC_NSTATS;
            sp[NMPC_ST_KKT_ERR] = E0; sp[NMPC_ST_PRIMAL_INF] = pinf; sp[NMPC_ST_DUAL_INF] = dinf; sp[NMPC_ST_COMPL] = c0;
            sp[NMPC_ST_MU] = mu; sp[NMPC_ST_N_REG] = n_reg; sp[NMPC_ST_N_RESTO] = n_resto; sp[NMPC_ST_N_SOC] = n_soc;
            sp[NMPC_ST_N_FACTOR] = n_fact; sp[NMPC_ST_N_LS] = n_ls;
        }
    }
    __device__ void run() { ipm_run(*this); }
};

// one thread per instance
template <class Model>
__global__ void __launch_bounds__(64) solve_kernel_small_ocp(const NmpcSolveParams P)
{
    for (int inst = blockIdx.x * blockDim.x + threadIdx.x; inst < P.B; inst += gridDim.x * blockDim.x) {
        ThreadSolver<Model> s(P, P.ws + (long long)inst * P.ws_stride);
        s.setup(inst);
        s.run();
    }
}

// The interior-point iteration (K2), shared by the warp-per-instance solver (solver_body.cuh, 1..10 robots) and the
// CTA-per-instance dense-block solver (block_solver.cuh, larger swarms).
//
// IPOPT's monotone-mu primal-dual filter line-search method (Waechter & Biegler 2006 with IPOPT's default options,
// SURVEY.md App. B) behind the reference's  sol = solver(x0=, p=, lbx=, ubx=, lbg=, ubg=)
// (centralized_six_robots_implementation.py:432): least-squares multiplier initialisation, scaled optimality error
// and termination tests, inertia correction, fraction-to-boundary, filter with switching / Armijo conditions,
// second-order corrections, kappa_sigma multiplier reset, bounded restoration substitute.
//
// The solver type S supplies the passes (init_point, factor_m, forward, eval, accept, ...) and the team-wide
// synchronisation; every thread of a team executes this function with identical control flow.
#pragma once

template <class S>
NMPC_DEV void ipm_run(S &s)
{
    const nmpc_opts &o = s.P.o;
    if (s.bounds_rejected()) return;
    s.init_point();
    // least-squares equality multipliers (W = 0, Sigma = I); discarded when too large
    {
        bool ok = s.factor_m(1, 0.0, 0.0, false);
        typename S::StepInfo si;
        if (ok) s.forward(0.0, 0.99, S::R_DZ, S::R_DS, S::R_YC, S::R_YD, si);
        const double ymax = s.mult_absmax();
        if (!ok || !(ymax <= o.constr_mult_init_max)) s.mult_zero();
        S::tsync();
    }
    double mu = o.mu_init, tau = fmax(o.tau_min, 1.0 - mu);
    double theta_max = -1.0, theta_min = -1.0, delta_last = 0.0, f_prev = 0.0;
    int iter = 0, st = NMPC_MAX_ITER, n_acc = 0;
    bool tiny_prev = false, at_top = false;   // at_top: between the two convoy barriers of an iteration
    double E0 = 0.0;
    const double mu_floor = fmin(o.tol, o.compl_inf_tol) / (o.barrier_tol_factor + 1.0);
    typename S::EvalOut E;
    E.dinf = E.c0 = E.cmu = E.ysum = E.zsum = 0.0;
    for (;;) {
        s.iter_sync(); at_top = true;
        s.eval(true, mu, 0.0, 0, 0, false, false, 0.0, E);
        const double smax = 100.0;
        const double sd = fmax(smax, (E.ysum + E.zsum) / fmax(1.0, s.ny_nzb)) / smax;
        const double sc = fmax(smax, E.zsum / fmax(1.0, s.nzb_cnt)) / smax;
        for (int pass = 0;; pass++) {
            E0 = fmax(fmax(E.dinf / sd, E.pinf), E.c0 / sc);
            const double Emu = fmax(fmax(E.dinf / sd, E.pinf), E.cmu / sc);
            if (pass == 0) {
                if (E0 <= o.tol && E.dinf / s.df <= o.dual_inf_tol && E.viol <= o.constr_viol_tol && E.c0 / s.df <= o.compl_inf_tol) { st = NMPC_SOLVED; goto finished; }
                bool acc = E0 <= o.acceptable_tol && E.dinf / s.df <= 1e10 && E.viol <= 1e-2 && E.c0 / s.df <= 1e-2 &&
                           (iter == 0 || fabs(E.f - f_prev) / fmax(1.0, fabs(E.f)) <= o.acceptable_obj_change_tol);
                n_acc = acc ? n_acc + 1 : 0;
                if (n_acc >= o.acceptable_iter) { st = NMPC_ACCEPTABLE; goto finished; }
                if (iter >= o.max_iter) { st = NMPC_MAX_ITER; goto finished; }
            }
            if (!(Emu <= o.barrier_tol_factor * mu) && !(tiny_prev && pass == 0)) break;
            const double nm = fmax(fmin(o.kappa_mu * mu, pow(mu, o.theta_mu)), mu_floor);
            if (nm >= mu) break;
            mu = nm; tau = fmax(o.tau_min, 1.0 - mu); s.fn = 0; tiny_prev = false;
            s.eval(true, mu, 0.0, 0, 0, false, false, 0.0, E);
        }
        f_prev = E.f;
        const double theta = E.theta;
        const double phi = s.df * E.f - mu * E.slog + o.kappa_d * mu * E.sdamp;
        if (theta_max < 0.0) { theta_max = 1e4 * fmax(1.0, theta); theta_min = 1e-4 * fmax(1.0, theta); }
        double *tr = (s.P.trace && iter < s.P.max_trace) ? s.P.trace + ((long long)s.inst * s.P.max_trace + iter) * NMPC_NTRACE : nullptr;
        if (tr && s.is_lead()) { tr[0] = mu; tr[1] = E0; tr[2] = theta; tr[3] = E.f; tr[4] = tr[5] = tr[6] = tr[7] = 0.0; }
        // ---- search direction with inertia correction ----
        double delta = 0.0;
        bool need_resto = false;
        for (;;) {
            if (s.factor_m(0, mu, delta, false)) break;
            if (delta == 0.0) delta = delta_last == 0.0 ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else delta *= (delta_last == 0.0 || 1e5 * delta_last < delta) ? 100.0 : 8.0;
            if (delta > 1e20) { need_resto = true; break; }
        }
        if (delta > 0.0 && !need_resto) { delta_last = delta; s.n_reg++; }
        s.mid_sync(); at_top = false;
        double alpha = 0.0, alpha_z = 0.0;
        int ls_count = 0;
        if (!need_resto) {
            typename S::StepInfo si;
            s.forward(mu, tau, S::R_DZ, S::R_DS, S::R_YTC, S::R_YTD, si);
            const double gbd = si.gbd;
            const bool tiny = si.tiny < 10.0 * 2.220446049250313e-16 && theta < 1e-4;
            double amin = 1e-5;
            if (gbd < 0.0) {
                amin = fmin(1e-5, 1e-8 * theta / (-gbd));
                if (theta <= theta_min) amin = fmin(amin, pow(theta, 1.1) / pow(-gbd, 2.3));
            }
            amin *= 0.05;
            bool accepted = false, armijo_step = false, use_soc = false;
            alpha = si.ap; alpha_z = si.az;
            if (tiny) { accepted = true; tiny_prev = true; }
            while (!accepted) {
                ls_count++;
                typename S::EvalOut Et;
                s.eval(false, mu, alpha, S::R_DZ, S::R_DS, true, false, 0.0, Et);
                const double th_t = Et.theta, ph_t = s.df * Et.f - mu * Et.slog + o.kappa_d * mu * Et.sdamp;
                const bool ftype = gbd < 0.0 && alpha * pow(-gbd, 2.3) > pow(theta, 1.1);
                if (s.trial_ok(th_t, ph_t, theta, phi, theta_max, theta_min, gbd, alpha, ftype)) {
                    accepted = true; armijo_step = ftype && theta <= theta_min; break;
                }
                if (ls_count == 1 && th_t >= theta && o.max_soc > 0) {  // second-order correction
                    double th_old = 0.0, th_tr = th_t, a_soc = alpha;
                    int cnt = 0;
                    s.soc_begin();
                    int rz = S::R_DZ, rs = S::R_DS;  // direction whose trial point feeds the next correction
                    while (cnt < o.max_soc && !accepted && (cnt == 0 || th_tr <= 0.99 * th_old)) {
                        th_old = th_tr;
                        typename S::EvalOut Ea;  // c_soc := a_soc c_soc + c(trial)
                        s.eval(false, mu, a_soc, rz, rs, true, true, a_soc, Ea);
                        s.n_soc++;
                        if (!s.factor_m(0, mu, delta, true)) break;
                        typename S::StepInfo s2;
                        s.forward(mu, tau, S::R_DZ2, S::R_DS2, S::R_YTC2, S::R_YTD2, s2);
                        a_soc = s2.ap; rz = S::R_DZ2; rs = S::R_DS2;
                        typename S::EvalOut E2;
                        s.eval(false, mu, a_soc, S::R_DZ2, S::R_DS2, true, false, 0.0, E2);
                        const double th2 = E2.theta, ph2 = s.df * E2.f - mu * E2.slog + o.kappa_d * mu * E2.sdamp;
                        if (s.trial_ok(th2, ph2, theta, phi, theta_max, theta_min, gbd, alpha, ftype)) {
                            accepted = true; armijo_step = ftype && theta <= theta_min; use_soc = true;
                            alpha = a_soc; alpha_z = s2.az;
                        } else { cnt++; th_tr = th2; }
                    }
                    if (accepted) break;
                }
                alpha *= 0.5;
                if (alpha < amin) break;
            }
            if (!accepted) need_resto = true;
            else {
                if (!tiny && !armijo_step) s.filter_add((1.0 - 1e-5) * theta, phi - 1e-8 * theta);
                if (!tiny) tiny_prev = false;
                if (use_soc) s.accept(alpha, alpha_z, mu, S::R_DZ2, S::R_DS2, S::R_YTC2, S::R_YTD2);
                else s.accept(alpha, alpha_z, mu, S::R_DZ, S::R_DS, S::R_YTC, S::R_YTD);
            }
        }
        if (need_resto) {
            // bounded substitute for IPOPT's restoration phase (see oracle/nmpc_oracle.c)
            s.n_resto++;
            s.filter_add((1.0 - 1e-5) * theta, phi - 1e-8 * theta);
            const double thR = theta;
            bool ok = false;
            {   // first candidate: rollout projection onto the dynamics (see oracle/nmpc_oracle.c)
                s.rollout_project();
                typename S::EvalOut Ep;
                s.eval(false, mu, 1.0, S::R_DZ, S::R_DS, true, false, 0.0, Ep);
                if (Ep.theta < thR) s.accept_primal(1.0, S::R_DZ, S::R_DS);
            }
            for (int r_it = 0; r_it < o.max_resto_iter; r_it++) {
                typename S::EvalOut Er;
                s.eval(false, mu, 0.0, 0, 0, false, false, 0.0, Er);
                const double th = Er.theta;
                if ((th <= 0.9 * thR || th <= 1e-9) &&
                    s.filter_ok(th, s.df * Er.f - mu * Er.slog + o.kappa_d * mu * Er.sdamp)) { ok = true; break; }
                if (!s.factor_m(2, mu, 0.0, false)) break;
                typename S::StepInfo sr;
                s.forward(mu, tau, S::R_DZ, S::R_DS, S::R_YTC, S::R_YTD, sr);
                double a = sr.ap, th_t = th;
                bool got = false;
                while (a > 1e-12) {
                    typename S::EvalOut Et;
                    s.eval(false, mu, a, S::R_DZ, S::R_DS, true, false, 0.0, Et);
                    th_t = Et.theta;
                    if (th_t <= (1.0 - 1e-4 * a) * th) { got = true; break; }
                    a *= 0.5;
                }
                if (!got) break;
                s.accept_primal(a, S::R_DZ, S::R_DS);
                if (th - th_t < 1e-14 * fmax(1.0, th)) break;
            }
            if (!ok) { st = NMPC_INFEASIBLE; iter++; goto finished; }
            s.resto_reset(mu);
            alpha = 0.0; alpha_z = 0.0;
        }
        s.n_ls += ls_count;
        if (tr && s.is_lead()) { tr[4] = alpha; tr[5] = alpha_z; tr[6] = delta; tr[7] = ls_count; }
        iter++;
    }
finished:
    if (at_top) s.mid_sync();   // two barriers per iteration on every path: the group stays in phase
    s.write_outputs(st, iter, E0, E.pinf, E.dinf, E.c0, mu);
}

"""Generates tests/golden/polish6_scipy.npz: an independent-solver pin of the BENCHMARK configuration
(6 robots, N = 20, T = 0.3, dmin = 0.3; sixth_scenario.py:127-135 with the horizon set to 20).

CasADi/IPOPT cannot be installed here (profiles/probe_casadi_r2.json), so the point the restated IPOPT (oracle/) and
the CUDA solver stop at is checked against two solvers that share no code with either of them:

  * SciPy SLSQP (Kraft's SQP, dense, analytic derivatives from oracle/nlp_numpy.py), and
  * SciPy trust-constr in its equality-constrained SQP mode (Byrd-Omojokun, exact Hessian of the Lagrangian) on the active
    face of x*, followed by the multiplier-sign and second-order-sufficiency checks that make the face result a certificate
    for the full problem (see face_sqp),

each STARTED FROM the oracle's x* (parity tier (a): "polish test").  If x* is the strict local minimiser
IPOPT would return, both must stay there; the fixture stores where they ended, so the GPU parity test
(tests/test_gpu_parity.py::test_benchmark_config_matches_independent_polish) can require
max|u - u_polish| <= 1e-4 and |f - f_polish| / f <= 1e-6 (north_star's tolerances) without SciPy in the loop.

A second, informational part solves a few instances from the reference's COLD START with SLSQP to report the
basin-agreement rate (the NLP is multi-modal).

Run from the repo root (about 5 minutes on 8 cores):   python tests/golden/make_polish_golden.py
"""
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):      # one BLAS thread per worker process
    os.environ[_v] = "1"
from multiprocessing import Pool  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.nlp_numpy import UnicycleNLP, synthetic_instances  # noqa: E402
from oracle.oracle_lib import Oracle  # noqa: E402

NR, NH, T, DMIN, VMAX, WMAX = 6, 20, 0.3, 0.3, 0.22, 2.84
N_SYNTH, N_COLD = 32, 8


def hexagon_desym():
    """C-6 hexagon swap (sixth_scenario.py:291-292,308-310), de-symmetrised as in bench.py's latency fixture."""
    s3 = np.sqrt(3) / 2
    st = np.array([[s3, 0.5, -2.618], [0, 1, -1.571], [-s3, 0.5, -0.524], [-s3, -0.5, 0.524], [0, -1, 1.571], [s3, -0.5, 2.618]])
    st = st + 0.02 * np.sin(1.0 + 2.0 * np.arange(18)).reshape(6, 3)
    goal = -st.copy(); goal[:, 2] = st[:, 2]
    return np.concatenate([st.ravel(), goal.ravel()])


def problem():
    nlp = UnicycleNLP(NR, NH, T)
    lbx, ubx, lbg, ubg = nlp.bounds(DMIN, VMAX, WMAX)
    eq = np.nonzero(lbg == ubg)[0]
    blk = nlp.blk
    ineq = np.array([r for r in range(nlp.mg) if lbg[r] != ubg[r] and r >= blk])      # real distance rows (block 0 holds constants)
    return nlp, lbx, ubx, lbg, ubg, eq, ineq


def slsqp(w0, p, maxiter, ftol):
    from scipy.optimize import Bounds, minimize
    nlp, lbx, ubx, lbg, ubg, eq, ineq = problem()
    cons = [dict(type="eq", fun=lambda w: nlp.g(w, p)[eq], jac=lambda w: nlp.jac_g(w, p)[eq]),
            dict(type="ineq", fun=lambda w: nlp.g(w, p)[ineq] - lbg[ineq], jac=lambda w: nlp.jac_g(w, p)[ineq])]
    res = minimize(lambda w: nlp.f(w, p), w0, jac=lambda w: nlp.grad_f(w, p), method="SLSQP", bounds=Bounds(lbx, ubx),
                   constraints=cons, options=dict(maxiter=maxiter, ftol=ftol))
    return res.x, float(res.fun), int(res.nit), int(res.status)


def face_sqp(x_star, p, tol_active=1e-6):
    """trust-constr on the ACTIVE FACE of x*: variables at a bound are fixed, active distance rows become equalities, so SciPy
    runs its Byrd-Omojokun equality-constrained SQP (no barrier).  The full trust-constr (its own interior point) is not a
    polisher: it re-initialises its slacks to 1, i.e. cold-starts from x*, and on this multi-modal NLP ends in another basin
    (hexagon: a KKT point 1.3 % away in f).  Together with the multiplier-sign and second-order checks below the face SQP
    certifies a strict local minimiser of the full problem.  Returns (x, f, nit, status, cert) with
    cert = (min multiplier slack of the active rows, min of the active-bound multipliers, smallest eigenvalue of the
    reduced Hessian, number of active bounds, number of active distance rows)."""
    from scipy.linalg import null_space
    from scipy.optimize import NonlinearConstraint, minimize
    from scipy.sparse import csr_matrix
    nlp, lbx, ubx, lbg, ubg, eq, ineq = problem()
    g = nlp.g(x_star, p)
    act_lo, act_hi = np.nonzero(x_star - lbx < tol_active)[0], np.nonzero(ubx - x_star < tol_active)[0]
    act_g = np.array([i for i in ineq if g[i] - lbg[i] < tol_active], dtype=int)
    rows = np.concatenate([eq, act_g]).astype(int)
    fixed, fixval = np.concatenate([act_lo, act_hi]), np.concatenate([lbx[act_lo], ubx[act_hi]])
    free = np.setdiff1d(np.arange(nlp.n), fixed)

    def full(y):
        w = np.empty(nlp.n); w[free] = y; w[fixed] = fixval
        return w

    def chess(y, v):
        lam = np.zeros(nlp.mg); lam[rows] = v
        return nlp.hess_lag(full(y), p, lam, sigma=0.0)[np.ix_(free, free)]

    H0 = nlp.hess_lag(x_star, p, np.zeros(nlp.mg), sigma=1.0)
    con = NonlinearConstraint(lambda y: nlp.g(full(y), p)[rows] - lbg[rows], 0.0, 0.0,
                              jac=lambda y: csr_matrix(nlp.jac_g(full(y), p)[np.ix_(rows, free)]), hess=chess)
    y, nit = x_star[free], 0
    for rnd in range(4):   # the SQP stops on its trust radius (xtol) in the flat omega directions: restart it until it stays put
        res = minimize(lambda y: nlp.f(full(y), p), y, jac=lambda y: nlp.grad_f(full(y), p)[free], hess=lambda y: H0[np.ix_(free, free)],
                       method="trust-constr", constraints=[con], options=dict(maxiter=300, gtol=1e-11, xtol=1e-15, initial_tr_radius=0.01 if rnd == 0 else 0.1))
        moved = np.abs(res.x - y).max()
        y, nit = res.x, nit + int(res.nit)
        if moved <= 1e-8:
            break
    x = full(res.x)
    # certificate at the polished point: L = f + v'c with c >= 0 active  =>  v <= 0;  fixed variables: dL/dx_i >= 0 at a lower
    # bound, <= 0 at an upper bound;  second-order sufficiency: Z'(H_f + sum v_i H_ci) Z > 0 on the null space of the active Jacobian
    v = res.v[0]
    lam = np.zeros(nlp.mg); lam[rows] = v
    J = nlp.jac_g(x, p)
    gl = nlp.grad_f(x, p) + J.T @ lam
    sign_rows = float((-v[len(eq):]).min()) if len(act_g) else np.inf
    sign_bnd = float(min([gl[i] for i in act_lo] + [-gl[i] for i in act_hi] + [np.inf]))
    Z = null_space(J[np.ix_(rows, free)])
    HL = nlp.hess_lag(x, p, lam, sigma=1.0)[np.ix_(free, free)]
    ev = float(np.linalg.eigvalsh(Z.T @ HL @ Z).min()) if Z.shape[1] else np.inf
    return x, float(res.fun), nit, int(res.status), (sign_rows, sign_bnd, ev, len(fixed), len(act_g))


def polish_one(args):
    x_star, p = args
    t0 = time.time()
    xs, fs, its, sts = slsqp(x_star, p, 100, 1e-15)
    xt, ft, itt, stt, cert = face_sqp(x_star, p)
    return xs, fs, its, sts, xt, ft, itt, stt, time.time() - t0, cert


def cold_one(args):
    w0, p = args
    return slsqp(w0, p, 250, 1e-13)


def main():
    orc = Oracle(NR, NH, T)
    nlp, lbx, ubx, lbg, ubg, eq, ineq = problem()
    P = np.concatenate([hexagon_desym()[None], synthetic_instances(N_SYNTH, NR, 20261018)])
    w0 = np.stack([orc.cold_start(q[:3 * NR]) for q in P])
    ref = orc.solve_batch(w0, P, lbx, ubx, lbg, ubg)
    assert np.all(ref["status"] == 0), ref["status"]
    print("oracle: iters", ref["iters"])
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        cold_async = pool.map_async(cold_one, [(w0[b], P[b]) for b in range(1, 1 + N_COLD)], chunksize=1)
        pol = pool.map(polish_one, [(ref["x"][b], P[b]) for b in range(len(P))], chunksize=1)
        print("polish done", flush=True)
        cold = cold_async.get()
    nu0 = 3 * NR * (NH + 1)
    x_s = np.stack([r[0] for r in pol]); f_s = np.array([r[1] for r in pol])
    x_t = np.stack([r[4] for r in pol]); f_t = np.array([r[5] for r in pol])
    for b, r in enumerate(pol):
        du_s = np.abs(r[0] - ref["x"][b])[nu0:].max(); du_t = np.abs(r[4] - ref["x"][b])[nu0:].max()
        print("inst %2d: f*=%.9f  SLSQP nit %3d st %d du %.2e df %.2e | face-SQP nit %3d st %d du %.2e df %.2e  mult>=%.1e bnd>=%.1e redH>=%.2e act %d+%d (%.0f s)" % (
            b, ref["f"][b], r[2], r[3], du_s, abs(r[1] - ref["f"][b]) / ref["f"][b], r[6], r[7], du_t, abs(r[5] - ref["f"][b]) / ref["f"][b],
            r[9][0], r[9][1], r[9][2], r[9][3], r[9][4], r[8]))
    x_c = np.stack([r[0] for r in cold]); f_c = np.array([r[1] for r in cold])
    same = [bool(np.abs(x_c[i] - ref["x"][1 + i])[nu0:].max() <= 1e-3) for i in range(N_COLD)]
    for i, r in enumerate(cold):
        print("cold %d: SLSQP f=%.6f nit %d st %d | oracle f=%.6f  same basin: %s" % (i, r[1], r[2], r[3], ref["f"][1 + i], same[i]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "polish6_scipy.npz"), P=P, x_oracle=ref["x"], f_oracle=ref["f"],
                        iters_oracle=ref["iters"], x_slsqp=x_s, f_slsqp=f_s, x_face=x_t, f_face=f_t,
                        cert=np.array([r[9] for r in pol], dtype=float),
                        x_cold_slsqp=x_c, f_cold_slsqp=f_c, cold_same_basin=np.array(same))


if __name__ == "__main__":
    main()

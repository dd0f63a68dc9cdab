"""Generates tests/golden/polish_scenarios_scipy.npz: the independent SLSQP pin of tests/golden/make_polish_golden.py for the FIRST MPC
step of the reference's own scenarios at their own horizons (C-1 ... C-6, C-2r, C-6r; de-symmetrised starts as in
tests/test_closed_loop.py).  SciPy SLSQP (analytic derivatives of oracle/nlp_numpy.py) is started from the restated IPOPT's x*; the
fixture stores where it ends.  Run from the repo root (a few minutes):   python tests/golden/make_scenario_polish_golden.py"""
import importlib.util
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = "2"
from multiprocessing import Pool  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.nlp_numpy import UnicycleNLP  # noqa: E402
from oracle.oracle_lib import Oracle  # noqa: E402

SIDS = ["C-1", "C-2", "C-2r", "C-3", "C-4", "C-5", "C-6", "C-6r"]


def scenarios():
    """mpc_loop.SCENARIOS without importing the package's CUDA binding (host-only module loaded by path)."""
    pkg = [d for d in os.listdir(ROOT) if d.endswith("_b200")][0]
    spec = importlib.util.spec_from_file_location("mpc_loop_only", os.path.join(ROOT, pkg, "mpc_loop.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m.SCENARIOS


def one(sid):
    from scipy.optimize import Bounds, minimize
    Nr, T, N, dmin, vmax, wmax, start, goal, tol = scenarios()[sid]
    start = np.asarray(start, float) + 0.02 * np.sin(1.0 + 2.0 * np.arange(3 * Nr))
    p = np.concatenate([start, np.asarray(goal, float)])
    nlp, orc = UnicycleNLP(Nr, N, T), Oracle(Nr, N, T)
    lbx, ubx, lbg, ubg = nlp.bounds(dmin, vmax, wmax)
    r = orc.solve(orc.cold_start(p[:3 * Nr]), p, lbx, ubx, lbg, ubg)
    assert r["status"] == 0, (sid, r["status"])
    eq = np.nonzero(lbg == ubg)[0]
    ineq = np.array([i for i in range(nlp.mg) if lbg[i] != ubg[i] and i >= nlp.blk], dtype=int)
    cons = [dict(type="eq", fun=lambda w: nlp.g(w, p)[eq], jac=lambda w: nlp.jac_g(w, p)[eq])]
    if len(ineq):
        cons.append(dict(type="ineq", fun=lambda w: nlp.g(w, p)[ineq] - lbg[ineq], jac=lambda w: nlp.jac_g(w, p)[ineq]))
    t0 = time.time()
    res = minimize(lambda w: nlp.f(w, p), r["x"], jac=lambda w: nlp.grad_f(w, p), method="SLSQP", bounds=Bounds(lbx, ubx),
                   constraints=cons, options=dict(maxiter=60, ftol=1e-15))
    nX = 3 * Nr * (N + 1)
    du, df = np.abs(res.x - r["x"])[nX:].max(), abs(res.fun - r["f"]) / max(1.0, abs(r["f"]))
    print("%-5s Nr %d N %3d n %4d: oracle iters %3d f* %.9f | SLSQP nit %2d st %d du %.2e df %.2e (%.0f s)" % (
        sid, Nr, N, nlp.n, r["iters"], r["f"], res.nit, res.status, du, df, time.time() - t0), flush=True)
    return sid, p, r["x"], r["f"], res.x, float(res.fun)


def main():
    with Pool(4) as pool:
        out = pool.map(one, SIDS, chunksize=1)
    d = {}
    for sid, p, xo, fo, xs, fs in out:
        k = sid.replace("-", "")
        d["p_" + k], d["x_oracle_" + k], d["f_oracle_" + k], d["x_slsqp_" + k], d["f_slsqp_" + k] = p, xo, fo, xs, fs
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "polish_scenarios_scipy.npz"), sids=np.array(SIDS), **d)


if __name__ == "__main__":
    main()

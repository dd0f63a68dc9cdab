"""SURVEY.md 8f-4 / BASELINE.json configs[0]: the Van der Pol multiple-shooting demo (mpc_pose_control_casadi.py:22-114) on the
thread-per-instance small-OCP solver, against an independent SLSQP solve of a NumPy restatement."""
import numpy as np
import pytest

from oracle.vdp_nlp import VdpNLP


def test_restated_rk4_interval_against_a_tight_ode_solve():
    """The demo evaluates F(x0=[0.2, 0.3], p=0.4) as its own sanity print (:62-64); the reference records no output, so the
    restated RK4 interval (4 steps of 0.125) is checked against SciPy's adaptive integrator at tight tolerances instead."""
    from scipy.integrate import solve_ivp
    nlp = VdpNLP()
    xf, qf = nlp.F([0.2, 0.3], 0.4)
    rhs = lambda t, y: [(1 - y[1] ** 2) * y[0] - y[1] + 0.4, y[0], y[0] ** 2 + y[1] ** 2 + 0.4 ** 2]
    ref = solve_ivp(rhs, [0.0, nlp.T / nlp.N], [0.2, 0.3, 0.0], rtol=1e-12, atol=1e-14).y[:, -1]
    np.testing.assert_allclose([xf[0], xf[1], qf], ref, atol=2e-5)      # RK4 global error at h = 0.125
    # gradient of the whole NLP objective and constraints by finite differences is used by the SLSQP reference: shapes
    w0 = nlp.demo_arrays()[0]
    f, g = nlp.fg(w0)
    assert g.shape == (40,) and np.isfinite(f)


@pytest.mark.gpu
def test_gpu_van_der_pol_demo_matches_slsqp(pkg):
    import torch
    assert torch.cuda.is_available()
    ocp = pkg.SmallOcp("van_der_pol", N=20, T=10.0, rk_steps=4)
    nlp = VdpNLP()
    assert (ocp.n, ocp.mg) == (nlp.n, nlp.mg) == (62, 40)
    w0, lbw, ubw, lbg, ubg = ocp.demo_arrays()
    for a, b in zip((w0, lbw, ubw, lbg, ubg), nlp.demo_arrays()):
        np.testing.assert_array_equal(a, b)
    out = ocp.solve_host(w0, lbw, ubw, lbg, ubg)
    assert out["status"][0] == 0, (out["status"], out["iters"], out["stats"])
    assert out["stats"][0, 0] <= 1e-8
    w = out["x"][0]
    f_np, g_np = nlp.fg(w)
    np.testing.assert_allclose(out["g"][0], g_np, atol=1e-12)
    assert abs(out["f"][0] - f_np) <= 1e-12 * max(1.0, abs(f_np))
    assert np.abs(g_np).max() <= 1e-8                                   # shooting rows closed
    assert w[0] == 0.0 and w[1] == 1.0                                   # fixed initial state (:79-80)
    assert (w[2::3] >= -1 - 1e-8).all() and (w[2::3] <= 1 + 1e-8).all() and (w[0::3] >= -0.25 - 1e-8).all()
    ref = nlp.solve_slsqp(w0, lbw, ubw)
    assert abs(ref.fun - out["f"][0]) <= 1e-6 * max(1.0, abs(ref.fun)), (ref.fun, out["f"][0])
    du = np.abs(ref.x - w)[2::3].max()
    if du > 1e-4:       # SLSQP without analytic derivatives: polish from the product's point, it must stay there
        ref = nlp.solve_slsqp(w, lbw, ubw)
        assert ref.fun >= out["f"][0] - 1e-8
        du = np.abs(ref.x - w)[2::3].max()
    assert du <= 1e-4, du
    # the demo's own call surface (:109-114): parameter-free solver(x0=, lbx=, ubx=, lbg=, ubg=), w_opt[0::3] etc.
    solver = pkg.nlpsol("solver", "ipopt", {"family": "van_der_pol", "N": 20, "T": 10.0, "rk_steps": 4})
    sol = solver(x0=list(w0), lbx=list(lbw), ubx=list(ubw), lbg=list(lbg), ubg=list(ubg))
    w_opt = sol["x"].full().flatten()
    np.testing.assert_array_equal(w_opt, w)
    # a batch of perturbed initial guesses converges to the same optimum (the problem has one minimiser in this region)
    rng = np.random.default_rng(0)
    W0 = np.tile(w0, (64, 1)); W0[:, 2:] += 0.1 * rng.normal(size=(64, nlp.n - 2))
    outb = ocp.solve_host(W0, lbw, ubw, lbg, ubg)
    assert (outb["status"] == 0).all()
    assert np.abs(outb["x"] - w[None]).max() <= 1e-5

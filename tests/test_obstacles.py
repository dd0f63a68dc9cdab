"""Static-obstacle family (first_/third_scenario_mpc_obstacle_avoidance.py): CPU checks of the restated
NLP, GPU parity of the product against an independent SLSQP solve of that restatement."""
import numpy as np
import pytest

from oracle.obstacle_nlp import ObstacleNLP

# third_scenario_mpc_obstacle_avoidance.py:97-119 (centre x, y, r_obs), rob_dim = 0.15 as in the first scenario (:60)
THIRD = [(-0.6, 3.3, 0.2), (0.6, 3.3, 0.125), (0.0, 2.3, 0.15), (1.0, 2.3, 0.15), (-0.6, 1.3, 0.2), (0.6, 1.3, 0.175)]
ROB_DIM = 0.15


def _obs(rows):
    return np.array([[x, y, r + ROB_DIM] for x, y, r in rows])


def test_restated_obstacle_nlp_derivatives():
    nlp = ObstacleNLP(12, 0.2, _obs(THIRD))
    assert (nlp.n, nlp.mg) == (3 * 13 + 2 * 12, 3 + 12 * 9)          # third scenario row count per stage: 3 + 6 (:175)
    rng = np.random.default_rng(1)
    w, p = rng.normal(size=nlp.n), rng.normal(size=6)
    eps = 1e-6
    Jfd = np.stack([(nlp.g(w + eps * e, p) - nlp.g(w - eps * e, p)) / (2 * eps) for e in np.eye(nlp.n)], axis=1)
    np.testing.assert_allclose(nlp.jac_g(w, p), Jfd, atol=1e-8)
    gfd = np.array([(nlp.f(w + eps * e, p) - nlp.f(w - eps * e, p)) / (2 * eps) for e in np.eye(nlp.n)])
    np.testing.assert_allclose(nlp.grad_f(w, p), gfd, atol=1e-6)
    # row layout: block 0 = X_0 - x0bar only; obstacle rows follow the three defect rows of every stage
    X, U = nlp.split(w)
    g = nlp.g(w, p)
    np.testing.assert_allclose(g[:3], X[0] - p[:3])
    np.testing.assert_allclose(g[3 + 3:3 + 9], np.hypot(X[0, 0] - nlp.obs[:, 0], X[0, 1] - nlp.obs[:, 1]) - nlp.obs[:, 2])


def test_slsqp_reference_solution_avoids_the_obstacle():
    nlp = ObstacleNLP(15, 0.3, [[0.45, 0.5, 0.3]])
    p = np.array([0.0, 0.0, 0.6, 1.2, 1.3, 0.0])
    lbx, ubx, lbg, ubg = nlp.bounds(0.05, 0.2, np.pi / 4)
    res = nlp.solve_slsqp(nlp.cold_start(p[:3]), p, lbx, ubx, lbg, ubg)
    assert res.success
    g = nlp.g(res.x, p)
    assert np.abs(g[lbg == ubg]).max() < 1e-8 and (g[lbg != ubg] >= 0.05 - 1e-8).all()
    assert (g[lbg != ubg] < 0.05 + 1e-6).any(), "the obstacle should be active on the way to the goal"


def _cases():
    one = dict(N=15, T=0.3, obs=np.array([[0.45, 0.5, 0.3]]), margin=0.05,
               P=np.array([[0.0, 0.0, 0.6, 1.2, 1.3, 0.0], [0.1, -0.1, 0.9, 1.0, 1.4, 0.3]]))
    six = dict(N=20, T=0.3, obs=_obs(THIRD), margin=0.1,
               P=np.array([[0.0, 0.6, 1.57, 0.1, 3.9, 1.57], [0.3, 0.7, 1.2, -0.2, 3.6, 1.57]]))
    return {"one": one, "six": six}


@pytest.mark.parametrize("case", ["one", "six"])
def test_c_oracle_covers_the_obstacle_family(case):
    """oracle/nmpc_oracle.c with obstacle rows (restated IPOPT, same algorithm as for the robot pairs): its solution is a KKT point
    of the NumPy restatement, in the scripts' row layout, and agrees with the independent SLSQP solve."""
    from oracle.oracle_lib import Oracle
    c = _cases()[case]
    nlp, orc = ObstacleNLP(c["N"], c["T"], c["obs"]), Oracle(1, c["N"], c["T"], obstacles=c["obs"])
    assert (orc.n, orc.mg) == (nlp.n, nlp.mg)
    lbx, ubx, lbg, ubg = nlp.bounds(c["margin"], 0.2, np.pi / 4)
    for p in c["P"]:
        r = orc.solve(nlp.cold_start(p[:3]), p, lbx, ubx, lbg, ubg)
        assert r["status"] == 0 and r["stats"][0] <= 1e-8
        np.testing.assert_allclose(r["g"], nlp.g(r["x"], p), atol=1e-12)
        assert abs(r["f"] - nlp.f(r["x"], p)) <= 1e-10 * max(1.0, abs(r["f"]))
        ref = nlp.solve_slsqp(nlp.cold_start(p[:3]), p, lbx, ubx, lbg, ubg)
        if np.abs(ref.x - r["x"])[nlp.nX:].max() > 1e-4:
            ref = nlp.solve_slsqp(r["x"], p, lbx, ubx, lbg, ubg)
            assert ref.fun >= r["f"] - 1e-9 * max(1.0, abs(r["f"]))
        assert abs(ref.fun - r["f"]) <= 1e-6 * max(1.0, abs(r["f"]))
        assert np.abs(ref.x - r["x"])[nlp.nX:].max() <= 1e-4


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _t(torch, a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["warp", "cta", "thread"])
@pytest.mark.parametrize("case", ["one", "six"])
def test_gpu_obstacle_family_matches_slsqp(pkg, torch_cuda, case, path):
    """The three kernels that serve this family: the warp-per-instance solver (the default: the obstacle rows take the lanes
    after the pair rows), the CTA-per-instance dense-block solver (nmpc_tuning.force_block_path; the default when the rows do
    not fit a team's lanes) and the thread-per-instance small-OCP solver (nmpc_tuning.thread_min_batch = 1)."""
    torch = torch_cuda
    tuning = {"warp": None, "cta": dict(force_block_path=1), "thread": dict(thread_min_batch=1)}[path]
    if case == "one":      # first scenario geometry, shortened horizon: obstacle between start and goal
        N, T, obs, margin = 15, 0.3, np.array([[0.45, 0.5, 0.3]]), 0.05
        P = np.array([[0.0, 0.0, 0.6, 1.2, 1.3, 0.0], [0.1, -0.1, 0.9, 1.0, 1.4, 0.3]])
    else:                  # third scenario: six obstacles (:97-119), margin 0.1 (:175)
        N, T, obs, margin = 20, 0.3, _obs(THIRD), 0.1
        P = np.array([[0.0, 0.6, 1.57, 0.1, 3.9, 1.57], [0.3, 0.7, 1.2, -0.2, 3.6, 1.57]])
    v_max, w_max = 0.2, np.pi / 4
    prob = pkg.Problem(1, N, T, obstacles=obs, tuning=tuning)
    nlp = ObstacleNLP(N, T, obs)
    assert (prob.n, prob.mg) == (nlp.n, nlp.mg)
    lbx, ubx, lbg, ubg = prob.bounds_obstacles(margin, v_max, w_max)
    for a, b in zip((lbx, ubx, lbg, ubg), nlp.bounds(margin, v_max, w_max)):
        np.testing.assert_array_equal(a, b)
    x0 = prob.cold_start(P[:, :3])
    out = prob.solve(_t(torch, x0), _t(torch, P), _t(torch, lbx), _t(torch, ubx), _t(torch, lbg), _t(torch, ubg))
    torch.cuda.synchronize()
    assert (out["status"].cpu().numpy() == 0).all(), (out["status"], out["iters"])
    assert out["stats"][:, 0].max().item() <= 1e-8
    x, f, g = out["x"].cpu().numpy(), out["f"].cpu().numpy(), out["g"].cpu().numpy()
    for b in range(P.shape[0]):
        np.testing.assert_allclose(g[b], nlp.g(x[b], P[b]), rtol=0, atol=1e-10)        # same rows, same layout
        assert abs(nlp.f(x[b], P[b]) - f[b]) <= 1e-9 * max(1.0, abs(f[b]))
        assert np.abs(g[b][lbg == ubg]).max() <= 1e-6 and (g[b][lbg != ubg] >= margin - 1e-6).all()
        ref = nlp.solve_slsqp(nlp.cold_start(P[b, :3]), P[b], lbx, ubx, lbg, ubg)
        du = np.abs(ref.x - x[b])[nlp.nX:].max()
        if abs(ref.fun - f[b]) > 1e-6 * max(1.0, abs(f[b])) or du > 1e-4:
            # SLSQP's stopping rule is on f: along the flat directions (omega near the end of the horizon, weight 0.05) it can
            # stop 1e-3 away from the minimiser.  Polished from the product's point it must stay there and find nothing lower.
            ref = nlp.solve_slsqp(x[b], P[b], lbx, ubx, lbg, ubg)
            du = np.abs(ref.x - x[b])[nlp.nX:].max()
            assert ref.fun >= f[b] - 1e-9 * max(1.0, abs(f[b])), (ref.fun, f[b])
        assert abs(ref.fun - f[b]) <= 1e-6 * max(1.0, abs(f[b])), (ref.fun, f[b])
        assert du <= 1e-4, du
    # and against the C oracle (the same restated IPOPT algorithm): same point, same multipliers, iteration counts within 3
    from oracle.oracle_lib import Oracle
    orc = Oracle(1, N, T, obstacles=obs)
    refb = orc.solve_batch(x0, P, lbx, ubx, lbg, ubg, want_duals=True)
    assert (refb["status"] == 0).all()
    assert np.abs(refb["x"] - x)[:, nlp.nX:].max() <= 1e-4
    assert (np.abs(refb["f"] - f) <= 1e-6 * np.maximum(1.0, np.abs(f))).all()
    assert np.abs(out["iters"].cpu().numpy() - refb["iters"]).max() <= 3, (out["iters"].cpu().numpy(), refb["iters"])
    np.testing.assert_allclose(out["lam_g"].cpu().numpy(), refb["lam_g"], rtol=0, atol=1e-5 * max(1.0, np.abs(refb["lam_g"]).max()))


@pytest.mark.gpu
def test_gpu_first_scenario_closed_loop_full_size(pkg, torch_cuda):
    """first_scenario_mpc_obstacle_avoidance.py at its own size (T = 0.1, N = 100, :58-63,96-99,150,160,198): closed loop from
    (0, 0, 0) towards (1.5, 1.5, 0) past the obstacle at (0.4, 1.1); every step solves to 1e-8 and keeps the clearance."""
    torch = torch_cuda
    N, T, margin = 100, 0.1, 0.05
    obs = np.array([[0.4, 1.1, 0.15 + ROB_DIM]])
    prob = pkg.Problem(1, N, T, obstacles=obs)
    lbx, ubx, lbg, ubg = prob.bounds_obstacles(margin, 0.2, np.pi / 4)
    P = np.array([[0.0, 0.0, 0.0, 1.5, 1.5, 0.0], [0.0, 0.8, 0.5, 1.5, 1.5, 0.0]])
    res = pkg.closed_loop(prob, _t(torch, P), _t(torch, lbx), _t(torch, ubx), _t(torch, lbg), _t(torch, ubg), steps=30, tol=5e-2)
    torch.cuda.synchronize()
    assert (res["status"].cpu().numpy() <= 1).all(), res["status"].cpu().numpy()
    traj = res["traj"].cpu().numpy()
    clear = np.hypot(traj[..., 0] - obs[0, 0], traj[..., 1] - obs[0, 1]) - obs[0, 2]
    assert clear.min() >= margin - 1e-6, clear.min()
    d0 = np.hypot(traj[0, :, 0] - 1.5, traj[0, :, 1] - 1.5)
    d1 = np.hypot(traj[-1, :, 0] - 1.5, traj[-1, :, 1] - 1.5)
    assert (d1 < d0 - 0.2).all(), (d0, d1)      # 30 steps of 0.1 s at <= 0.2 m/s: at most 0.6 m, less while turning past the obstacle


@pytest.mark.gpu
def test_gpu_obstacle_family_through_the_nlpsol_shim(pkg, torch_cuda):
    """The scripts' call surface for this family: descriptor in place of the SX block, then the unchanged
    sol = solver(x0=, lbx=, ubx=, lbg=, ubg=, p=) / sol['x'][a:b].full() (first_scenario_mpc_obstacle_avoidance.py:143,220-231)."""
    N, T, margin = 15, 0.3, 0.05
    obs = [[0.45, 0.5, 0.3]]
    solver = pkg.nlpsol("solver", "ipopt", {"family": "unicycle_obstacles", "Nr": 1, "N": N, "T": T, "obstacles": obs},
                        {"print_time": 0, "ipopt": {"max_iter": 2000, "print_level": 0, "acceptable_tol": 1e-8, "acceptable_obj_change_tol": 1e-6}})
    nlp = ObstacleNLP(N, T, obs)
    lbx, ubx, lbg, ubg = nlp.bounds(margin, 0.2, np.pi / 4)
    p = np.array([0.0, 0.0, 0.6, 1.2, 1.3, 0.0])
    sol = solver(x0=nlp.cold_start(p[:3]).reshape(-1, 1), lbx=lbx.reshape(-1, 1), ubx=ubx.reshape(-1, 1), lbg=lbg.reshape(1, -1),
                 ubg=ubg.reshape(1, -1), p=p)
    assert solver.stats()["success"]
    u = sol["x"][3 * (N + 1):].full()
    assert u.shape == (2 * N, 1)
    w = np.asarray(sol["x"].full()).ravel()
    np.testing.assert_allclose(np.asarray(sol["g"].full()).ravel(), nlp.g(w, p), atol=1e-10)
    with pytest.raises(pkg.NmpcError):
        solver.problem.jac_pattern()       # the stand-alone derivative record is not offered for this family


@pytest.mark.gpu
def test_gpu_two_robots_with_obstacles_on_the_warp_path(pkg, torch_cuda):
    """Pair rows and obstacle rows together (one pair row + 2 x 2 obstacle rows per stage) on the warp-per-instance path, against
    the C oracle and, forced onto it, the dense-block path: two robots swap sides around two obstacles."""
    from oracle.oracle_lib import Oracle
    torch = torch_cuda
    Nr, N, T, margin, dmin = 2, 12, 0.3, 0.05, 0.3
    obs = np.array([[0.0, 0.35, 0.25], [0.0, -0.4, 0.25]])
    rng = np.random.default_rng(2)
    P = np.array([[-0.8, 0.05, 0.0, 0.8, -0.05, 3.1, 0.8, 0.0, 0.0, -0.8, 0.0, 3.1]]) + np.concatenate([0.05 * rng.normal(size=(6, 6)), np.zeros((6, 6))], axis=1)
    outs = []
    for tuning in (None, dict(force_block_path=1)):
        prob = pkg.Problem(Nr, N, T, obstacles=obs, tuning=tuning)
        lbx, ubx, lbg, ubg = prob.bounds_obstacles(margin, 0.2, np.pi / 4, dmin=dmin)
        x0 = prob.cold_start(P[:, :6])
        out = prob.solve(_t(torch, x0), _t(torch, P), _t(torch, lbx), _t(torch, ubx), _t(torch, lbg), _t(torch, ubg))
        torch.cuda.synchronize()
        outs.append(out)
    orc = Oracle(Nr, N, T, obstacles=obs)
    ref = orc.solve_batch(x0, P, lbx, ubx, lbg, ubg, want_duals=True)
    assert (ref["status"] == 0).all()
    nX = 6 * (N + 1)
    for out in outs:
        assert (out["status"].cpu().numpy() == 0).all() and out["stats"][:, 0].max().item() <= 1e-8
        x, f, g = out["x"].cpu().numpy(), out["f"].cpu().numpy(), out["g"].cpu().numpy()
        assert np.abs(x - ref["x"])[:, nX:].max() <= 1e-4
        assert (np.abs(f - ref["f"]) <= 1e-6 * np.maximum(1.0, np.abs(ref["f"]))).all()
        np.testing.assert_allclose(g, ref["g"], atol=1e-7)
        assert np.abs(out["iters"].cpu().numpy() - ref["iters"]).max() <= 3
        ineq = lbg != ubg
        assert (g[:, ineq] >= lbg[ineq][None] - 1e-6).all()

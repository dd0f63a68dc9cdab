"""The product's device source (csrc/solver_body.cuh) stepped through a CPU emulation of a warp
(tests/emul/) and compared with the oracle: CPU-side evidence that the kernel's algorithm is the
oracle's, iteration by iteration.  Test infrastructure only -- the product never runs this way."""
import numpy as np
import pytest

from oracle.oracle_lib import Oracle
from tests.emul.emul_lib import emu_solve

CASES = [
    (1, 25, 0.25, 0.3, [0, 0, 0, 2.5, 2.0, 1.57]),
    (2, 10, 0.1, 0.25, [-1, -1, 0.785, 1, 1, 2.356, 1, 1, 0.785, -1, -1, -2.356]),
    (3, 8, 0.3, 0.3, [-1, -1, 1.57, 0, -1, 1.57, 1, -1, 1.57, 1, 2, 0, 0, 1, 0, -1, 0.5, 0]),
]


@pytest.mark.parametrize("Nr,N,T,dmin,p", CASES)
@pytest.mark.parametrize("reverse", [0, 1])
def test_emulated_kernel_follows_the_oracle(Nr, N, T, dmin, p, reverse):
    p = np.asarray(p, float)
    o = Oracle(Nr, N, T)
    lbx, ubx, lbg, ubg = o.bounds(dmin, 0.22, 2.84)
    w0 = o.cold_start(p[:3 * Nr])
    r = o.solve(w0, p, lbx, ubx, lbg, ubg, trace=True)
    e = emu_solve(Nr, N, T, w0, p, lbx, ubx, lbg, ubg, trace=True, reverse=reverse)
    assert e["rc"] == 0 and e["status"][0] == r["status"] == 0
    assert abs(int(e["iters"][0]) - r["iters"]) <= 2
    nX = 3 * Nr * (N + 1)
    assert np.abs(e["x"][0] - r["x"])[nX:].max() <= 1e-7
    assert abs(e["f"][0] - r["f"]) / r["f"] <= 1e-9
    np.testing.assert_allclose(e["lam_g"][0], r["lam_g"], atol=1e-5)
    m = min(int(e["iters"][0]), r["iters"]) - 3
    # same barrier parameter, step sizes, regularisation and line-search counts along the way
    np.testing.assert_allclose(e["trace"][0][:m, [0, 6, 7]], r["trace"][:m, [0, 6, 7]], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(e["trace"][0][:m, [2, 3, 4, 5]], r["trace"][:m, [2, 3, 4, 5]], rtol=1e-5, atol=1e-9)


def test_emulated_kernel_six_robots_short_horizon():
    Nr, N, T = 6, 6, 0.3
    o = Oracle(Nr, N, T)
    lbx, ubx, lbg, ubg = o.bounds(0.3, 0.22, 2.84)
    s3 = np.sqrt(3) / 2
    start = np.array([[s3, 0.5, -2.618], [0, 1, -1.571], [-s3, 0.5, -0.524], [-s3, -0.5, 0.524], [0, -1, 1.571], [s3, -0.5, 2.618]])
    goal = -start.copy(); goal[:, 2] = start[:, 2]
    p = np.concatenate([start.ravel(), goal.ravel()])
    w0 = o.cold_start(p[:18])
    r = o.solve(w0, p, lbx, ubx, lbg, ubg)
    e = emu_solve(Nr, N, T, w0, p, lbx, ubx, lbg, ubg)
    assert e["status"][0] == r["status"] == 0
    assert np.abs(e["x"][0] - r["x"])[18 * 7:].max() <= 1e-6
    assert int(e["stats"][0][8]) == int(r["stats"][8])      # same number of factorisations


def test_emulated_kernel_rejects_bad_bounds():
    o = Oracle(2, 3, 0.1)
    lbx, ubx, lbg, ubg = o.bounds(0.25, 0.22, 2.84)
    p = np.array([0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0.0])
    bad = lbx.copy(); bad[0] = 20.0
    e = emu_solve(2, 3, 0.1, o.cold_start(p[:6]), p, bad, ubx, lbg, ubg)
    assert e["rc"] == -2 and e["status"][0] == -2


def test_emulated_two_warp_team_eight_robots():
    """Nr = 8 (mpc_online_casadi_tb3_eight_multi_centralized_collision_free.py:148-156: N = 5, T = 0.02): 64-lane team."""
    Nr, N, T = 8, 5, 0.02
    o = Oracle(Nr, N, T)
    lbx, ubx, lbg, ubg = o.bounds(0.25, 0.22, 2.84)
    a = np.arange(8) * np.pi / 4
    st = np.stack([np.cos(a), np.sin(a), (a + 2 * np.pi) % (2 * np.pi) - np.pi], 1)
    g = -st.copy(); g[:, 2] = st[:, 2]
    p = np.concatenate([st.ravel() + 0.01 * np.sin(np.arange(24)), g.ravel()])
    w0 = o.cold_start(p[:24])
    r = o.solve(w0, p, lbx, ubx, lbg, ubg)
    for rev in (0, 1):
        e = emu_solve(Nr, N, T, w0, p, lbx, ubx, lbg, ubg, reverse=rev)
        assert e["status"][0] == r["status"] == 0 and int(e["iters"][0]) == r["iters"]
        assert np.abs(e["x"][0] - r["x"]).max() <= 1e-9


@pytest.mark.parametrize("case", ["one", "six", "two_robots"])
@pytest.mark.parametrize("reverse", [0, 1])
def test_emulated_kernel_obstacle_family(case, reverse):
    """Static circular obstacles (first_/third_scenario_mpc_obstacle_avoidance.py:96-152) on the warp-per-instance path: the
    obstacle rows occupy the lanes after the pair rows.  Same iterates as the C oracle (its restated IPOPT with obstacle rows),
    in the scripts' g layout (no inequality rows in block 0)."""
    third = [(-0.6, 3.3, 0.2), (0.6, 3.3, 0.125), (0.0, 2.3, 0.15), (1.0, 2.3, 0.15), (-0.6, 1.3, 0.2), (0.6, 1.3, 0.175)]
    if case == "one":
        Nr, N, T, obs, margin, dmin = 1, 15, 0.3, np.array([[0.45, 0.5, 0.3]]), 0.05, 0.0
        p = np.array([0.0, 0.0, 0.6, 1.2, 1.3, 0.0])
    elif case == "six":
        Nr, N, T, obs, margin, dmin = 1, 20, 0.3, np.array([[x, y, r + 0.15] for x, y, r in third]), 0.1, 0.0
        p = np.array([0.0, 0.6, 1.57, 0.1, 3.9, 1.57])
    else:      # two robots swapping sides around two obstacles: pair row + 2 x 2 obstacle rows per stage
        Nr, N, T, obs, margin, dmin = 2, 12, 0.3, np.array([[0.0, 0.35, 0.25], [0.0, -0.4, 0.25]]), 0.05, 0.3
        p = np.array([-0.8, 0.05, 0.0, 0.8, -0.05, 3.1, 0.8, 0.0, 0.0, -0.8, 0.0, 3.1])
    o = Oracle(Nr, N, T, obstacles=obs)
    ns, M, nobs = 3 * Nr, Nr * (Nr - 1) // 2, len(obs)
    inf = np.inf
    lbx = np.concatenate([np.tile([-10.0, -10.0, -2 * np.pi], Nr * (N + 1)), np.tile([-0.2, -np.pi / 4], Nr * N)])
    ubx = -lbx
    blk_lo = np.concatenate([np.zeros(ns), np.full(M, dmin * dmin), np.full(Nr * nobs, margin)])
    blk_hi = np.concatenate([np.zeros(ns), np.full(M + Nr * nobs, inf)])
    lbg, ubg = np.concatenate([np.zeros(ns), np.tile(blk_lo, N)]), np.concatenate([np.zeros(ns), np.tile(blk_hi, N)])
    assert o.mg == len(lbg)
    w0 = o.cold_start(p[:ns])
    r = o.solve(w0, p, lbx, ubx, lbg, ubg)
    e = emu_solve(Nr, N, T, w0, p, lbx, ubx, lbg, ubg, reverse=reverse, obstacles=obs)
    assert e["rc"] == 0 and e["status"][0] == r["status"] == 0
    assert abs(int(e["iters"][0]) - r["iters"]) <= 2
    nX = ns * (N + 1)
    assert np.abs(e["x"][0] - r["x"])[nX:].max() <= 1e-7
    assert abs(e["f"][0] - r["f"]) / max(1.0, r["f"]) <= 1e-9
    np.testing.assert_allclose(e["g"][0], r["g"], atol=1e-9)
    np.testing.assert_allclose(e["lam_g"][0], r["lam_g"], atol=1e-5)
    assert e["g"][0][lbg != ubg].min() >= margin * 0 + min(margin, dmin * dmin if M else margin) - 1e-8


def test_emulated_kernel_long_horizon_single_robot():
    """decentralized_first_scenario.py:94-95,190-192: N = 200, T = 0.05, positions boxed to +-2 -- 201 stage records per instance."""
    Nr, N, T = 1, 200, 0.05
    o = Oracle(Nr, N, T)
    lbx, ubx, lbg, ubg = o.bounds(0.15, 0.22, 2.84)
    lbx[:3 * (N + 1)].reshape(-1, 3)[:, :2] = -2.0
    ubx[:3 * (N + 1)].reshape(-1, 3)[:, :2] = 2.0
    P = np.array([[0, 0, 0, 1.0, 0.5, 0.0], [-0.5, 0.5, 3.14, -0.5, -0.75, 0.785]])
    w0 = np.stack([o.cold_start(q[:3]) for q in P])
    r = o.solve_batch(w0, P, lbx, ubx, lbg, ubg)
    e = emu_solve(Nr, N, T, w0, P, lbx, ubx, lbg, ubg)
    assert e["rc"] == 0 and np.all(e["status"] == 0) and np.all(r["status"] == 0)
    assert np.abs(e["iters"] - r["iters"]).max() <= 2
    assert np.abs(e["x"] - r["x"])[:, 3 * (N + 1):].max() <= 1e-6
    assert (np.abs(e["f"] - r["f"]) / r["f"]).max() <= 1e-9


def test_emulated_kernel_per_instance_bounds():
    """bounds_batched = 1: each instance reads its own bound rows (its own stage-major copy in the scratch)."""
    Nr, N, T = 3, 6, 0.2
    o = Oracle(Nr, N, T)
    rows = [o.bounds(dmin, v, w) for dmin, v, w in ((0.3, 0.22, 2.84), (0.15, 0.4, 1.0), (0.35, 0.1, 2.0))]
    lbx, ubx, lbg, ubg = (np.stack([q[i] for q in rows]) for i in range(4))
    p = np.array([-1, -1, 1.57, 0, -1, 1.57, 1, -1, 1.57, 1, 2, 0, 0, 1, 0, -1, 0.5, 0], float)
    P = np.stack([p, p, p])
    w0 = np.stack([o.cold_start(p[:9])] * 3)
    r = o.solve_batch(w0, P, lbx, ubx, lbg, ubg)
    e = emu_solve(Nr, N, T, w0, P, lbx, ubx, lbg, ubg)
    assert e["rc"] == 0 and np.all(e["status"] == r["status"]) and np.all(r["status"] == 0)
    nX = 3 * Nr * (N + 1)
    assert np.abs(e["x"] - r["x"])[:, nX:].max() <= 1e-6
    assert np.abs(e["x"][:, nX::2]).max(axis=1) == pytest.approx([0.22, 0.4, 0.1], abs=1e-6)
    for b in range(3):     # and the shared-bounds entry gives the same numbers
        one = emu_solve(Nr, N, T, w0[b], P[b], lbx[b], ubx[b], lbg[b], ubg[b])
        np.testing.assert_array_equal(one["x"][0], e["x"][b])

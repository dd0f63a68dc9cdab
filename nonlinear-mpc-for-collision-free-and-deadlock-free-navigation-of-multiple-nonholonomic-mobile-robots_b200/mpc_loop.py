"""The reference's MPC loop body, solver-agnostic (host code).

Mirrors AllScripts/centralized_six_robots_implementation.py:416-465 (and casadi_test.py:143-180,
whose in-process Euler plant replaces the ROS odometry):

    while ||x0 - xs|| > tol:
        p  = [x0; xs]                                        :419
        w0 = [vec(X0'); vec(u0')]                            :423
        sol = solver(x0=w0, p=p, lbx, ubx, lbg, ubg)         :432
        u  = reshape(sol['x'][ns(N+1):], nc, N)'             :436-437
        apply u[0]; u0 = [u[1:]; u[-1]]  (shift)             :160-169,450
        x0 <- x0 + T f(x0, u[0])         (plant)             casadi_test.py:17-26
        X0 = [X[1:]; X[N-1]]                                 :465  (row N-1, not N)
"""
from __future__ import annotations

import numpy as np

SCENARIOS = {
    # id: (Nr, T, N, dmin, v_max, w_max, start, goal, stop_tol)        SURVEY.md Appendix C
    "C-1": (1, 0.25, 25, 0.3, 0.22, 2.84, [0, 0, 0], [2.5, 2.0, 1.57], 5e-2),                        # casadi_test.py:34-39,115-117
    "C-2": (2, 0.1, 50, 0.25, 0.22, 2.84, [-1, -1, 0.785, 1, 1, 2.356], [1, 1, 0.785, -1, -1, -2.356], 5e-2),   # second_scenario.py
    "C-3": (3, 0.05, 50, 0.3, 0.22, 2.84, [-1, -1, 1.57, 0, -1, 1.57, 1, -1, 1.57], [2, 2, 0, 2, 1, 0, 2, 0, 0], 5e-2),   # third_scenario.py
    "C-4": (4, 0.1, 50, 0.3, 0.22, 2.84, [-1, 1, -0.785, 1, 1, -2.356, 1, -1, 2.356, -1, -1, 0.785],
            [1, -1, -0.785, -1, -1, -2.356, -1, 1, 2.356, 1, 1, 0.785], 5e-2),                        # fourth_scenario.py (square, diagonal swap)
    "C-6": (6, 0.3, 35, 0.3, 0.22, 2.84,
            [0.866, 0.5, -2.618, 0, 1, -1.571, -0.866, 0.5, -0.524, -0.866, -0.5, 0.524, 0, -1, 1.571, 0.866, -0.5, 2.618],
            [-0.866, -0.5, -2.618, 0, -1, -1.571, 0.866, -0.5, -0.524, 0.866, 0.5, 0.524, 0, 1, 1.571, -0.866, 0.5, 2.618], 1e-1),  # sixth_scenario.py
    # fifth_scenario.py:115-123 (T, N, dmin, bounds), :253-254 (start: V formation at x < 0), :268-269 (goal: mirrored V), :297 (stop 1e-1)
    "C-5": (5, 0.1, 35, 0.3, 0.22, 2.84, [-0.5, 1.0, 0, -1.0, 0.5, 0, -1.5, 0, 0, -1.0, -0.5, 0, -0.5, -1.0, 0],
            [0.5, -1.0, 0, 1.0, -0.5, 0, 1.5, 0, 0, 1.0, 0.5, 0, 0.5, 1.0, 0], 1e-1),
    # centralized_two_robots_implementation.py:101-109, :213-214 (frame origins = true starts), :224 (goal), :252 (stop 5e-2)
    "C-2r": (2, 0.05, 70, 0.15, 0.22, 2.84, [-0.7112, -0.7112, 0.785, 0.7112, 0.7112, -2.356],
             [0.7112, 0.7112, 0.785, -0.7112, -0.7112, -2.356], 5e-2),
    # centralized_six_robots_implementation.py:197-205 (dmin 0.4, v_max 0.15, omega_max 1.5), :364-369 (frame origins = true
    # starts), :386-388 (goal), :416 (stop 1e-1): the real-robot constants
    "C-6r": (6, 0.3, 35, 0.4, 0.15, 1.5,
             [0.7, 0.4, -2.618, 0.0, 0.8, -1.57, -0.7, 0.4, -0.523, -0.7, -0.4, 0.523, 0.0, -0.8, 1.57, 0.7, -0.4, 2.618],
             [-0.7, -0.4, -2.618, 0.0, -0.8, -1.57, 0.7, -0.4, -0.523, 0.7, 0.4, 0.523, 0.0, 0.8, 1.57, -0.7, 0.4, 2.618], 1e-1),
    # mpc_online_casadi_tb3_ten_multi_centralized_collision_avoidance.py:169-177 (T 0.1, N 20, dmin 0.3), :409-410 (goal: two
    # rows of five), :438 (stop 1e-1).  The script's start poses are unusable as written (robots 6-10 repeat 1-5, :389-399):
    # robots 1-5 keep theirs (four corners and the centre), robots 6-10 start on the edge midpoints and at (2, 0).
    "C-10": (10, 0.1, 20, 0.3, 0.22, 2.84,
             [-1, 1, -0.785, 1, 1, -2.356, 1, -1, 2.356, -1, -1, 0.785, 0, 0, 0,
              0, 1, -1.571, 1, 0, 3.1416, 0, -1, 1.571, -1, 0, 0, 2, 0, 3.1416],
             [-1.5, 1, 1.57, -0.5, 1, 1.57, 0.5, 1, 1.57, 1.5, 1, 1.57, 2.5, 1, 1.57,
              -1.5, -1, -1.57, -0.5, -1, -1.57, 0.5, -1, -1.57, 1.5, -1, -1.57, 2.5, 2.5, 0.0], 1e-1),
}


def bounds(Nr, N, dmin, v_max, w_max, xy_box=10.0):
    """args = {lbx, ubx, lbg, ubg} exactly as the reference shapes them (:349-352)."""
    ns, nc, M = 3 * Nr, 2 * Nr, Nr * (Nr - 1) // 2
    inf = np.inf
    lbx = np.concatenate([np.tile([-xy_box, -xy_box, -inf], Nr * (N + 1)), np.tile([-v_max, -w_max], Nr * N)]).reshape(-1, 1)
    lbg = np.tile(np.concatenate([np.zeros(ns), np.full(M, dmin * dmin)]), N + 1).reshape(1, -1)
    ubg = np.tile(np.concatenate([np.zeros(ns), np.full(M, inf)]), N + 1).reshape(1, -1)
    return dict(lbx=lbx, ubx=-lbx, lbg=lbg, ubg=ubg)


def run_mpc(solver, Nr, T, N, start, goal, args, tol, max_steps):
    """`solver` has nlpsol's call surface and returns a dict whose 'x' supports [a:b] and .full()."""
    ns, nc = 3 * Nr, 2 * Nr
    x0 = np.asarray(start, float).reshape(ns, 1)
    xs = np.asarray(goal, float).reshape(ns, 1)
    u0 = np.zeros((N, nc))                              # first guess: zeros (:398)
    X0 = np.tile(x0.T, (N + 1, 1))                      # X_k = x0 (:400)
    xx, u_cl = [x0[:, 0].copy()], []
    it = 0
    while np.linalg.norm(x0 - xs) > tol and it < max_steps:
        p = np.concatenate([x0, xs], axis=0)
        w0 = np.concatenate([X0.reshape(-1, 1), u0.reshape(-1, 1)], axis=0)       # row-per-stage == column-major of X (ns x N+1)
        sol = solver(x0=w0, p=p, lbx=args["lbx"], ubx=args["ubx"], lbg=args["lbg"], ubg=args["ubg"])
        u = np.asarray(sol["x"][ns * (N + 1):].full()).reshape(N, nc)             # (:436-437)
        X = np.asarray(sol["x"][:ns * (N + 1)].full()).reshape(N + 1, ns)         # (:460-461)
        u_cl.append(u[0].copy())
        u0 = np.concatenate([u[1:], u[-1:]], axis=0)                              # shift()
        st = x0[:, 0].copy()
        for i in range(Nr):                                                       # Euler plant
            th = st[3 * i + 2]
            st[3 * i] += T * u[0, 2 * i] * np.cos(th)
            st[3 * i + 1] += T * u[0, 2 * i] * np.sin(th)
            st[3 * i + 2] += T * u[0, 2 * i + 1]
        x0 = st.reshape(ns, 1)
        xx.append(st.copy())
        X0 = np.concatenate([X[1:], X[N - 1:N]], axis=0)                          # (:465)
        it += 1
    return np.asarray(xx), np.asarray(u_cl)


def min_pair_distance(xx, Nr):
    if Nr < 2:
        return np.inf
    pos = xx.reshape(len(xx), Nr, 3)[:, :, :2]
    d = np.linalg.norm(pos[:, :, None] - pos[:, None], axis=-1) + np.eye(Nr)[None] * 1e9
    return float(d.min())

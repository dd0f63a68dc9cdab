"""CPU restatement of the reference's static-obstacle NLP (family F) -- TEST INFRASTRUCTURE ONLY.

first_scenario_mpc_obstacle_avoidance.py:75-152 and third_scenario_mpc_obstacle_avoidance.py:95-175: ONE unicycle, the
same Euler multiple-shooting template as the multi-robot scripts, plus per stage k = 0..N-1 one row per static obstacle

    sqrt((x_k - ox)^2 + (y_k - oy)^2) - rob_dim - r_obs   >=  margin            (:125, margin 0.05 / 0.1 in lbg :150)

Row layout of g (as the scripts build it): [X_0 - x0bar (3)], then per k: [X_{k+1} - X_k - T f(X_k, U_k) (3); obstacle rows].
theta is boxed to +-2 pi (:151).  Only tests/ may import this module; the product never does.

PARITY UNPINNED, like the rest of oracle/: CasADi / IPOPT are not installable here and the reference ships no recorded
outputs.  The independent solver for this family is SciPy's SLSQP on this restatement (analytic derivatives), polished
to a KKT point; the product's answer is compared with it and re-checked against the NLP itself.
"""
import numpy as np
from scipy.optimize import minimize


class ObstacleNLP:
    def __init__(self, N, T, obstacles, Q=(1.0, 5.0, 0.1), R=(0.5, 0.05)):
        self.N, self.T = int(N), float(T)
        self.obs = np.asarray(obstacles, float).reshape(-1, 3)      # (ox, oy, clearance = rob_dim + r_obs)
        self.no = self.obs.shape[0]
        self.Q, self.R = np.asarray(Q, float), np.asarray(R, float)
        self.nX = 3 * (self.N + 1)
        self.n = self.nX + 2 * self.N
        self.blk = 3 + self.no
        self.mg = 3 + self.N * self.blk

    def split(self, w):
        return w[:self.nX].reshape(self.N + 1, 3), w[self.nX:].reshape(self.N, 2)

    def f(self, w, p):
        X, U = self.split(w)
        e = X[:self.N] - p[3:6]
        return float((e * e * self.Q).sum() + (U * U * self.R).sum())

    def grad_f(self, w, p):
        X, U = self.split(w)
        gX = np.zeros_like(X)
        gX[:self.N] = 2 * self.Q * (X[:self.N] - p[3:6])
        return np.concatenate([gX.ravel(), (2 * self.R * U).ravel()])

    def g(self, w, p):
        X, U = self.split(w)
        out = np.empty(self.mg)
        out[:3] = X[0] - p[:3]
        for k in range(self.N):
            x, y, th = X[k]
            v, om = U[k]
            o = 3 + k * self.blk
            out[o:o + 3] = X[k + 1] - (X[k] + self.T * np.array([v * np.cos(th), v * np.sin(th), om]))
            out[o + 3:o + self.blk] = np.hypot(x - self.obs[:, 0], y - self.obs[:, 1]) - self.obs[:, 2]
        return out

    def jac_g(self, w, p):
        X, U = self.split(w)
        J = np.zeros((self.mg, self.n))
        J[0:3, 0:3] = np.eye(3)
        T = self.T
        for k in range(self.N):
            x, y, th = X[k]
            v, om = U[k]
            o, ix, iu = 3 + k * self.blk, 3 * k, self.nX + 2 * k
            J[o:o + 3, ix + 3:ix + 6] = np.eye(3)
            J[o:o + 3, ix:ix + 3] = -np.eye(3)
            J[o, ix + 2] = T * v * np.sin(th); J[o + 1, ix + 2] = -T * v * np.cos(th)
            J[o, iu] = -T * np.cos(th); J[o + 1, iu] = -T * np.sin(th); J[o + 2, iu + 1] = -T
            dx, dy = x - self.obs[:, 0], y - self.obs[:, 1]
            rho = np.hypot(dx, dy)
            J[o + 3:o + self.blk, ix] = dx / rho
            J[o + 3:o + self.blk, ix + 1] = dy / rho
        return J

    def bounds(self, margin, v_max, w_max, xy_box=10.0, th_box=2 * np.pi):
        lbx = np.concatenate([np.tile([-xy_box, -xy_box, -th_box], self.N + 1), np.tile([-v_max, -w_max], self.N)])
        blk_lo = np.concatenate([np.zeros(3), np.full(self.no, margin)])
        blk_hi = np.concatenate([np.zeros(3), np.full(self.no, np.inf)])
        return lbx, -lbx, np.concatenate([np.zeros(3), np.tile(blk_lo, self.N)]), np.concatenate([np.zeros(3), np.tile(blk_hi, self.N)])

    def cold_start(self, x0):
        return np.concatenate([np.tile(np.asarray(x0, float), self.N + 1), np.zeros(2 * self.N)])

    def solve_slsqp(self, w0, p, lbx, ubx, lbg, ubg, tol=1e-12, maxiter=2000):
        """Independent solve: SLSQP on the same functions, equality rows (lbg == ubg) and one-sided inequality rows."""
        eq = np.where(lbg == ubg)[0]
        lo = np.where((lbg > -np.inf) & (lbg != ubg))[0]
        cons = [{"type": "eq", "fun": lambda w: self.g(w, p)[eq] - lbg[eq], "jac": lambda w: self.jac_g(w, p)[eq]},
                {"type": "ineq", "fun": lambda w: self.g(w, p)[lo] - lbg[lo], "jac": lambda w: self.jac_g(w, p)[lo]}]
        res = minimize(lambda w: self.f(w, p), w0, jac=lambda w: self.grad_f(w, p), method="SLSQP", constraints=cons,
                       bounds=list(zip(lbx, ubx)), options={"ftol": tol, "maxiter": maxiter})
        return res

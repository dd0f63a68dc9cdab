"""CPU restatement of the reference's Van der Pol multiple-shooting NLP -- TEST INFRASTRUCTURE ONLY.

mpc_pose_control_casadi.py:22-114: x1' = (1 - x2^2) x1 - x2 + u, x2' = x1, L = x1^2 + x2^2 + u^2 (:25-35); T = 10, N = 20, four RK4
steps per interval on the state and on the cost quadrature (:45-59); w = [X_0, U_0, X_1, ..., U_{N-1}, X_N] interleaved (:77-106);
J = sum of the interval quadratures; g = F(X_k, U_k) - X_{k+1}; X_0 = (0, 1) fixed by bounds, -1 <= u <= 1, x1 >= -0.25.
PARITY UNPINNED (no CasADi here, no recorded outputs): the independent solver is SciPy SLSQP on this restatement.
Only tests/ may import this module."""
import numpy as np
from scipy.optimize import minimize


class VdpNLP:
    def __init__(self, N=20, T=10.0, M=4):
        self.N, self.T, self.M = int(N), float(T), int(M)
        self.DT = self.T / self.N / self.M
        self.n, self.mg = 3 * self.N + 2, 2 * self.N

    @staticmethod
    def ode(x, u):
        return np.array([(1 - x[1] ** 2) * x[0] - x[1] + u, x[0]]), x[0] ** 2 + x[1] ** 2 + u ** 2

    def F(self, x, u):
        X, Q, DT = np.array(x, float), 0.0, self.DT
        for _ in range(self.M):
            k1, q1 = self.ode(X, u)
            k2, q2 = self.ode(X + DT / 2 * k1, u)
            k3, q3 = self.ode(X + DT / 2 * k2, u)
            k4, q4 = self.ode(X + DT * k3, u)
            X = X + DT / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
            Q = Q + DT / 6 * (q1 + 2 * q2 + 2 * q3 + q4)
        return X, Q

    def fg(self, w):
        f, g = 0.0, np.empty(self.mg)
        for k in range(self.N):
            xf, qf = self.F(w[3 * k:3 * k + 2], w[3 * k + 2])
            f += qf
            g[2 * k:2 * k + 2] = xf - w[3 * k + 3:3 * k + 5]
        return f, g

    def demo_arrays(self):
        inf = np.inf
        w0, lbw, ubw = [0.0, 1.0], [0.0, 1.0], [0.0, 1.0]
        for _ in range(self.N):
            w0 += [0.0, 0.0, 0.0]; lbw += [-1.0, -0.25, -inf]; ubw += [1.0, inf, inf]
        return np.array(w0), np.array(lbw), np.array(ubw), np.zeros(self.mg), np.zeros(self.mg)

    def solve_slsqp(self, w0, lbw, ubw, tol=1e-13, maxiter=1000):
        cons = [{"type": "eq", "fun": lambda w: self.fg(w)[1]}]
        return minimize(lambda w: self.fg(w)[0], w0, method="SLSQP", constraints=cons, bounds=list(zip(lbw, ubw)),
                        options={"ftol": tol, "maxiter": maxiter})

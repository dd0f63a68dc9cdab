"""NumPy restatement of the reference's multiple-shooting NLP (TEST INFRASTRUCTURE ONLY).

This file is part of ``oracle/``: it may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg.  The product
path (the CUDA library) never touches it.

PARITY UNPINNED: the reference ships no tests, golden vectors or recorded
outputs, and CasADi/IPOPT are not installable here.  What
this file restates is the *problem* the reference hands to ``nlpsol``; the
derivatives are checked against finite differences in ``tests/``.

Reference (all under /root/reference/AllScripts/):
  * decision-vector layout  w = [vec(X) ; vec(U)], column-major
      centralized_six_robots_implementation.py:240-245,339
  * cost                    centralized_six_robots_implementation.py:252-266,314
  * constraints             centralized_six_robots_implementation.py:278,282-331
  * bounds                  centralized_six_robots_implementation.py:349-352
  * single robot            casadi_test.py:34-109
"""
from __future__ import annotations

import numpy as np


class UnicycleNLP:
    """Dimensions and index maps of the Nr-robot, horizon-N unicycle NLP."""

    def __init__(self, Nr, N, T, Q=(1.0, 5.0, 0.1), R=(0.5, 0.05)):
        self.Nr, self.N, self.T = int(Nr), int(N), float(T)
        self.Q = np.asarray(Q, float)
        self.R = np.asarray(R, float)
        self.ns, self.nc = 3 * self.Nr, 2 * self.Nr
        self.M = self.Nr * (self.Nr - 1) // 2
        self.n = self.ns * (self.N + 1) + self.nc * self.N
        self.blk = self.ns + self.M
        self.mg = (self.N + 1) * self.blk
        # pairs in lexicographic order i<j  (d12, d13, ..., d56;
        # centralized_six_robots_implementation.py:288-306)
        self.pairs = [(i, j) for i in range(self.Nr) for j in range(i + 1, self.Nr)]

    # -- index helpers -----------------------------------------------------
    def ix(self, k, i, c):
        return k * self.ns + 3 * i + c

    def iu(self, k, i, c):
        return self.ns * (self.N + 1) + k * self.nc + 2 * i + c

    def split(self, w):
        w = np.asarray(w, float).reshape(-1)
        X = w[: self.ns * (self.N + 1)].reshape(self.N + 1, self.Nr, 3)
        U = w[self.ns * (self.N + 1):].reshape(self.N, self.Nr, 2)
        return X, U

    # -- objective ---------------------------------------------------------
    def f(self, w, p):
        X, U = self.split(w)
        xs = np.asarray(p, float).reshape(-1)[self.ns:].reshape(self.Nr, 3)
        e = X[: self.N] - xs[None]
        return float(np.sum(self.Q * e * e) + np.sum(self.R * U * U))

    def grad_f(self, w, p):
        X, U = self.split(w)
        xs = np.asarray(p, float).reshape(-1)[self.ns:].reshape(self.Nr, 3)
        gX = np.zeros_like(X)
        gX[: self.N] = 2.0 * self.Q * (X[: self.N] - xs[None])
        gU = 2.0 * self.R * U
        return np.concatenate([gX.reshape(-1), gU.reshape(-1)])

    # -- constraints -------------------------------------------------------
    def g(self, w, p):
        X, U = self.split(w)
        x0 = np.asarray(p, float).reshape(-1)[: self.ns]
        out = np.zeros(self.mg)
        out[: self.ns] = X[0].reshape(-1) - x0
        out[self.ns: self.blk] = 3.5
        T = self.T
        for k in range(self.N):
            b = (k + 1) * self.blk
            th = X[k, :, 2]
            v, om = U[k, :, 0], U[k, :, 1]
            nxt = np.stack([X[k, :, 0] + T * v * np.cos(th),
                            X[k, :, 1] + T * v * np.sin(th),
                            th + T * om], axis=1)
            out[b: b + self.ns] = (X[k + 1] - nxt).reshape(-1)
            for q, (i, j) in enumerate(self.pairs):
                dx = X[k, i, 0] - X[k, j, 0]
                dy = X[k, i, 1] - X[k, j, 1]
                out[b + self.ns + q] = dx * dx + dy * dy
        return out

    def jac_g(self, w, p):
        """Dense Jacobian (mg x n); tests derive the CCS triplets from it."""
        X, U = self.split(w)
        J = np.zeros((self.mg, self.n))
        T = self.T
        for r in range(self.ns):
            J[r, r] = 1.0
        for k in range(self.N):
            b = (k + 1) * self.blk
            for i in range(self.Nr):
                th, v = X[k, i, 2], U[k, i, 0]
                rx, ry, rt = b + 3 * i, b + 3 * i + 1, b + 3 * i + 2
                J[rx, self.ix(k + 1, i, 0)] = 1.0
                J[rx, self.ix(k, i, 0)] = -1.0
                J[rx, self.ix(k, i, 2)] = T * v * np.sin(th)
                J[rx, self.iu(k, i, 0)] = -T * np.cos(th)
                J[ry, self.ix(k + 1, i, 1)] = 1.0
                J[ry, self.ix(k, i, 1)] = -1.0
                J[ry, self.ix(k, i, 2)] = -T * v * np.cos(th)
                J[ry, self.iu(k, i, 0)] = -T * np.sin(th)
                J[rt, self.ix(k + 1, i, 2)] = 1.0
                J[rt, self.ix(k, i, 2)] = -1.0
                J[rt, self.iu(k, i, 1)] = -T
            for q, (i, j) in enumerate(self.pairs):
                dx = X[k, i, 0] - X[k, j, 0]
                dy = X[k, i, 1] - X[k, j, 1]
                r = b + self.ns + q
                J[r, self.ix(k, i, 0)] = 2 * dx
                J[r, self.ix(k, j, 0)] = -2 * dx
                J[r, self.ix(k, i, 1)] = 2 * dy
                J[r, self.ix(k, j, 1)] = -2 * dy
        return J

    def hess_lag(self, w, p, lam_g, sigma=1.0):
        """Dense Hessian of  sigma*f + lam_g' g  (CasADi sign convention)."""
        X, U = self.split(w)
        lam = np.asarray(lam_g, float).reshape(-1)
        H = np.zeros((self.n, self.n))
        T = self.T
        for k in range(self.N):
            b = (k + 1) * self.blk
            for i in range(self.Nr):
                th, v = X[k, i, 2], U[k, i, 0]
                lx, ly = lam[b + 3 * i], lam[b + 3 * i + 1]
                ixx, iyy, itt = self.ix(k, i, 0), self.ix(k, i, 1), self.ix(k, i, 2)
                iv, iw = self.iu(k, i, 0), self.iu(k, i, 1)
                H[ixx, ixx] += sigma * 2 * self.Q[0]
                H[iyy, iyy] += sigma * 2 * self.Q[1]
                H[itt, itt] += sigma * 2 * self.Q[2] + T * v * (lx * np.cos(th) + ly * np.sin(th))
                H[iv, iv] += sigma * 2 * self.R[0]
                H[iw, iw] += sigma * 2 * self.R[1]
                c = T * (lx * np.sin(th) - ly * np.cos(th))
                H[itt, iv] += c
                H[iv, itt] += c
            for q, (i, j) in enumerate(self.pairs):
                mu = lam[b + self.ns + q]
                for c in (0, 1):
                    a, bb = self.ix(k, i, c), self.ix(k, j, c)
                    H[a, a] += 2 * mu
                    H[bb, bb] += 2 * mu
                    H[a, bb] -= 2 * mu
                    H[bb, a] -= 2 * mu
        return H

    # -- bounds as the reference builds them ------------------------------
    def bounds(self, dmin, v_max, w_max, xy_box=10.0):
        """args = {lbx, ubx, lbg, ubg}; centralized_six...py:349-352."""
        inf = np.inf
        lbx = np.concatenate([np.tile([-xy_box, -xy_box, -inf], self.Nr * (self.N + 1)),
                              np.tile([-v_max, -w_max], self.Nr * self.N)])
        ubx = -lbx
        lbg = np.tile(np.concatenate([np.zeros(self.ns), np.full(self.M, dmin * dmin)]), self.N + 1)
        ubg = np.tile(np.concatenate([np.zeros(self.ns), np.full(self.M, inf)]), self.N + 1)
        return lbx, ubx, lbg, ubg

    def cold_start(self, x0):
        """X_k = x0 for all k, U = 0 (centralized_six...py:398-400,423)."""
        x0 = np.asarray(x0, float).reshape(-1)
        return np.concatenate([np.tile(x0, self.N + 1), np.zeros(self.nc * self.N)])

    # -- sparsity (CCS) ----------------------------------------------------
    def jac_ccs(self):
        """(colptr, rowidx) of the structural Jacobian, CasADi-style CCS."""
        rng = np.random.default_rng(1)
        w = rng.uniform(0.3, 1.3, self.n)
        p = rng.uniform(-1, 1, 2 * self.ns)
        S = self.jac_g(w, p) != 0
        return _ccs(S)

    def hess_ccs_lower(self):
        rng = np.random.default_rng(2)
        w = rng.uniform(0.3, 1.3, self.n)
        p = rng.uniform(-1, 1, 2 * self.ns)
        lam = rng.uniform(0.5, 1.5, self.mg)
        S = np.tril(self.hess_lag(w, p, lam) != 0)
        return _ccs(S)


def _ccs(S):
    colptr = [0]
    rows = []
    for c in range(S.shape[1]):
        r = np.nonzero(S[:, c])[0]
        rows.extend(r.tolist())
        colptr.append(len(rows))
    return np.asarray(colptr, np.int32), np.asarray(rows, np.int32)


def shift_warm_start(nlp, w_opt):
    """Reference warm start between MPC steps.

    u0 = [u[1:]; u[-1]]            centralized_six...py:160-169  (shift)
    X0 = [X[1:]; X[N-1]]           centralized_six...py:465  (appends row N-1, not N)
    """
    X, U = nlp.split(w_opt)
    Xn = np.concatenate([X[1:], X[nlp.N - 1: nlp.N]], axis=0)
    Un = np.concatenate([U[1:], U[-1:]], axis=0)
    return np.concatenate([Xn.reshape(-1), Un.reshape(-1)])


def euler_plant(nlp, x0, u0):
    """x <- x + T f(x,u)   casadi_test.py:17-26."""
    x = np.asarray(x0, float).reshape(nlp.Nr, 3).copy()
    u = np.asarray(u0, float).reshape(nlp.Nr, 2)
    th = x[:, 2].copy()
    x[:, 0] += nlp.T * u[:, 0] * np.cos(th)
    x[:, 1] += nlp.T * u[:, 0] * np.sin(th)
    x[:, 2] += nlp.T * u[:, 1]
    return x.reshape(-1)


def synthetic_instances(B, Nr=6, seed=20261018, box=2.0, sep=0.5):
    """BASELINE.md workload: B start/goal sets, rejection sampled."""
    rng = np.random.default_rng(seed)

    def draw():
        while True:
            xy = rng.uniform(-box, box, (Nr, 2))
            d = np.linalg.norm(xy[:, None] - xy[None], axis=-1) + np.eye(Nr) * 1e9
            if d.min() >= sep:
                th = rng.uniform(-np.pi, np.pi, (Nr, 1))
                return np.concatenate([xy, th], axis=1).reshape(-1)

    P = np.empty((B, 6 * Nr))
    for b in range(B):
        P[b, : 3 * Nr] = draw()
        P[b, 3 * Nr:] = draw()
    return P

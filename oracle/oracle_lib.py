"""ctypes binding of the C oracle (oracle/nmpc_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Importable only from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs.  PARITY UNPINNED (see nmpc_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_NSTATS = 12
ORC_NTRACE = 8
STATUS = {0: "SOLVED", 1: "ACCEPTABLE", 2: "MAX_ITER", 3: "INFEASIBLE", 4: "NUMERICAL"}


class Desc(C.Structure):
    _fields_ = [("Nr", C.c_int), ("N", C.c_int), ("T", C.c_double), ("Q", C.c_double * 3), ("R", C.c_double * 2),
                ("nobs", C.c_int), ("obs", C.c_void_p)]


class Opts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("acceptable_tol", C.c_double),
                ("acceptable_iter", C.c_int), ("acceptable_obj_change_tol", C.c_double),
                ("dual_inf_tol", C.c_double), ("constr_viol_tol", C.c_double), ("compl_inf_tol", C.c_double),
                ("mu_init", C.c_double), ("kappa_mu", C.c_double), ("theta_mu", C.c_double),
                ("barrier_tol_factor", C.c_double), ("tau_min", C.c_double), ("bound_push", C.c_double),
                ("bound_frac", C.c_double), ("bound_relax_factor", C.c_double),
                ("bound_mult_init_val", C.c_double), ("constr_mult_init_max", C.c_double),
                ("kappa_sigma", C.c_double), ("kappa_d", C.c_double), ("nlp_scaling_max_gradient", C.c_double),
                ("max_soc", C.c_int), ("max_resto_iter", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "nmpc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        try:
            _LIB = C.CDLL(so)
        except OSError:
            _LIB = C.CDLL(build(force=True))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _LIB.orc_solve.argtypes = [C.POINTER(Desc), C.POINTER(Opts)] + [dp] * 11 + [ip, ip, dp, dp, C.c_int]
        _LIB.orc_solve_batch.argtypes = ([C.POINTER(Desc), C.POINTER(Opts), C.c_int] + [dp] * 6 + [C.c_int]
                                         + [dp] * 5 + [ip, ip, dp, C.c_int])
        _LIB.orc_eval.argtypes = [C.POINTER(Desc)] + [dp] * 8
        _LIB.orc_eval.restype = None
        _LIB.orc_shift.argtypes = [C.POINTER(Desc), dp, dp]
        _LIB.orc_shift.restype = None
        _LIB.orc_plant.argtypes = [C.POINTER(Desc), dp, dp, dp]
        _LIB.orc_plant.restype = None
        _LIB.orc_jac_pattern.argtypes = [C.POINTER(Desc), ip, ip]
        _LIB.orc_hess_pattern.argtypes = [C.POINTER(Desc), ip, ip]
        _LIB.orc_kkt_step.argtypes = [C.POINTER(Desc)] + [dp] * 5 + [C.c_double, dp, dp, C.c_double] + [dp] * 6
        _LIB.orc_default_opts.argtypes = [C.POINTER(Opts)]
        for fn in ("orc_n", "orc_mg", "orc_nnz_jac", "orc_nnz_hess"):
            getattr(_LIB, fn).argtypes = [C.POINTER(Desc)]
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Oracle:
    """Nr-robot, horizon-N unicycle NLP + restated IPOPT on the CPU."""

    def __init__(self, Nr, N, T, Q=(1.0, 5.0, 0.1), R=(0.5, 0.05), obstacles=None, **opts):
        """obstacles: [n_obs, 3] (centre x, y, clearance): the static-obstacle family (first_scenario_mpc_obstacle_avoidance.py)."""
        self.L = lib()
        self.obstacles = None if obstacles is None else np.ascontiguousarray(np.asarray(obstacles, dtype=np.float64).reshape(-1, 3))
        self.d = Desc(int(Nr), int(N), float(T), (C.c_double * 3)(*Q), (C.c_double * 2)(*R),
                      0 if self.obstacles is None else int(self.obstacles.shape[0]),
                      None if self.obstacles is None else self.obstacles.ctypes.data)
        self.o = Opts()
        self.L.orc_default_opts(C.byref(self.o))
        for k, v in opts.items():
            setattr(self.o, k, v)
        self.Nr, self.N, self.T = int(Nr), int(N), float(T)
        self.ns, self.nc = 3 * self.Nr, 2 * self.Nr
        self.M = self.Nr * (self.Nr - 1) // 2
        self.n = self.L.orc_n(C.byref(self.d))
        self.mg = self.L.orc_mg(C.byref(self.d))
        self.nnz_jac = self.L.orc_nnz_jac(C.byref(self.d))
        self.nnz_hess = self.L.orc_nnz_hess(C.byref(self.d))

    # reference bounds (centralized_six_robots_implementation.py:349-352)
    def bounds(self, dmin, v_max, w_max, xy_box=10.0):
        inf = np.inf
        lbx = np.concatenate([np.tile([-xy_box, -xy_box, -inf], self.Nr * (self.N + 1)),
                              np.tile([-v_max, -w_max], self.Nr * self.N)])
        ubx = -lbx
        lbg = np.tile(np.concatenate([np.zeros(self.ns), np.full(self.M, dmin * dmin)]), self.N + 1)
        ubg = np.tile(np.concatenate([np.zeros(self.ns), np.full(self.M, inf)]), self.N + 1)
        return lbx, ubx, lbg, ubg

    def cold_start(self, x0):
        x0 = _f64(x0).reshape(-1)
        return np.concatenate([np.tile(x0, self.N + 1), np.zeros(self.nc * self.N)])

    def eval(self, w, p, lam_g=None, want_jac=True, want_hess=True):
        w, p = _f64(w), _f64(p)
        f = np.zeros(1)
        grad, g = np.zeros(self.n), np.zeros(self.mg)
        jv = np.zeros(self.nnz_jac) if want_jac else None
        hv = np.zeros(self.nnz_hess) if (want_hess and lam_g is not None) else None
        lam = _f64(lam_g) if lam_g is not None else None
        self.L.orc_eval(C.byref(self.d), _dp(w), _dp(p), _dp(lam), _dp(f), _dp(grad), _dp(g), _dp(jv), _dp(hv))
        return dict(f=float(f[0]), grad=grad, g=g, jac=jv, hess=hv)

    def jac_pattern(self):
        cp, ri = np.zeros(self.n + 1, np.int32), np.zeros(self.nnz_jac, np.int32)
        self.L.orc_jac_pattern(C.byref(self.d), _ip(cp), _ip(ri))
        return cp, ri

    def hess_pattern(self):
        cp, ri = np.zeros(self.n + 1, np.int32), np.zeros(self.nnz_hess, np.int32)
        self.L.orc_hess_pattern(C.byref(self.d), _ip(cp), _ip(ri))
        return cp, ri

    def solve(self, x0, p, lbx, ubx, lbg, ubg, trace=False):
        x0, p, lbx, ubx, lbg, ubg = map(_f64, (x0, p, lbx, ubx, lbg, ubg))
        x, f, g = np.zeros(self.n), np.zeros(1), np.zeros(self.mg)
        lam_x, lam_g = np.zeros(self.n), np.zeros(self.mg)
        st, it = np.zeros(1, np.int32), np.zeros(1, np.int32)
        stats = np.zeros(ORC_NSTATS)
        ntr = int(self.o.max_iter) + 1 if trace else 0
        tr = np.zeros((ntr, ORC_NTRACE)) if trace else None
        rc = self.L.orc_solve(C.byref(self.d), C.byref(self.o), _dp(x0), _dp(p), _dp(lbx), _dp(ubx), _dp(lbg),
                              _dp(ubg), _dp(x), _dp(f), _dp(g), _dp(lam_x), _dp(lam_g), _ip(st), _ip(it),
                              _dp(stats), _dp(tr), ntr)
        if rc:
            raise ValueError("orc_solve API error %d" % rc)
        out = dict(x=x, f=float(f[0]), g=g, lam_x=lam_x, lam_g=lam_g, status=int(st[0]), iters=int(it[0]), stats=stats)
        if trace:
            out["trace"] = tr[: int(it[0]) + 1]
        return out

    def solve_batch(self, x0, p, lbx, ubx, lbg, ubg, nthreads=0, want_duals=False):
        x0, p, lbx, ubx, lbg, ubg = map(_f64, (x0, p, lbx, ubx, lbg, ubg))
        B = x0.shape[0]
        batched = 1 if lbx.ndim == 2 else 0
        x, f, g = np.zeros((B, self.n)), np.zeros(B), np.zeros((B, self.mg))
        lam_x = np.zeros((B, self.n)) if want_duals else None
        lam_g = np.zeros((B, self.mg)) if want_duals else None
        st, it = np.zeros(B, np.int32), np.zeros(B, np.int32)
        stats = np.zeros((B, ORC_NSTATS))
        rc = self.L.orc_solve_batch(C.byref(self.d), C.byref(self.o), B, _dp(x0), _dp(p), _dp(lbx), _dp(ubx),
                                    _dp(lbg), _dp(ubg), batched, _dp(x), _dp(f), _dp(g), _dp(lam_x), _dp(lam_g),
                                    _ip(st), _ip(it), _dp(stats), int(nthreads))
        if rc:
            raise ValueError("orc_solve_batch API error %d" % rc)
        return dict(x=x, f=f, g=g, lam_x=lam_x, lam_g=lam_g, status=st, iters=it, stats=stats)

    def shift(self, x_prev):
        xp = _f64(x_prev)
        out = np.zeros(self.n)
        self.L.orc_shift(C.byref(self.d), _dp(xp), _dp(out))
        return out

    def plant(self, state, u0):
        s, u = _f64(state), _f64(u0)
        out = np.zeros(self.ns)
        self.L.orc_plant(C.byref(self.d), _dp(s), _dp(u), _dp(out))
        return out

    def kkt_step(self, p, lbg, ubg, w, lam_g, obj_scale, sig_x, sig_s, delta_w, gx, gs, rg):
        a = list(map(_f64, (p, lbg, ubg, w, lam_g)))
        sx, ss, gx, gs, rg = map(_f64, (sig_x, sig_s, gx, gs, rg))
        dx, ds, yl = np.zeros(self.n), np.zeros(self.mg), np.zeros(self.mg)
        rc = self.L.orc_kkt_step(C.byref(self.d), *[_dp(v) for v in a], float(obj_scale), _dp(sx), _dp(ss),
                                 float(delta_w), _dp(gx), _dp(gs), _dp(rg), _dp(dx), _dp(ds), _dp(yl))
        return rc, dx, ds, yl

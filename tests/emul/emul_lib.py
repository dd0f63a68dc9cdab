"""ctypes binding of the CPU warp emulation (tests/emul/) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.oracle_lib import Desc, Opts, lib as _orc_lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NSTATS, NTRACE = 11, 8


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libnmpc_emul.so"])
        _LIB = C.CDLL(os.path.join(_HERE, "libnmpc_emul.so"))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _LIB.emu_solve.argtypes = ([C.POINTER(Desc), C.POINTER(Opts), C.c_int] + [dp] * 6 + [C.c_int] + [dp] * 5
                                   + [ip, ip, dp, dp, C.c_int, C.c_int, C.c_int, dp])
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def emu_solve(Nr, N, T, x0, p, lbx, ubx, lbg, ubg, Q=(1.0, 5.0, 0.1), R=(0.5, 0.05), trace=False, reverse=0, obstacles=None, **opts):
    """obstacles: [n_obs, 3] (centre x, y, clearance) -> the static-obstacle family (its own g layout, see nmpc_create_obstacles)."""
    L = lib()
    obs = None if obstacles is None else np.ascontiguousarray(np.asarray(obstacles, dtype=np.float64).reshape(-1, 3))
    d = Desc(int(Nr), int(N), float(T), (C.c_double * 3)(*Q), (C.c_double * 2)(*R))
    o = Opts()
    _orc_lib().orc_default_opts(C.byref(o))   # nmpc_opts and orc_opts have the same layout
    for k, v in opts.items():
        setattr(o, k, v)
    f64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    x0, p, lbx, ubx, lbg, ubg = map(f64, (x0, p, lbx, ubx, lbg, ubg))
    x0 = np.atleast_2d(x0); p = np.atleast_2d(p)
    B, n = x0.shape
    mg = lbg.shape[-1]
    x, f, g = np.zeros((B, n)), np.zeros(B), np.zeros((B, mg))
    lam_x, lam_g = np.zeros((B, n)), np.zeros((B, mg))
    st, it = np.full(B, -99, np.int32), np.zeros(B, np.int32)
    stats = np.zeros((B, NSTATS))
    ntr = int(o.max_iter) + 1 if trace else 0
    tr = np.zeros((B, ntr, NTRACE)) if trace else None
    rc = L.emu_solve(C.byref(d), C.byref(o), B, _dp(x0), _dp(p), _dp(lbx), _dp(ubx), _dp(lbg), _dp(ubg),
                     1 if lbx.ndim == 2 else 0, _dp(x), _dp(f), _dp(g), _dp(lam_x), _dp(lam_g),
                     st.ctypes.data_as(C.POINTER(C.c_int)), it.ctypes.data_as(C.POINTER(C.c_int)), _dp(stats),
                     _dp(tr), ntr, int(reverse), 0 if obs is None else int(obs.shape[0]), _dp(obs))
    return dict(rc=rc, x=x, f=f, g=g, lam_x=lam_x, lam_g=lam_g, status=st, iters=it, stats=stats, trace=tr)

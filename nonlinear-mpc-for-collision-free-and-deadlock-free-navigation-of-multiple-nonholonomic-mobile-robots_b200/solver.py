"""Host side of the hot path: the nlpsol-compatible call surface and the batched device API.

Mirrors what the reference does around its solver call
(AllScripts/centralized_six_robots_implementation.py):
    solver = nlpsol('solver', 'ipopt', nlp_prob, opts)        :345-346
    sol = solver(x0=..., p=..., lbx=..., ubx=..., lbg=..., ubg=...)   :432
    sol['x'][a:b] ... .full()                                  :436,440,460
so that the scenario loops (:416-510) run unchanged on top of the CUDA library.  CasADi is not a
dependency: the factory takes a problem descriptor in place of the symbolic dict.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import NSTATS, NTRACE, Desc, NmpcError, check, default_opts, default_tuning, lib


# ----------------------------------------------------------------------------------------------
# the few CasADi names the reference's loop bodies use on numeric data
# ----------------------------------------------------------------------------------------------
class DM:
    """Dense numeric matrix with CasADi's column-major semantics: [a:b] slicing on the flattened
    column-major vector, .full() -> ndarray, usable wherever NumPy expects an array."""

    def __init__(self, a):
        a = np.asarray(a.full() if isinstance(a, DM) else a, dtype=float)
        self._a = a.reshape(-1, 1) if a.ndim < 2 else a

    def full(self):
        return self._a.copy()

    @property
    def shape(self):
        return self._a.shape

    def numel(self):
        return self._a.size

    def __array__(self, dtype=None, copy=None):
        return self._a.astype(dtype) if dtype is not None else self._a

    def __len__(self):
        return self._a.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            return DM(np.atleast_2d(self._a[idx]))
        flat = self._a.reshape(-1, order="F")
        return DM(np.atleast_1d(flat[idx]).reshape(-1, 1))

    def __float__(self):
        return float(self._a.reshape(-1)[0])

    def __repr__(self):
        return "DM(%r)" % (self._a,)


def _num(a):
    return np.asarray(a.full() if isinstance(a, DM) else a, dtype=float)


def reshape(a, *shape):
    """CasADi reshape: column-major."""
    if len(shape) == 1:
        shape = tuple(shape[0])
    return DM(np.reshape(np.atleast_2d(_num(a)), (int(shape[0]), int(shape[1])), order="F"))


def repmat(a, n, m=1):
    return DM(np.tile(np.atleast_2d(_num(a)), (int(n), int(m))))


def vertcat(*args):
    return DM(np.concatenate([np.atleast_2d(_num(a)).reshape(-1, np.atleast_2d(_num(a)).shape[-1]) if np.ndim(_num(a)) > 1
                              else _num(a).reshape(-1, 1) for a in args], axis=0))


def horzcat(*args):
    return DM(np.concatenate([np.atleast_2d(_num(a)) for a in args], axis=1))


def _flat(a, size, name):
    """Accept row or column vectors, scalars (broadcast) and DM; flatten column-major."""
    v = _num(a)
    if v.size == 1:
        return np.full(size, float(v.reshape(-1)[0]))
    v = v.reshape(-1, order="F")
    if v.size != size:
        raise ValueError("%s has %d entries, expected %d" % (name, v.size, size))
    return np.ascontiguousarray(v, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
class Problem:
    """Handle on the CUDA library for one (Nr, N, T, Q, R) problem family."""

    def __init__(self, Nr, N, T, Q=(1.0, 5.0, 0.1), R=(0.5, 0.05), obstacles=None, tuning=None, **opts):
        """obstacles: optional [n_obs, 3] array of static circular obstacles (centre x, y, clearance = rob_dim + r_obs): the
        reference's obstacle-avoidance family (first_scenario_mpc_obstacle_avoidance.py:96-152), see nmpc_create_obstacles.
        tuning: optional dict of nmpc_tuning fields (convoy, ctas_per_sm, force_block_path, thread_min_batch)."""
        self.L = lib()
        self.desc = Desc(int(Nr), int(N), float(T), (C.c_double * 3)(*[float(q) for q in Q]),
                         (C.c_double * 2)(*[float(r) for r in R]))
        self.opts = default_opts(**opts)
        self.h = C.c_void_p()
        self.obstacles = None if obstacles is None else np.ascontiguousarray(np.asarray(obstacles, dtype=np.float64).reshape(-1, 3))
        self.tuning = default_tuning(**(tuning or {}))
        check(self.L.nmpc_create_tuned(C.byref(self.desc), C.byref(self.opts), C.byref(self.tuning),
                                       0 if self.obstacles is None else int(self.obstacles.shape[0]),
                                       None if self.obstacles is None else self.obstacles.ctypes.data, C.byref(self.h)))
        self.nobs = 0 if self.obstacles is None else int(self.obstacles.shape[0])
        self.Nr, self.N, self.T = int(Nr), int(N), float(T)
        self.ns, self.nc = 3 * self.Nr, 2 * self.Nr
        self.M = self.Nr * (self.Nr - 1) // 2
        self.n, self.mg, self.np_ = self.L.nmpc_n(self.h), self.L.nmpc_mg(self.h), self.L.nmpc_np(self.h)
        self.nnz_jac, self.nnz_hess = self.L.nmpc_nnz_jac(self.h), self.L.nmpc_nnz_hess(self.h)
        self._ws = None

    def __del__(self):
        try:
            if self.h:
                self.L.nmpc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # reference bounds (centralized_six_robots_implementation.py:349-352)
    def bounds(self, dmin, v_max, w_max, xy_box=10.0):
        inf = np.inf
        lbx = np.concatenate([np.tile([-xy_box, -xy_box, -inf], self.Nr * (self.N + 1)),
                              np.tile([-v_max, -w_max], self.Nr * self.N)])
        ubx = -lbx
        lbg = np.tile(np.concatenate([np.zeros(self.ns), np.full(self.M, dmin * dmin)]), self.N + 1)
        ubg = np.tile(np.concatenate([np.zeros(self.ns), np.full(self.M, inf)]), self.N + 1)
        return lbx, ubx, lbg, ubg

    def bounds_obstacles(self, margin, v_max, w_max, dmin=0.0, xy_box=10.0, th_box=2.0 * np.pi):
        """Bounds of the obstacle-avoidance scripts (first_scenario_mpc_obstacle_avoidance.py:150-152): theta boxed to +-2 pi,
        equality rows 0, obstacle rows >= margin (0.05 there, 0.1 in the third scenario), pair rows (Nr > 1) >= dmin^2."""
        inf = np.inf
        lbx = np.concatenate([np.tile([-xy_box, -xy_box, -th_box], self.Nr * (self.N + 1)),
                              np.tile([-v_max, -w_max], self.Nr * self.N)])
        ubx = -lbx
        blk_lo = np.concatenate([np.zeros(self.ns), np.full(self.M, dmin * dmin), np.full(self.Nr * self.nobs, margin)])
        blk_hi = np.concatenate([np.zeros(self.ns), np.full(self.M + self.Nr * self.nobs, inf)])
        lbg = np.concatenate([np.zeros(self.ns), np.tile(blk_lo, self.N)])
        ubg = np.concatenate([np.zeros(self.ns), np.tile(blk_hi, self.N)])
        return lbx, ubx, lbg, ubg

    def cold_start(self, x0):
        """X_k = x0 for all k, U = 0 (:398-400,423)."""
        x0 = np.asarray(x0, float)
        if x0.ndim == 1:
            return np.concatenate([np.tile(x0, self.N + 1), np.zeros(self.nc * self.N)])
        return np.concatenate([np.tile(x0, (1, self.N + 1)), np.zeros((x0.shape[0], self.nc * self.N))], axis=1)

    def launch_count(self):
        return int(self.L.nmpc_launch_count(self.h))

    def jac_pattern(self):
        cp, ri = np.zeros(self.n + 1, np.int32), np.zeros(self.nnz_jac, np.int32)
        check(self.L.nmpc_jac_pattern(self.h, cp.ctypes.data, ri.ctypes.data))
        return cp, ri

    def hess_pattern(self):
        cp, ri = np.zeros(self.n + 1, np.int32), np.zeros(self.nnz_hess, np.int32)
        check(self.L.nmpc_hess_pattern(self.h, cp.ctypes.data, ri.ctypes.data))
        return cp, ri

    # ---- host-buffer path (what a CasADi-style caller has) -------------------------------------
    def solve_host(self, x0, p, lbx, ubx, lbg, ubg, want=("f", "g", "lam_x", "lam_g", "stats"), out=None):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x0, p = np.atleast_2d(f64(x0)), np.atleast_2d(f64(p))
        lbx, ubx, lbg, ubg = f64(lbx), f64(ubx), f64(lbg), f64(ubg)
        B = x0.shape[0]
        if getattr(self, "_order", None) is not None and self._order.numel() != B:
            raise ValueError("the scheduling order set with set_order() has %d entries, the batch has %d" % (self._order.numel(), B))
        if x0.shape[1] != self.n or p.shape != (B, self.np_):
            raise ValueError("x0 must be [B,%d] and p [B,%d]" % (self.n, self.np_))
        batched = 1 if lbx.ndim == 2 else 0
        for a, m, nm in ((lbx, self.n, "lbx"), (ubx, self.n, "ubx"), (lbg, self.mg, "lbg"), (ubg, self.mg, "ubg")):
            if a.shape != ((B, m) if batched else (m,)):
                raise ValueError("%s has shape %s" % (nm, a.shape))
        o = out if out is not None else {}
        o.setdefault("x", np.empty((B, self.n)))
        for k, shp in (("f", (B,)), ("g", (B, self.mg)), ("lam_x", (B, self.n)), ("lam_g", (B, self.mg)), ("stats", (B, NSTATS))):
            if k in want:
                o.setdefault(k, np.empty(shp))
        o.setdefault("status", np.empty(B, np.int32))
        o.setdefault("iters", np.empty(B, np.int32))
        ptr = lambda k: o[k].ctypes.data if k in o else None
        check(self.L.nmpc_solve_host(self.h, B, x0.ctypes.data, p.ctypes.data, lbx.ctypes.data, ubx.ctypes.data,
                                     lbg.ctypes.data, ubg.ctypes.data, batched, ptr("x"), ptr("f"), ptr("g"),
                                     ptr("lam_x"), ptr("lam_g"), ptr("status"), ptr("iters"), ptr("stats")))
        return o

    # ---- device path: torch CUDA tensors, zero-copy, asynchronous on the current stream --------
    def _torch(self):
        import torch
        if not torch.cuda.is_available():
            raise NmpcError("no CUDA device: the batched path has no CPU fallback")
        return torch

    def workspace(self, B, batched_bounds=False, device=None):
        torch = self._torch()
        need = (self.L.nmpc_workspace_bytes_batched_bounds if batched_bounds else self.L.nmpc_workspace_bytes)(self.h, int(B))
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=device)
        return self._ws

    def _stream(self, device):
        """Raw stream handle of torch's current stream ON THE TENSORS' DEVICE (the handle checks the device itself)."""
        return self._torch().cuda.current_stream(device).cuda_stream

    def solve(self, x0, p, lbx, ubx, lbg, ubg, want=("f", "g", "lam_x", "lam_g", "stats"), out=None, trace=0):
        """x0 [B,n], p [B,6Nr], bounds [n]/[mg] or [B,.]: float64 CUDA tensors.  Returns CUDA tensors."""
        torch = self._torch()
        B = x0.shape[0]
        for t in (x0, p, lbx, ubx, lbg, ubg):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError("inputs must be contiguous float64 CUDA tensors")
        if getattr(self, "_order", None) is not None and self._order.numel() != B:
            raise ValueError("the scheduling order set with set_order() has %d entries, the batch has %d" % (self._order.numel(), B))
        batched = 1 if lbx.dim() == 2 else 0
        dev = x0.device
        if any(t.device != dev for t in (p, lbx, ubx, lbg, ubg)):
            raise ValueError("all inputs must live on the same CUDA device")
        ws = self.workspace(B, bool(batched), dev)
        o = out if out is not None else {}
        o.setdefault("x", torch.empty((B, self.n), dtype=torch.float64, device=dev))
        for k, shp in (("f", (B,)), ("g", (B, self.mg)), ("lam_x", (B, self.n)), ("lam_g", (B, self.mg)), ("stats", (B, NSTATS))):
            if k in want:
                o.setdefault(k, torch.empty(shp, dtype=torch.float64, device=dev))
        o.setdefault("status", torch.empty(B, dtype=torch.int32, device=dev))
        o.setdefault("iters", torch.empty(B, dtype=torch.int32, device=dev))
        ptr = lambda k: o[k].data_ptr() if k in o else None
        stream = self._stream(dev)
        args = [self.h, B, x0.data_ptr(), p.data_ptr(), lbx.data_ptr(), ubx.data_ptr(), lbg.data_ptr(), ubg.data_ptr(), batched,
                ptr("x"), ptr("f"), ptr("g"), ptr("lam_x"), ptr("lam_g"), ptr("status"), ptr("iters"), ptr("stats")]
        if trace:
            o["trace"] = torch.zeros((B, int(trace), NTRACE), dtype=torch.float64, device=dev)
            check(self.L.nmpc_solve_trace(*args, o["trace"].data_ptr(), int(trace), ws.data_ptr(), ws.numel(), stream))
        else:
            check(self.L.nmpc_solve(*args, ws.data_ptr(), ws.numel(), stream))
        return o

    def set_order(self, order=None):
        """Scheduling hint (nmpc_set_order): int32 CUDA tensor [B], a permutation of the instances, longest expected solve
        first; None restores index order.  The tensor is kept alive by the handle."""
        if order is not None:
            torch = self._torch()
            if not (order.is_cuda and order.dtype == torch.int32 and order.is_contiguous()):
                raise ValueError("order must be a contiguous int32 CUDA tensor")
        self._order = order
        check(self.L.nmpc_set_order(self.h, None if order is None else order.data_ptr(), 0 if order is None else int(order.numel())))

    def order_from_iters(self, iters):
        """Longest-first order from the iteration counts of the previous MPC step (the closed-loop predictor)."""
        torch = self._torch()
        return torch.argsort(iters, descending=True, stable=True).to(torch.int32).contiguous()

    def shift(self, x_prev, out=None):
        torch = self._torch()
        out = torch.empty_like(x_prev) if out is None else out
        check(self.L.nmpc_shift(self.h, x_prev.shape[0], x_prev.data_ptr(), out.data_ptr(), self._stream(x_prev.device)))
        return out

    def plant(self, state, x_opt, out=None):
        torch = self._torch()
        out = torch.empty_like(state) if out is None else out
        check(self.L.nmpc_plant(self.h, state.shape[0], state.data_ptr(), x_opt.data_ptr(), out.data_ptr(), self._stream(state.device)))
        return out

    def eval(self, w, p, lam_g=None, want=("f", "grad", "g", "jac", "hess")):
        torch = self._torch()
        B, dev = w.shape[0], w.device
        o = {}
        shapes = dict(f=(B,), grad=(B, self.n), g=(B, self.mg), jac=(B, self.nnz_jac), hess=(B, self.nnz_hess))
        for k in want:
            if k == "hess" and lam_g is None:
                continue
            o[k] = torch.empty(shapes[k], dtype=torch.float64, device=dev)
        ptr = lambda k: o[k].data_ptr() if k in o else None
        check(self.L.nmpc_eval(self.h, B, w.data_ptr(), p.data_ptr(), lam_g.data_ptr() if lam_g is not None else None,
                               ptr("f"), ptr("grad"), ptr("g"), ptr("jac"), ptr("hess"), self._stream(dev)))
        return o


def closed_loop(prob, P, lbx, ubx, lbg, ubg, steps, tol=1e-1, dmin=None):
    """Batched closed-loop driver, device resident: the reference's loop body
    (centralized_six_robots_implementation.py:416-465 with the Euler plant of casadi_test.py:17-26) for B
    independent instances at once.  Per step: solve -> apply u_0 through the plant -> reference shift as the next
    guess; instances whose ||x - xs|| <= tol stop moving (the loop guard at :416).  Returns a dict of CUDA tensors:
    traj [steps+1,B,3Nr], u [steps,B,2Nr], status [steps,B], iters [steps,B], active [steps,B], min_dist [B]."""
    torch = prob._torch()
    B, ns, nc, N = P.shape[0], prob.ns, prob.nc, prob.N
    p = P.clone()
    x0 = torch.cat([p[:, :ns].repeat(1, N + 1), torch.zeros((B, nc * N), dtype=torch.float64, device=P.device)], dim=1).contiguous()
    traj = torch.empty((steps + 1, B, ns), dtype=torch.float64, device=P.device)
    us = torch.zeros((steps, B, nc), dtype=torch.float64, device=P.device)
    st = torch.zeros((steps, B), dtype=torch.int32, device=P.device)
    its = torch.zeros((steps, B), dtype=torch.int32, device=P.device)
    act = torch.zeros((steps, B), dtype=torch.bool, device=P.device)
    traj[0] = p[:, :ns]
    out = {}
    for t in range(steps):
        if t == 0:
            prob.set_order(None)
        active = (p[:, :ns] - p[:, ns:]).norm(dim=1) > tol
        act[t] = active
        prob.solve(x0, p, lbx, ubx, lbg, ubg, want=(), out=out)
        u0 = out["x"][:, ns * (N + 1):ns * (N + 1) + nc]
        nxt = prob.plant(p[:, :ns].contiguous(), out["x"])
        p[:, :ns] = torch.where(active[:, None], nxt, p[:, :ns])
        us[t] = torch.where(active[:, None], u0, torch.zeros_like(u0))
        st[t], its[t] = out["status"], out["iters"]
        prob.set_order(prob.order_from_iters(out["iters"]))      # longest solves first in the next step
        x0 = prob.shift(out["x"])
        traj[t + 1] = p[:, :ns]
    prob.set_order(None)
    res = dict(traj=traj, u=us, status=st, iters=its, active=act)
    if prob.Nr > 1:
        pos = traj.reshape(steps + 1, B, prob.Nr, 3)[..., :2]
        d = (pos[:, :, :, None, :] - pos[:, :, None, :, :]).norm(dim=-1)
        d = d + torch.eye(prob.Nr, device=P.device, dtype=torch.float64)[None, None] * 1e9
        res["min_dist"] = d.amin(dim=(0, 2, 3))
    return res


class SmallOcp(Problem):
    """Small generic OCP, one GPU thread per instance (nmpc_create_ocp).  model 'van_der_pol' is the direct-multiple-shooting
    demo of mpc_pose_control_casadi.py:22-114: interleaved decision vector [X_0, U_0, ..., X_N] (n = 3N + 2), g = F(X_k, U_k) -
    X_{k+1} (mg = 2N), initial state fixed through its bounds, no parameter vector."""
    MODELS = {"van_der_pol": 1}

    def __init__(self, model="van_der_pol", N=20, T=10.0, rk_steps=4, **opts):
        self.L = lib()
        self.opts = default_opts(**({"max_iter": 3000} | opts))      # the demo sets no options: IPOPT's default max_iter
        self.h = C.c_void_p()
        check(self.L.nmpc_create_ocp(self.MODELS[model], int(N), float(T), int(rk_steps), C.byref(self.opts), C.byref(self.h)))
        self.model, self.Nr, self.N, self.T, self.rk_steps = model, 0, int(N), float(T), int(rk_steps)
        self.ns, self.nc, self.M, self.nobs, self.obstacles = 2, 1, 0, 0, None
        self.n, self.mg, self.np_ = self.L.nmpc_n(self.h), self.L.nmpc_mg(self.h), 0
        self.nnz_jac = self.nnz_hess = 0
        self._ws = None

    def demo_arrays(self):
        """w0, lbw, ubw, lbg, ubg exactly as the demo assembles them (mpc_pose_control_casadi.py:77-106)."""
        inf = np.inf
        w0, lbw, ubw = [0.0, 1.0], [0.0, 1.0], [0.0, 1.0]
        for _ in range(self.N):
            w0 += [0.0, 0.0, 0.0]; lbw += [-1.0, -0.25, -inf]; ubw += [1.0, inf, inf]
        return np.array(w0), np.array(lbw), np.array(ubw), np.zeros(self.mg), np.zeros(self.mg)

    def solve_host(self, x0, lbx, ubx, lbg, ubg, want=("f", "g", "lam_x", "lam_g", "stats"), out=None):
        x0 = np.atleast_2d(np.ascontiguousarray(x0, dtype=np.float64))
        return Problem.solve_host(self, x0, np.zeros((x0.shape[0], 0)), lbx, ubx, lbg, ubg, want=want, out=out)


class NlpSolver:
    """What nlpsol(...) returns: callable with the reference's keyword arguments (:432)."""

    def __init__(self, name, problem):
        self.name, self.problem = name, problem
        self._stats = {}

    def __call__(self, x0=0.0, p=None, lbx=-np.inf, ubx=np.inf, lbg=-np.inf, ubg=np.inf, lam_x0=None, lam_g0=None):
        P = self.problem
        if P.np_ == 0:      # parameter-free problem (the Van der Pol demo calls solver(x0=, lbx=, ubx=, lbg=, ubg=), :113)
            o = P.solve_host(_flat(x0, P.n, "x0")[None], _flat(lbx, P.n, "lbx"), _flat(ubx, P.n, "ubx"), _flat(lbg, P.mg, "lbg"),
                             _flat(ubg, P.mg, "ubg"))
        else:
            if p is None:
                raise ValueError("p = [x0; xs] is required")
            o = P.solve_host(_flat(x0, P.n, "x0")[None], _flat(p, P.np_, "p")[None], _flat(lbx, P.n, "lbx"),
                             _flat(ubx, P.n, "ubx"), _flat(lbg, P.mg, "lbg"), _flat(ubg, P.mg, "ubg"))
        st = int(o["status"][0])
        self._stats = dict(return_status=_cabi.STATUS.get(st, str(st)), success=st in (0, 1),
                           iter_count=int(o["iters"][0]), kkt_error=float(o["stats"][0, 0]))
        lam_p = np.zeros(P.np_)
        if P.np_:
            # dL/dp with L = f + lam_g'g: the initial-condition rows are X_0 - p[0:ns]; the cost holds xs = p[ns:] in
            # sum_{k<N} (X_k - xs)'Q(X_k - xs)  (centralized_six_robots_implementation.py:278,314)
            lam_p[:P.ns] = -o["lam_g"][0, :P.ns]
            Xk = o["x"][0, :P.ns * (P.N + 1)].reshape(P.N + 1, P.ns)[:P.N]
            Qd = np.tile(np.asarray(list(P.desc.Q), float), P.Nr)
            lam_p[P.ns:] = -2.0 * Qd * (Xk - np.asarray(_flat(p, P.np_, "p"))[P.ns:][None]).sum(axis=0)
        return {"x": DM(o["x"][0]), "f": DM(o["f"][:1]), "g": DM(o["g"][0]), "lam_x": DM(o["lam_x"][0]),
                "lam_g": DM(o["lam_g"][0]), "lam_p": DM(lam_p)}

    def stats(self):
        return dict(self._stats)


_IPOPT_KEYS = {f[0] for f in _cabi.Opts._fields_}


def nlpsol(name, plugin, nlp, opts=None):
    """Drop-in for casadi.nlpsol(name, 'ipopt', nlp_prob, opts) (:345-346).

    `nlp` is a problem descriptor instead of CasADi's symbolic dict:
        {'family': 'unicycle_centralized', 'Nr', 'N', 'T', 'Q': (3,), 'R': (2,)}
    `opts` uses the reference's layout: {'print_time': 0, 'ipopt': {'max_iter': 2000, ...}}.
    """
    if plugin not in ("ipopt", "b200ipm"):
        raise ValueError("plugin %r: this library implements the interior-point path only" % (plugin,))
    if isinstance(nlp, dict) and nlp.get("family") == "van_der_pol":
        ip = dict((opts or {}).get("ipopt", {}))
        kw = {k: v for k, v in ip.items() if k in _IPOPT_KEYS}
        return NlpSolver(name, SmallOcp("van_der_pol", nlp.get("N", 20), nlp.get("T", 10.0), nlp.get("rk_steps", 4), **kw))
    if not isinstance(nlp, dict) or nlp.get("family", "unicycle_centralized") not in ("unicycle_centralized", "unicycle_obstacles"):
        raise ValueError("nlp must be a descriptor {'family':'unicycle_centralized'|'unicycle_obstacles','Nr','N','T',...}")
    if nlp.get("family") == "unicycle_obstacles" and nlp.get("obstacles") is None:
        raise ValueError("family 'unicycle_obstacles' needs 'obstacles': [[x, y, clearance], ...]")
    ip = dict((opts or {}).get("ipopt", {}))
    ip.update({k: v for k, v in (opts or {}).items() if k in _IPOPT_KEYS})
    kw = {k: v for k, v in ip.items() if k in _IPOPT_KEYS}          # print_level etc. are accepted and ignored
    prob = Problem(nlp["Nr"], nlp["N"], nlp["T"], nlp.get("Q", (1.0, 5.0, 0.1)), nlp.get("R", (0.5, 0.05)),
                   obstacles=nlp.get("obstacles") if nlp.get("family") == "unicycle_obstacles" else None, **kw)
    return NlpSolver(name, prob)

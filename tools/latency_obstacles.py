"""Closed-loop single-instance latency of the static-obstacle family at the first scenario's own size (T = 0.1, N = 100)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
N, T, margin = int(sys.argv[1]) if len(sys.argv) > 1 else 100, 0.1, 0.05
obs = np.array([[0.4, 1.1, 0.30]])
prob = pkg.Problem(1, N, T, obstacles=obs)
lbx, ubx, lbg, ubg = prob.bounds_obstacles(margin, 0.2, np.pi / 4)
p = np.array([[0.0, 0.0, 0.0, 1.5, 1.5, 0.0]])
w = prob.cold_start(p[:, :3])
times, its, o = [], [], {}
for step in range(40):
    t0 = time.perf_counter()
    prob.solve_host(w, p, lbx, ubx, lbg, ubg, want=(), out=o)
    times.append(time.perf_counter() - t0); its.append(int(o["iters"][0]))
    x = o["x"][0]; X = x[:3 * (N + 1)].reshape(N + 1, 3); U = x[3 * (N + 1):].reshape(N, 2)
    th = p[0, 2]
    p[0, 0] += T * U[0, 0] * np.cos(th); p[0, 1] += T * U[0, 0] * np.sin(th); p[0, 2] += T * U[0, 1]
    w = np.concatenate([np.concatenate([X[1:], X[N - 1:N]]).ravel(), np.concatenate([U[1:], U[-1:]]).ravel()])[None]
print("N=%d: p50 %.2f ms  p95 %.2f ms  first(cold) %.1f ms  mean warm iters %.1f  status %s" % (N, 1e3 * np.median(times[1:]), 1e3 * np.percentile(times[1:], 95), 1e3 * times[0], np.mean(its[1:]), o["status"]))
import ctypes
prof = (ctypes.c_longlong * 16)()
prob.L.nmpc_debug_block_profile(prof, 1)
print("phase Mcycles [prepass, matvec, build, cholesky, trsm, syrk, forward, eval]:", [round(v / 1e6, 1) for v in prof[:8]])

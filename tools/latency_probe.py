import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
NR, NH, T = 6, 20, 0.3
prob = pkg.Problem(NR, NH, T)
lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
s3 = np.sqrt(3) / 2
st = np.array([[s3, 0.5, -2.618], [0, 1, -1.571], [-s3, 0.5, -0.524], [-s3, -0.5, 0.524], [0, -1, 1.571], [s3, -0.5, 2.618]])
st = st + 0.02 * np.sin(1.0 + 2.0 * np.arange(18)).reshape(6, 3)
goal = -st.copy(); goal[:, 2] = st[:, 2]
p1 = np.concatenate([st.ravel(), goal.ravel()])[None]
w1_ = prob.cold_start(p1[:, :18])
times, o1, its = [], {}, []
for step in range(60):
    t0 = time.perf_counter()
    prob.solve_host(w1_, p1, lbx, ubx, lbg, ubg, want=(), out=o1)
    times.append(time.perf_counter() - t0); its.append(int(o1["iters"][0]))
    u0 = o1["x"][0, 18 * (NH + 1):18 * (NH + 1) + 12]
    for i in range(6):
        th = p1[0, 3 * i + 2]
        p1[0, 3 * i] += T * u0[2 * i] * np.cos(th); p1[0, 3 * i + 1] += T * u0[2 * i] * np.sin(th); p1[0, 3 * i + 2] += T * u0[2 * i + 1]
    X = o1["x"][0, :18 * (NH + 1)].reshape(NH + 1, 18); U = o1["x"][0, 18 * (NH + 1):].reshape(NH, 12)
    w1_ = np.concatenate([np.concatenate([X[1:], X[NH - 1:NH]]).ravel(), np.concatenate([U[1:], U[-1:]]).ravel()])[None]
print("p50 %.3f ms  p95 %.3f ms  first %.1f ms  mean iters %.1f" % (1e3 * np.median(times[1:]), 1e3 * np.percentile(times[1:], 95), 1e3 * times[0], np.mean(its[1:])))

"""Batch throughput of the static-obstacle family (one robot, third scenario's six obstacles, N = 20) on both kernels."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
THIRD = [(-0.6, 3.3, 0.2), (0.6, 3.3, 0.125), (0.0, 2.3, 0.15), (1.0, 2.3, 0.15), (-0.6, 1.3, 0.2), (0.6, 1.3, 0.175)]
obs = np.array([[x, y, r + 0.15] for x, y, r in THIRD])
rng = np.random.default_rng(0)
P = np.tile(np.array([[0.0, 0.6, 1.57, 0.1, 3.9, 1.57]]), (B, 1)); P[:, :2] += 0.1 * rng.normal(size=(B, 2)); P[:, 3:5] += 0.2 * rng.normal(size=(B, 2))
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')
for label, tune in (("thread per instance", dict(thread_min_batch=1)), ("CTA per instance", dict(force_block_path=1)), ("warp per instance (default)", None)):
    prob = pkg.Problem(1, 20, 0.3, obstacles=obs, tuning=tune)
    lbx, ubx, lbg, ubg = prob.bounds_obstacles(0.1, 0.2, np.pi / 4)
    args = [t(prob.cold_start(P[:, :3])), t(P), t(lbx), t(ubx), t(lbg), t(ubg)]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time(); out = prob.solve(*args); torch.cuda.synchronize(); dt = time.time() - t0
    print("%-28s B=%d: %.3f s -> %.0f solves/s, solved %.3f, mean iters %.1f" % (label, B, dt, B / dt, (out["status"] == 0).double().mean().item(), out["iters"].double().mean().item()))

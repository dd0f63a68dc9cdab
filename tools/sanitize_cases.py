"""Small solves of every kernel family, for compute-sanitizer runs (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
from oracle.nlp_numpy import synthetic_instances
which = sys.argv[1] if len(sys.argv) > 1 else "all"
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')

def run(Nr, N, B, box, **kw):
    P = synthetic_instances(B, Nr=Nr, seed=5, box=box)
    prob = pkg.Problem(Nr, N, 0.3, **kw)
    lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
    out = prob.solve(t(prob.cold_start(P[:, :3 * Nr])), t(P), t(lbx), t(ubx), t(lbg), t(ubg))
    torch.cuda.synchronize()
    print(Nr, N, B, out["status"].cpu().numpy(), out["iters"].cpu().numpy())
    return prob, out

if which in ("all", "warp"):
    prob, out = run(2, 6, 5, 2.0, max_iter=12)
    w = prob.shift(out["x"]); s = prob.plant(t(np.zeros((5, 6))), out["x"]); torch.cuda.synchronize()
    run(6, 8, 5, 2.0, max_iter=8)
if which in ("all", "team"):
    run(8, 5, 2, 2.5, max_iter=6)
if which in ("all", "block"):
    run(12, 5, 2, 3.0, max_iter=6)
if which in ("all", "eval"):
    prob = pkg.Problem(6, 20, 0.3)
    g = torch.Generator(device="cuda").manual_seed(1)
    w = torch.randn((7, prob.n), dtype=torch.float64, device="cuda", generator=g)
    p = torch.randn((7, prob.np_), dtype=torch.float64, device="cuda", generator=g)
    lam = torch.randn((7, prob.mg), dtype=torch.float64, device="cuda", generator=g)
    o = prob.eval(w, p, lam); torch.cuda.synchronize(); print("eval ok", o["f"][:2].cpu().numpy())
if which in ("all", "obs"):
    prob = pkg.Problem(1, 8, 0.3, obstacles=[[0.45, 0.5, 0.3]], max_iter=10)
    lbx, ubx, lbg, ubg = prob.bounds_obstacles(0.05, 0.2, 0.78)
    P = np.array([[0.0, 0.0, 0.6, 1.2, 1.3, 0.0]])
    out = prob.solve(t(prob.cold_start(P[:, :3])), t(P), t(lbx), t(ubx), t(lbg), t(ubg)); torch.cuda.synchronize()
    print("obs", out["status"].cpu().numpy(), out["iters"].cpu().numpy())
if which in ("all", "ocp"):
    ocp = pkg.SmallOcp("van_der_pol", N=6, T=3.0, rk_steps=2, max_iter=10)
    w0, lbw, ubw, lbg, ubg = ocp.demo_arrays()
    out = ocp.solve_host(np.tile(w0, (3, 1)), lbw, ubw, lbg, ubg)
    print("ocp", out["status"], out["iters"])

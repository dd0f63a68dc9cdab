import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
from oracle.nlp_numpy import synthetic_instances
Nr, N, T = 64, 20, 0.3
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P = synthetic_instances(B, Nr=Nr, seed=20261018, box=8.0)
prob = pkg.Problem(Nr, N, T)
lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
x0 = prob.cold_start(P[:, :3 * Nr])
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')
args = [t(x0), t(P), t(lbx), t(ubx), t(lbg), t(ubg)]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = prob.solve(*args)
    torch.cuda.synchronize(); dt = time.time() - t0
    prof = (__import__("ctypes").c_longlong * 16)()
    prob.L.nmpc_debug_block_profile(prof, 1)
    print("phase Mcycles [prepass, matvec, build, chol-rest, trsm, syrk, forward, eval | chol: diag, rows, trailing | trsm: copy-in, V, main]:", [round(v / 1e6, 1) for v in prof[:14]])
    print("B=%d  %.3f s  status %s iters %s nfact %s kkt %.2e" % (B, dt, out['status'].cpu().numpy()[:8], out['iters'].cpu().numpy()[:8], out['stats'][:8, 8].cpu().numpy(), out['stats'][:, 0].max().item()), flush=True)

"""Throughput of a cold-start batch for any robot count:  python tools/time_batch.py Nr B [N] [box]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
from oracle.nlp_numpy import synthetic_instances
Nr, B = int(sys.argv[1]), int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 20
box = float(sys.argv[4]) if len(sys.argv) > 4 else max(2.0, 0.9 * np.sqrt(Nr))
P = synthetic_instances(min(B, 256), Nr=Nr, seed=20261018, box=box)
P = np.tile(P, ((B + len(P) - 1) // len(P), 1))[:B]
prob = pkg.Problem(Nr, N, 0.3)
lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
x0 = prob.cold_start(P[:, :3 * Nr])
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')
args = [t(x0), t(P), t(lbx), t(ubx), t(lbg), t(ubg)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    out = prob.solve(*args)
    torch.cuda.synchronize(); dt = time.time() - t0
print("Nr=%d N=%d B=%d: %.3f s -> %.1f solves/s; solved %.3f; mean iters %.1f; mean factorisations %.1f" % (
    Nr, N, B, dt, B / dt, (out['status'] == 0).double().mean().item(), out['iters'].double().mean().item(), out['stats'][:, 8].mean().item()))

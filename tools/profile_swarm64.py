"""One short 64-robot solve (iteration cap from argv) for ncu captures of solve_kernel_block."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
from oracle.nlp_numpy import synthetic_instances
Nr, N, T = 64, 20, 0.3
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = synthetic_instances(1, Nr=Nr, seed=20261018, box=8.0)
prob = pkg.Problem(Nr, N, T, max_iter=iters)
lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
x0 = prob.cold_start(P[:, :3 * Nr])
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')
out = prob.solve(t(x0), t(P), t(lbx), t(ubx), t(lbg), t(ubg))
torch.cuda.synchronize()
print(out['status'].cpu().numpy(), out['iters'].cpu().numpy(), out['stats'][0, 8].item())

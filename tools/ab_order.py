"""Does a cold-start ordering key shorten the tail?  Index order vs longest-predicted-first by the objective of the initial guess
(N x the Q-weighted squared start-goal error), the benchmark workload:   python tools/ab_order.py [B]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
P = pkg.workload.synthetic_instances(B, 6)
prob = pkg.Problem(6, 20, 0.3)
lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')
args = [t(prob.cold_start(P[:, :18])), t(P), t(lbx), t(ubx), t(lbg), t(ubg)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda:0')
out = {}
def timed(n=3):
    ms = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.solve(*args, want=("stats",), out=out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return B / np.mean(ms) * 1e3
timed(2)
d = (args[1][:, 18:] - args[1][:, :18]).reshape(B, 6, 3)
f0 = (d[..., 0] ** 2 + 5.0 * d[..., 1] ** 2 + 0.1 * d[..., 2] ** 2).sum(1)
orders = {"index": None, "by f0": torch.argsort(f0, descending=True).to(torch.int32).contiguous(),
          "by true factorisations": torch.argsort(out["stats"][:, 8], descending=True).to(torch.int32).contiguous(),
          "reverse f0 (worst case)": torch.argsort(f0, descending=False).to(torch.int32).contiguous()}
for rep in range(2):
    for name, o in orders.items():
        prob.set_order(o)
        print("%-26s %.0f solves/s" % (name, timed()))
prob.set_order(None)

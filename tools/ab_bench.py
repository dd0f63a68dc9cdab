"""Quick A/B timing of the benchmark workload (6 robots, N = 20, cold start; bench.py's instances) on whatever
lib/libnmpc_b200.so is in the tree:   python tools/ab_bench.py [B] [convoy] [ctas_per_sm] [label]
Prints one line: cold solves/s (CUDA events, best and mean of 3, L2 flushed), warm solves/s, iterations."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
from oracle.nlp_numpy import synthetic_instances
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
tune = {}
if len(sys.argv) > 2 and int(sys.argv[2]) >= 0: tune["convoy"] = int(sys.argv[2])
if len(sys.argv) > 3 and int(sys.argv[3]) > 0: tune["ctas_per_sm"] = int(sys.argv[3])
label = sys.argv[4] if len(sys.argv) > 4 else ""
P = synthetic_instances(B, 6, 20261018)
prob = pkg.Problem(6, 20, 0.3, tuning=tune)
lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')
args = [t(prob.cold_start(P[:, :18])), t(P), t(lbx), t(ubx), t(lbg), t(ubg)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda:0')
out = {}
def timed(a, n=3):
    ms = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.solve(*a, want=("stats",), out=out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms
timed(args, 2)
ms = timed(args)
it, nf, ok = out["iters"].double().mean().item(), out["stats"][:, 8].mean().item(), (out["status"] == 0).double().mean().item()
xo = out["x"].clone()
p2 = args[1].clone(); p2[:, :18] = prob.plant(args[1][:, :18].contiguous(), xo)
wargs = [prob.shift(xo), p2] + args[2:]
prob.set_order(prob.order_from_iters(out["iters"]))
timed(wargs, 1)
wms = timed(wargs)
print("%-28s B=%d cold best %.0f mean %.0f solves/s (iters %.1f fact %.1f solved %.4f) | warm %.0f solves/s (iters %.1f)" % (
    label, B, B / min(ms) * 1e3, B / np.mean(ms) * 1e3, it, nf, ok, B / np.mean(wms) * 1e3, out["iters"].double().mean().item()))

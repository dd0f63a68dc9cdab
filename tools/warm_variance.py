import sys, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
pkg = ge.load_package()
from oracle.nlp_numpy import synthetic_instances
Nr, N, T, B = 6, 20, 0.3, 8192
P = synthetic_instances(B, Nr=Nr, seed=20261018)
prob = pkg.Problem(Nr, N, T)
lbx, ubx, lbg, ubg = prob.bounds(0.3, 0.22, 2.84)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda:0')
d = [t(prob.cold_start(P[:, :18])), t(P), t(lbx), t(ubx), t(lbg), t(ubg)]
out = prob.solve(*d)
p2 = d[1].clone(); p2[:, :18] = prob.plant(d[1][:, :18].contiguous(), out["x"]); x0w = prob.shift(out["x"])
ow = prob.solve(x0w, p2, *d[2:])
torch.cuda.synchronize()
it = ow["iters"].cpu().numpy(); nf = ow["stats"][:, 8].cpu().numpy()
print("warm iters: mean %.1f max %d; factorisations mean %.1f max %d; top10 iters %s; top10 nfact %s" % (it.mean(), it.max(), nf.mean(), nf.max(), np.sort(it)[-10:], np.sort(nf)[-10:]))
def timeit(label):
    ms = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.solve(x0w, p2, *d[2:], want=("stats",)); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(label, ["%.1f" % m for m in ms], "-> %.0f solves/s (mean)" % (B / (np.mean(ms) * 1e-3)))
timeit("index order        ")
prob.set_order(prob.order_from_iters(out["iters"])); timeit("LPT by cold iters  ")
prob.set_order(prob.order_from_iters(ow["iters"])); timeit("LPT by warm iters  ")
prob.set_order(prob.order_from_iters(ow["stats"][:, 8].to(torch.int32))); timeit("LPT by warm nfact  ")

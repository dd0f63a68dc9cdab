#!/usr/bin/env python
"""Roofline of the streaming kernels either side of the solver (K1 eval, K4 shift, plant) on one B200.

Algorithmic bytes: eval reads w + lam_g + p and writes f, grad, g, jac, hess = 52,136 B per point at
Nr=6, N=20; shift reads and writes w (2 x 4,944 B).  Timed with CUDA events, inputs larger than L2 (B = 65,536 ->
3.4 GB of eval traffic), 3 warm-ups, best-of-5 reported next to the mean."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), float(np.mean(ts))


def main():
    pkg = ge.load_package()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    prob = pkg.Problem(6, 20, 0.3)
    g = torch.Generator(device="cuda").manual_seed(1)
    w = torch.randn((B, prob.n), dtype=torch.float64, device="cuda", generator=g)
    p = torch.randn((B, prob.np_), dtype=torch.float64, device="cuda", generator=g)
    lam = torch.randn((B, prob.mg), dtype=torch.float64, device="cuda", generator=g)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    out = {}
    eval_bytes = 8 * (prob.n + prob.mg + prob.np_ + 1 + prob.n + prob.mg + prob.nnz_jac + prob.nnz_hess)
    # outputs are allocated once (the timed region is the kernel, not the allocator)
    o = prob.eval(w, p, lam)
    L = pkg.lib()
    st = torch.cuda.current_stream().cuda_stream

    def ev():
        pkg._cabi.check(L.nmpc_eval(prob.h, B, w.data_ptr(), p.data_ptr(), lam.data_ptr(), o["f"].data_ptr(), o["grad"].data_ptr(),
                                    o["g"].data_ptr(), o["jac"].data_ptr(), o["hess"].data_ptr(), st))
    best, mean = timeit(ev)
    out["eval"] = dict(bytes_per_point=eval_bytes, B=B, best_s=best, mean_s=mean, gbs=eval_bytes * B / best / 1e9, frac=eval_bytes * B / best / 1e9 / peak,
                       evals_per_s=B / best)
    xn = torch.empty_like(w)
    best, mean = timeit(lambda: prob.shift(w, xn))
    out["shift"] = dict(bytes_per_point=16 * prob.n, B=B, best_s=best, mean_s=mean, gbs=16 * prob.n * B / best / 1e9, frac=16 * prob.n * B / best / 1e9 / peak)
    s = torch.randn((B, prob.ns), dtype=torch.float64, device="cuda", generator=g)
    so = torch.empty_like(s)
    best, mean = timeit(lambda: prob.plant(s, w, so))
    out["plant"] = dict(bytes_per_point=8 * (2 * prob.ns + prob.nc), B=B, best_s=best, mean_s=mean)
    out["hbm_peak_gbs"] = peak
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Generates tests/golden/swarm64_oracle.npz: the CPU oracle's solution of the two 64-robot, N = 20 instances used by
tests/test_gpu_block.py::test_block_path_64_robot_swarm (BASELINE.json configs[4]; SURVEY.md 8d recipe: starts and
goals uniform in [-8, 8]^2 with separation >= 0.5, seed 20261018).  The oracle needs several minutes per instance at
this size, so its outputs are committed as a fixture.  Run from the repo root:  python tests/golden/make_swarm64_golden.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.nlp_numpy import synthetic_instances  # noqa: E402
from oracle.oracle_lib import Oracle  # noqa: E402

Nr, N, T, B = 64, 20, 0.3, 2
P = synthetic_instances(B, Nr=Nr, seed=20261018, box=8.0)
orc = Oracle(Nr, N, T)
lbx, ubx, lbg, ubg = orc.bounds(0.3, 0.22, 2.84)
x0 = np.stack([orc.cold_start(P[b, :3 * Nr]) for b in range(B)])
t0 = time.time()
ref = orc.solve_batch(x0, P, lbx, ubx, lbg, ubg)
print("oracle: %.1f s, status %s, iters %s, f %s" % (time.time() - t0, ref["status"], ref["iters"], ref["f"]))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "swarm64_oracle.npz"), P=P, x=ref["x"], f=ref["f"],
                    status=ref["status"], iters=ref["iters"])

"""Synthetic workloads and scenario fixtures of the benchmark (host code, NumPy only).

* ``synthetic_instances``: the BASELINE.md / SURVEY.md 8d recipe -- B start/goal sets for Nr robots, drawn from ONE
  ``np.random.default_rng(seed)`` stream (seed 20261018): per instance, rejection-sample Nr start positions uniform in
  [-box, box]^2 with all pairwise distances >= sep, headings uniform in [-pi, pi); goals the same way, independently.
  Instance b of the table does not depend on B (prefix property), so ranks can take slices of the 65,536-instance table.
* ``hexagon_swap``: the six-robot antipodal swap of sixth_scenario.py:291-292,308-310 (latency fixture), optionally
  de-symmetrised (the exactly symmetric layout has mirror-image optima between which rounding decides).
"""
from __future__ import annotations

import numpy as np

BENCH_SEED = 20261018
BENCH_TABLE = 65536      # instances in the full table (8 GPUs x 8,192)


def synthetic_instances(B, Nr=6, seed=BENCH_SEED, box=2.0, sep=0.5):
    rng = np.random.default_rng(seed)

    def draw():
        while True:
            xy = rng.uniform(-box, box, (Nr, 2))
            d = np.linalg.norm(xy[:, None] - xy[None], axis=-1) + np.eye(Nr) * 1e9
            if d.min() >= sep:
                th = rng.uniform(-np.pi, np.pi, (Nr, 1))
                return np.concatenate([xy, th], axis=1).reshape(-1)

    P = np.empty((B, 6 * Nr))
    for b in range(B):
        P[b, : 3 * Nr] = draw()
        P[b, 3 * Nr:] = draw()
    return P


def bench_shard(rank, world, per_gpu=None, Nr=6, seed=BENCH_SEED):
    """Rank `rank`'s slice of the benchmark table (sharding.shard_range over world * per_gpu instances, 65,536 by
    default at world = 8): every rank draws the same stream and keeps its own contiguous part."""
    from .sharding import shard_range
    per_gpu = BENCH_TABLE // 8 if per_gpu is None else int(per_gpu)
    lo, hi = shard_range(world * per_gpu, rank, world)
    return synthetic_instances(hi, Nr, seed)[lo:hi], (lo, hi)


def hexagon_swap(desym=True):
    s3 = np.sqrt(3) / 2
    st = np.array([[s3, 0.5, -2.618], [0, 1, -1.571], [-s3, 0.5, -0.524], [-s3, -0.5, 0.524], [0, -1, 1.571], [s3, -0.5, 2.618]])
    if desym:
        st = st + 0.02 * np.sin(1.0 + 2.0 * np.arange(18)).reshape(6, 3)
    goal = -st.copy(); goal[:, 2] = st[:, 2]
    return np.concatenate([st.ravel(), goal.ravel()])

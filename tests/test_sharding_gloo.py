"""world_size-2 gloo test of the N>1 host logic: contiguous sharding, MAX-over-ranks timing, final gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from oracle.nlp_numpy import synthetic_instances
    from oracle.oracle_lib import Oracle
    B = 7
    lo, hi = pkg.sharding.shard_range(B, rank, world)
    P = synthetic_instances(B, 2, seed=5)[lo:hi]          # every rank derives the same instance table, takes its slice
    o = Oracle(2, 6, 0.1)
    lbx, ubx, lbg, ubg = o.bounds(0.25, 0.22, 2.84)
    x0 = np.stack([o.cold_start(p[:6]) for p in P])
    r = o.solve_batch(x0, P, lbx, ubx, lbg, ubg, nthreads=1)   # the CPU oracle stands in for the GPU solve
    u0 = torch.tensor(r["x"][:, 6 * 7:6 * 7 + 4])
    allu = pkg.sharding.gather_first_controls(u0, dist)
    val, tmax, n = pkg.sharding.job_throughput(hi - lo, 0.5 + rank, dist)
    q.put((rank, lo, hi, allu.numpy(), val, tmax, n, r["status"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, u_a, v0, t0, n0, s0), (r1, lo1, hi1, u_b, v1, t1, n1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 7)                     # contiguous, balanced, complete
    assert n0 == n1 == 7 and t0 == t1 == 1.5 and v0 == v1 == 7 / 1.5   # sum of units / max of time
    np.testing.assert_array_equal(u_a, u_b)                          # every rank holds the gathered result
    assert u_a.shape == (7, 4) and s0 == [0] * 4 and s1 == [0] * 3
    # the gathered controls equal a single-process solve of the whole batch
    sys.path.insert(0, ROOT)
    from oracle.nlp_numpy import synthetic_instances
    from oracle.oracle_lib import Oracle
    P = synthetic_instances(7, 2, seed=5)
    o = Oracle(2, 6, 0.1)
    lbx, ubx, lbg, ubg = o.bounds(0.25, 0.22, 2.84)
    r = o.solve_batch(np.stack([o.cold_start(p[:6]) for p in P]), P, lbx, ubx, lbg, ubg, nthreads=1)
    np.testing.assert_array_equal(u_a, r["x"][:, 42:46])


def test_shard_range_properties(pkg):
    for B in (1, 7, 8192, 65536):
        for w in (1, 2, 3, 8):
            edges = [pkg.sharding.shard_range(B, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_benchmark_table_is_one_stream_and_shards_tile_it(pkg):
    """ONE default_rng(20261018) table; the package's generator (used by bench.py) and the oracle's copy (used by
    the tests) are the same function, instance b does not depend on the table size, and the ranks' shards tile the table."""
    import numpy as np
    from oracle.nlp_numpy import synthetic_instances as oracle_gen
    wl = pkg.workload
    full = wl.synthetic_instances(96, 6)
    np.testing.assert_array_equal(full, oracle_gen(96, 6, 20261018))
    np.testing.assert_array_equal(full[:40], wl.synthetic_instances(40, 6))
    parts = [wl.bench_shard(r, 4, per_gpu=24)[0] for r in range(4)]
    np.testing.assert_array_equal(np.concatenate(parts), full)
    assert wl.bench_shard(2, 4, per_gpu=24)[1] == (48, 72)

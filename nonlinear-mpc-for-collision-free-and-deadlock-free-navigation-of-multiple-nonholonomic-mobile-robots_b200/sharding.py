"""Host-side sharding of independent NMPC instances over ranks (one process per GPU).

Instances (scenarios, seeds, start/goal sets) are independent NLPs, so the path shards with no
data-path collective: rank r of G takes a contiguous slice; the only collectives are
the timing reduction (MAX over ranks) and an optional gather of first controls / statuses."""
from __future__ import annotations


def shard_range(B, rank, world):
    """Contiguous, balanced split: the first B % world ranks get one extra instance."""
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def job_throughput(n_local, seconds_local, dist=None, device=None):
    """Whole-job units/s = sum of units over ranks / max of time over ranks."""
    import torch
    t = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
    n = torch.tensor([float(n_local)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return n.item() / t.item(), t.item(), n.item()


def gather_first_controls(u0_local, dist=None):
    """Optional final gather of u_0 [B_local, 2Nr] onto every rank (the one exchange a consumer may need)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return u0_local
    sizes = [torch.zeros(1, dtype=torch.int64, device=u0_local.device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([u0_local.shape[0]], dtype=torch.int64, device=u0_local.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx, u0_local.shape[1]), dtype=u0_local.dtype, device=u0_local.device)
    pad[:u0_local.shape[0]] = u0_local
    out = [torch.zeros_like(pad) for _ in sizes]
    dist.all_gather(out, pad)
    return torch.cat([o[:int(s.item())] for o, s in zip(out, sizes)], dim=0)

"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the box, gloo in CPU tests).

* rows of (X, y) are split per rank; each evaluation needs ONE all-reduce of the pass-1 partial [m*m + m + 3] and one of the
  pass-2 gradient partial [d+2+m*d]; every rank runs the m x m section redundantly (identical inputs => identical P, u).
* HMC chains are split across ranks; there is no collective in the sampling loop, only the final trace gather.
"""
import torch
import torch.distributed as dist


def shard_rows(n, rank, world):
    """Contiguous row span [lo, hi) of rank `rank`; sizes differ by at most one row."""
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_chains(n_chains, rank, world):
    lo, hi = shard_rows(n_chains, rank, world)
    return list(range(lo, hi))


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_sum_(t, group=None):
    if is_distributed() or group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_chains(local, n_chains, group=None):
    """Concatenate per-rank chain results (first dim = local chains) in rank order on every rank."""
    if not is_distributed():
        return local
    world = dist.get_world_size(group)
    sizes = [shard_rows(n_chains, r, world) for r in range(world)]
    mx = max(b - a for a, b in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:b - a] for o, (a, b) in zip(outs, sizes)], dim=0)

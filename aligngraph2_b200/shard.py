"""Read sharding and timing reduction for the multi-GPU runs.

The mecat2ref+ path shards by reads with no data-path collective: every rank holds the whole reference and a
contiguous range of the reads (concatenating the ranks' records in rank order keeps the file order).  The only
communication is the bookkeeping of a measurement: a barrier, the max over ranks of the elapsed time and the sum of
the work counters.  Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def read_range(n_reads: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of [0, n_reads): the first n_reads % world ranks get one read more."""
    base, extra = divmod(n_reads, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_measurement(elapsed_ms: float, counters: dict, device="cpu") -> tuple[float, dict]:
    """(max over ranks of elapsed_ms, sum over ranks of every counter).  No-op outside torch.distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return elapsed_ms, dict(counters)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    keys = sorted(counters)
    c = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), {k: float(v) for k, v in zip(keys, c.tolist())}

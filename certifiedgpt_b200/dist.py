"""Sharding of the Monte-Carlo draws across ranks: one process per GPU, the N draws
of a `_sample_noise` call are split into contiguous slices of the GLOBAL sample range, noise is keyed
by global sample index, and only the int64 label-count vector is all-reduced."""
import torch


def shard_range(base: int, num: int, rank: int, world: int):
    """Rank `rank`'s slice [lo, hi) of the global sample range [base, base + num)."""
    assert 0 <= rank < world and num >= 0
    return base + (num * rank) // world, base + (num * (rank + 1)) // world


def rank_world(process_group):
    """process_group: None (single process), True (default group) or a torch ProcessGroup."""
    if process_group is None:
        return 0, 1
    import torch.distributed as dist
    g = None if process_group is True else process_group
    return dist.get_rank(g), dist.get_world_size(g)


def allreduce_counts(counts: torch.Tensor, process_group):
    """Sum of the per-rank count vectors (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    if process_group is None:
        return counts
    import torch.distributed as dist
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=None if process_group is True else process_group)
    return counts

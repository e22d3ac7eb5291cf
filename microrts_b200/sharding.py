"""Multi-GPU plumbing: games are independent, so a run shards by contiguous game-index range with one batch per
GPU/rank and no data-path collective.  The only collective is one all-reduce (sum) of the win/score counters.

Game g of the global batch always uses seed `seed0 + g`, so results do not depend on the number of ranks.
"""
import numpy as np

STAT_KEYS = ["wins_p0", "wins_p1", "draws", "games_finished", "cycles", "decisions", "unit_cycles", "errors"]


def shard(n_total, rank, world):
    """Contiguous range [first, first+count) of rank's games; remainders go to the lowest ranks."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def global_seeds(seed0, first, count):
    return np.arange(first, first + count, dtype=np.int64) + np.int64(seed0)


def reduce_stats(stats, device=None, group=None):
    """Sum the per-rank counters of BatchedGameState.stats() over all ranks (NCCL for CUDA tensors, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(stats[k]) for k in STAT_KEYS], dtype=torch.int64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_KEYS, t.tolist()))

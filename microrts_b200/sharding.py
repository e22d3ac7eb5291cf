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


class Communicator:
    """The run's one NCCL communicator, created and used behind the C ABI (mrts_nccl_unique_id / mrts_nccl_comm_create /
    mrts_batch_stats_allreduce, include/microrts_cuda.h).  `exchange` is the host's own plumbing for handing rank 0's 128-byte
    unique id to the other ranks: a callable taking the id (bytes) on rank 0 and None elsewhere, returning the id on every rank
    (bench.py passes a torch.distributed object broadcast; a Java host would pass an array between its per-GPU threads)."""

    def __init__(self, rank, world, device, exchange):
        import ctypes as C
        from . import _ffi
        from .api import _check
        L = _ffi.lib()
        uid = None
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            _check(L.mrts_nccl_unique_id(buf))
            uid = bytes(buf)
        uid = exchange(uid)
        assert isinstance(uid, (bytes, bytearray)) and len(uid) == 128
        h = C.c_void_p()
        _check(L.mrts_nccl_comm_create((C.c_uint8 * 128).from_buffer_copy(uid), world, rank, device, C.byref(h)))
        self._h, self.rank, self.world = h, rank, world

    def all_reduce_stats(self, batch):
        """Counters of `batch` summed over every rank's batch (one ncclAllReduce of 8 int64 inside the library)."""
        from . import _ffi
        from .api import _check
        out = np.zeros(8, dtype=np.int64)
        _check(_ffi.lib().mrts_batch_stats_allreduce(batch._h, out.ctypes.data, self._h))
        return dict(zip(STAT_KEYS, out.tolist()))

    def close(self):
        if self._h is not None:
            from . import _ffi
            _ffi.lib().mrts_nccl_comm_destroy(self._h)
            self._h = None


def reduce_stats_host(stats, group=None):
    """CPU-side sum of per-rank counter dicts over a gloo group: the world_size > 1 host-logic tests on machines without a GPU
    (tests/test_multi_rank.py).  GPU runs use Communicator.all_reduce_stats, which reduces inside libmicrorts_cuda.so."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(stats[k]) for k in STAT_KEYS], dtype=torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        assert dist.get_backend(group) == "gloo", "GPU runs reduce through Communicator (the C ABI), not through torch.distributed"
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_KEYS, t.tolist()))

"""Multi-GPU layout of the env batch: one process per GPU, every rank owns a contiguous range of
GLOBAL env ids, no collective on the step path.  Philox streams are keyed by the
global id (SsdConfig.env_id_offset), so a trajectory does not depend on how the batch is cut.
The only exchange is the end-of-run reduction of the stats counters (ssd_stats)."""
import os

from ._names import STAT_NAMES


def dist_env():
    """(rank, world_size, local_rank) as torchrun exports them; (0, 1, 0) for a plain launch."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total_envs, world_size, rank):
    """[begin, end) of the global env ids owned by `rank` when `total_envs` are split as evenly as
    possible: the first total_envs % world_size ranks hold one env more."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside 0..%d" % (rank, world_size - 1))
    if total_envs < 0:
        raise ValueError("total_envs must be non-negative")
    q, r = divmod(int(total_envs), int(world_size))
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def weak_range(envs_per_gpu, rank):
    """Weak scaling (bench.py): every rank owns envs_per_gpu envs of its own."""
    return rank * int(envs_per_gpu), (rank + 1) * int(envs_per_gpu)


def make_shard(cfg, total_envs, device, seed=0, rank=None, world_size=None, **kw):
    """BatchedSSDEnv over this rank's slice of a `total_envs` batch."""
    from .batched import BatchedSSDEnv
    r, w, _ = dist_env()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    begin, end = shard_range(total_envs, world_size, rank)
    if end == begin:
        raise ValueError("rank %d owns no environments (%d envs over %d ranks)" % (rank, total_envs, world_size))
    return BatchedSSDEnv(cfg, end - begin, device=device, seed=seed, env_id_offset=begin, **kw)


def reduce_stats(stats, device=None, group=None):
    """Sum the per-rank stats dicts (BatchedSSDEnv.stats()) over all ranks: the one collective of a
    run (NCCL all-reduce of 8 int64 on GPU jobs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    vec = torch.tensor([int(stats[k]) for k in STAT_NAMES], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_NAMES, (int(x) for x in vec.tolist())))


def max_over_ranks(value, device=None, group=None):
    """Max of a float over ranks (device timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

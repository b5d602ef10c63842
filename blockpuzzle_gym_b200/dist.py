"""Multi-GPU plumbing: env sharding by global index and the single statistics all-reduce.

Envs are independent (the reference keeps them in a Python list, rollout.py:33, with seeds
seed + 1000*idx, rollout.py:206-210), so rank r of R owns the contiguous global env range
shard_range(total, r, R) and passes its start as env_index_offset; Philox is keyed by the global index,
so results do not depend on R.  The only collective is one all-reduce (sum) of the float64[8]
statistics vector, replacing mpi_moments (train.py:21-26,67-73).
"""
import torch
import torch.distributed as dist

from ._lib import STAT_NAMES


def shard_range(total_envs, rank, world):
    """Contiguous [start, stop) of global env indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total_envs), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_stats(stats, group=None):
    """Sum the statistics vector over all ranks in place (NCCL on GPU tensors, gloo on CPU tensors)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def stats_dict(stats):
    v = stats.detach().cpu().tolist()
    d = {k: v[i] for i, k in enumerate(STAT_NAMES)}
    d["success_rate"] = d["successes"] / d["episodes"] if d["episodes"] else float("nan")  # rollout.py:163-167
    return d

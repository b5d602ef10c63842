"""World-size-2 gloo test (CPU) of the multi-GPU host logic: env sharding by global index and the one
statistics all-reduce.  The per-rank env work is done by the oracle here (no GPU in this container)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blockpuzzle_gym_b200.dist import allreduce_stats, shard_range, stats_dict


def test_shard_ranges_partition_the_envs():
    for total in (1, 7, 8, 4096, 8388608, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(8388608, 3, 8) == (3 * 1048576, 4 * 1048576)


def _worker(rank, world, port, total, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import coracle
    lo, hi = shard_range(total, rank, world)
    env = coracle.OracleVecEnv("BlocksTouch-v0", hi - lo, seed=5, env_index_offset=lo)
    env.reset()
    st = env.run_random(steps)                      # [episodes, successes, steps]
    stats = torch.zeros(8, dtype=torch.float64)
    stats[0], stats[1], stats[2] = float(st[0]), float(st[1]), float(st[2])
    allreduce_stats(stats)
    state = env.get_state()
    q.put((rank, stats.numpy().copy(), state.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce():
    total, steps, world = 48, 100, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import coracle
    full = coracle.OracleVecEnv("BlocksTouch-v0", total, seed=5)
    full.reset()
    st = full.run_random(steps)
    for rank, stats, _ in res:                      # every rank holds the global sums
        assert stats[0] == st[0] and stats[1] == st[1] and stats[2] == st[2] == total * steps
    assert res[0][2] + res[1][2] == full.get_state().tobytes()   # sharded == unsharded, bit for bit
    d = stats_dict(torch.from_numpy(res[0][1]))
    assert d["episodes"] == total * (steps // 50) and 0.0 <= d["success_rate"] <= 1.0

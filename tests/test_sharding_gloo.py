"""CPU, world_size 2 (gloo): the N>1 host logic — shard layout and the statistics all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from gym_po.sharding import allreduce_stats, rank_world, shard_envs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert rank_world() == (rank, world)
        n, off = shard_envs(total, rank, world)
        # every rank learns every shard and checks that they tile [0, total) exactly
        mine = torch.tensor([n, off], dtype=torch.int64)
        allv = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allv, mine)
        pos = 0
        for nn, oo in (v.tolist() for v in allv):
            assert oo == pos and oo % 512 == 0
            pos += nn
        assert pos == total
        # per-rank episode statistics -> whole-job statistics
        stats = torch.tensor([10.0 * (rank + 1), 5.0 * (rank + 1), 100.0 * (rank + 1), 3.0, float(n), 0, 0, 0], dtype=torch.float64)
        out = allreduce_stats(stats)
        tri = world * (world + 1) / 2
        assert out["episodes"] == 10.0 * tri and out["sum_return"] == 5.0 * tri and out["env_steps"] == float(total)
        assert abs(out["mean_return"] - 0.5) < 1e-12 and abs(out["mean_length"] - 10.0) < 1e-12
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [1 << 16, 1000, 70_001])
def test_two_rank_sharding_and_stats_allreduce(total):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total), nprocs=2, join=True)


def test_shard_envs_properties():
    from gym_po.sharding import shard_envs
    for total in (1, 511, 512, 513, 4096, 1 << 22, (1 << 24) + 7):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                n, off = shard_envs(total, r, world)
                assert off == pos or n == 0
                assert off % 512 == 0
                pos += n
            assert pos == total
    with pytest.raises(ValueError):
        shard_envs(10, 2, 2)


def test_numa_binding_is_best_effort():
    """bind_to_gpu_numa_node never raises: without NVML / sysfs information it reports why nothing was done."""
    import os
    from gym_po.sharding import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    msg = bind_to_gpu_numa_node(0)
    assert isinstance(msg, str) and msg
    assert os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)

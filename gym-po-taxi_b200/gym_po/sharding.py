"""Multi-GPU plumbing: one process per GPU, the env batch split contiguously, no per-step collective.

The vectorized step never reads another env's state (reference extended_taxi.py:244-287,
rooms/rooms.py:198-222: elementwise ops and gathers from static tables only), so the batch axis shards
trivially.  Rank g owns global envs ``[offset_g, offset_g + n_g)``; Philox streams are keyed by the GLOBAL
env id (``env_offset`` in the C ABI), so trajectories do not depend on the number of GPUs.  The only
collective is a sum all-reduce of the 8-double episode-statistics vector at logging cadence
(``torch.distributed`` / NCCL over NVLink; gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from ._native import ENV_ALIGN

STAT_FIELDS = ("episodes", "sum_return", "sum_length", "sum_return_sq", "env_steps", "reserved0", "reserved1", "reserved2")


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_envs(total_envs: int, rank: int, world: int, align: int = ENV_ALIGN):
    """Contiguous split of ``total_envs`` into ``world`` shards whose offsets are multiples of ``align``
    (the kernels' tile size).  Returns ``(num_envs, env_offset)`` for ``rank``; shards differ by at most
    one tile and the last one takes the ragged remainder."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    tiles = -(-total_envs // align)
    base, extra = divmod(tiles, world)
    my_tiles = base + (1 if rank < extra else 0)
    first_tile = rank * base + min(rank, extra)
    offset = first_tile * align
    n = max(0, min(total_envs, offset + my_tiles * align) - offset)
    return n, offset


def allreduce_stats(stats: torch.Tensor) -> dict:
    """Sum the per-rank statistics vector over all ranks (in place) and return it as a dict with the
    derived means.  A no-op reduction when torch.distributed is not initialised (single GPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    v = stats.detach().cpu().tolist()
    out = dict(zip(STAT_FIELDS, v))
    n = max(out["episodes"], 1.0)
    out["mean_return"] = out["sum_return"] / n
    out["mean_length"] = out["sum_length"] / n
    return out

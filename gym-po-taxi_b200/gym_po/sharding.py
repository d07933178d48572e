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


def bind_to_gpu_numa_node(device_index: int) -> str:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (one process per GPU: the host path's
    pinned buffers are then allocated next to the GPU's PCIe root and the ranks stop sharing one socket's memory
    controllers).  Best effort: returns a short description, or the reason nothing was done."""
    if os.environ.get("GPT_NO_NUMA_BIND"):
        return "disabled (GPT_NO_NUMA_BIND)"
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:   # virtualised PCI topology: ask the driver instead (nvidia-smi topo prints a CPU-affinity column)
            return _bind_from_smi_topo(device_index)
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"NUMA node {node} has no CPU in this process's cpuset"
        os.sched_setaffinity(0, cpus)
        return f"bound to NUMA node {node} ({len(cpus)} CPUs)"
    except Exception as e:  # pragma: no cover - depends on the box
        return f"not bound ({type(e).__name__}: {e})"


def _parse_cpulist(spec: str) -> set:
    cpus = set()
    for part in spec.split(","):
        lo, _, hi = part.strip().partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _bind_from_smi_topo(device_index: int) -> str:
    """Fallback of bind_to_gpu_numa_node: the "CPU Affinity" column of ``nvidia-smi topo -m`` for this GPU."""
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception as e:  # pragma: no cover - depends on the box
        return f"no NUMA information for the GPU (sysfs: -1; nvidia-smi topo: {type(e).__name__})"
    return bind_from_topo_text(out, device_index)


def bind_from_topo_text(text: str, device_index: int, apply: bool = True) -> str:
    """Parses the matrix ``nvidia-smi topo -m`` prints (tab separated; header has a "CPU Affinity" column) and pins
    the process to GPU<device_index>'s CPU list.  Split out (and ``apply=False``) so the CPU tests can feed it text."""
    import re
    clean = re.sub(r"\x1b\[[0-9;]*m", "", text)
    lines = [ln for ln in clean.splitlines() if ln.strip()]
    header = next((ln for ln in lines if "CPU Affinity" in ln), None)
    row = next((ln for ln in lines if ln is not header and re.match(rf"^GPU{device_index}\b", ln.strip())), None)
    if header is None or row is None:
        return "no NUMA information for the GPU (sysfs: -1; nvidia-smi topo: no CPU Affinity column)"
    # matrix cells are X / NV# / SYS / NODE / PHB / PXB / PIX: the first all-digit list in the row is the CPU affinity
    spec = next((c.strip() for c in re.split(r"\t+|\s{2,}", row) if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", c.strip())), None)
    if spec is None:
        return "no NUMA information for the GPU (sysfs: -1; nvidia-smi topo: CPU affinity not given)"
    cpus = _parse_cpulist(spec)
    cpus &= os.sched_getaffinity(0)
    if not cpus:
        return "nvidia-smi topo: the GPU's CPU list has no CPU in this process's cpuset"
    if apply:
        os.sched_setaffinity(0, cpus)
    return f"bound to the GPU's CPU affinity from nvidia-smi topo ({len(cpus)} CPUs)"


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

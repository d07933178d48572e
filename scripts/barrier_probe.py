"""Why did the first launches after dist.barrier() cost more at N > 1?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/barrier_probe.py

Times the 20-step Taxi block (two fused launches of 10 steps, 2^22 envs per GPU) with CUDA events under different
preludes and prints, per rank, the median of 15 trials of each in microseconds:
  busy        block issued back to back behind identical work (GPU never idle)
  sync        torch.cuda.synchronize() right before the first event (GPU idle for microseconds)
  idle_10ms / idle_300ms   synchronize, host sleeps, then the block (GPU idle: clocks / power state may drop)
  barrier     dist.barrier() + synchronize (what bench.py r01 did)
  barrier+1   dist.barrier() + synchronize + one untimed block (what bench.py does now)
"""
import os
import statistics
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gym-po-taxi_b200")]
from gym_po.envs import TaxiVecEnv  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, T, K = 1 << 22, 10, 20
env = TaxiVecEnv(B, seed=0, env_offset=rank * B)
cap = env.capacity
acts = torch.randint(0, 5, (T, cap), dtype=torch.int8, device=dev)
out = {n: torch.zeros((T,) + tuple(env._arrays[n].shape), dtype=env._arrays[n].dtype, device=dev) for n in ("obs", "reward", "terminated", "truncated")}
env.reset(seed=0)
env._arrays["elapsed"][:B] = torch.randint(0, 201, (B,), device=dev, dtype=torch.int32)


def block():
    for _ in range(K // T):
        env.step_many(acts, out)


def timed(prelude):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prelude()
    e0.record()
    block()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


preludes = {
    "busy": lambda: [block() for _ in range(8)],
    "sync": torch.cuda.synchronize,
    "idle_10ms": lambda: (torch.cuda.synchronize(), time.sleep(0.01)),
    "idle_300ms": lambda: (torch.cuda.synchronize(), time.sleep(0.3)),
    "barrier": barrier,
    "barrier+1": lambda: (barrier(), block()),
}
for _ in range(50):
    block()
torch.cuda.synchronize()
res = {}
for name, pre in preludes.items():
    res[name] = statistics.median(timed(pre) for _ in range(15))
for r in range(world):
    if r == rank:
        print(f"rank {rank}: " + "  ".join(f"{k}={v:.1f}us" for k, v in res.items()), flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()

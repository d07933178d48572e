"""Diagnostic timing of the fused Taxi launch under different episode dynamics (how much of the launch is the rare
reset / respawn path?):  python scripts/taxi_fused_probe.py [T=10]
  random   random actions, de-synchronised episode phases (the bench's workload)
  noreset  no-op actions and an unreachable time limit: no env ever resets"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-po-taxi_b200"))
from gym_po.envs import TaxiVecEnv  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
b = 1 << 22
dev = torch.device("cuda:0")
for mode in ("random", "noreset"):
    env = TaxiVecEnv(b, seed=0, time_limit=200 if mode == "random" else 1 << 30)
    gen = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, 5, (T, env.capacity), dtype=torch.int8, device=dev, generator=gen) if mode == "random" else \
        torch.full((T, env.capacity), 5, dtype=torch.int8, device=dev)
    out = {n: torch.zeros((T,) + tuple(env._arrays[n].shape), dtype=env._arrays[n].dtype, device=dev) for n in ("obs", "reward", "terminated", "truncated")}
    env.reset(seed=0)
    if mode == "random":
        env._arrays["elapsed"][:b] = torch.randint(0, 201, (b,), device=dev, generator=gen, dtype=torch.int32)
    for _ in range(20):
        env.step_many(acts, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        env.step_many(acts, out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{mode:8s} T={T}: {us:7.1f} us per launch, {b * T / us / 1e3:6.1f} G env-steps/s, {(11 * T + 18) * env.capacity / us / 1e3:6.0f} GB/s", flush=True)
    env.close()

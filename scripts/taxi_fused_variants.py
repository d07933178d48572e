"""Fused Taxi launch, TMA I/O against per-thread I/O, on the non-default configurations (8x8 map, Hansen observations,
three passengers) and launch lengths:  python scripts/taxi_fused_variants.py"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-po-taxi_b200"))
from gym_po.envs import EXTENDED_TAXI_MAP, TaxiVecEnv  # noqa: E402

b = 1 << 22
dev = torch.device("cuda:0")
cases = [("5x5", dict(), 10), ("5x5 hansen", dict(hansen_obs=True), 10), ("5x5 3 passengers", dict(num_passengers=3), 10),
         ("8x8", dict(map=EXTENDED_TAXI_MAP), 10), ("5x5 T=5", dict(), 5), ("5x5 T=20", dict(), 20), ("5x5 T=40", dict(), 40)]
for name, kw, T in cases:
    for io in ("tma", "threads"):
        env = TaxiVecEnv(b, seed=0, **kw)
        env.set_fused_steps(io)
        gen = torch.Generator(device=dev).manual_seed(1)
        acts = torch.randint(0, 5, (T, env.capacity), dtype=torch.int8, device=dev, generator=gen)
        out = {n: torch.zeros((T,) + tuple(env._arrays[n].shape), dtype=env._arrays[n].dtype, device=dev) for n in ("obs", "reward", "terminated", "truncated")}
        env.reset(seed=0)
        env._arrays["elapsed"][:b] = torch.randint(0, env.time_limit + 1, (b,), device=dev, generator=gen, dtype=torch.int32)
        for _ in range(10):
            env.step_many(acts, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(20, 2000 // T)
        e0.record()
        for _ in range(n):
            env.step_many(acts, out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"{name:18s} {io:8s} T={T:2d}: {us:7.1f} us per launch, {b * T / us / 1e3:6.1f} G env-steps/s, {(11 * T + 18) * env.capacity / us / 1e3:6.0f} GB/s", flush=True)
        env.close()
        del env, out, acts
        torch.cuda.empty_cache()

"""Probe: host-side cost of one env.step() call from Python (tiny batch, so the kernel itself is negligible)."""
import sys, time
sys.path.insert(0, "gym-po-taxi_b200")
import torch
from gym_po.envs import TaxiVecEnv, RoomsEnv

for name, env, n_act in (("taxi", TaxiVecEnv(512, device="cuda:0", seed=0), 5), ("rooms", RoomsEnv(512, device="cuda:0", seed=0), 8)):
    env.reset(seed=0)
    a = torch.randint(0, n_act, (env.capacity,), dtype=torch.int8, device="cuda:0")
    for _ in range(2000):
        env.step(a)
    torch.cuda.synchronize()
    n = 20000
    t0 = time.perf_counter()
    for _ in range(n):
        env.step(a)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: {1e6 * (t1 - t0) / n:.2f} us per step() call issued, {1e6 * (t2 - t0) / n:.2f} us incl. drain")

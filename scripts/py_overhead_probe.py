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

# graph mode: 32 captured steps replayed; per-step cost without any Python in the loop
for b in (512, 1 << 20):
    env = TaxiVecEnv(b, device="cuda:0", seed=0)
    env.reset(seed=0)
    env.set_graph_mode(True)
    a = torch.randint(0, 5, (env.capacity,), dtype=torch.int8, device="cuda:0")
    env.step(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(32):
            env.step(a)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        g.replay()
    torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / (200 * 32)
    eager = TaxiVecEnv(b, device="cuda:0", seed=0)
    eager.reset(seed=0)
    for _ in range(500):
        eager.step(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(6400):
        eager.step(a)
    torch.cuda.synchronize()
    per_e = (time.perf_counter() - t0) / 6400
    print(f"taxi B={b}: graph replay {1e6 * per:.2f} us per step, eager Python loop {1e6 * per_e:.2f} us per step")

"""TEST/BENCH INFRASTRUCTURE — times the oracle port (the CPU restatement of the reference's numpy
vectorized step) on host cores.  Used only by bench.py's ``cpu_baseline`` leg and ``--impl reference``.

    python -m oracle.cpu_bench --family taxi --envs 65536 --steps 150 --warmup 20 --start-at <monotonic>

One process = one core (numpy elementwise ops are single-threaded).  bench.py launches one worker per
available core and aggregates: total env-steps / (last end - first start).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def make_env(family, b, seed):
    import oracle
    draws = oracle.GeneratorDraws(seed=seed)
    if family == "taxi":
        return oracle.TaxiOracle(b, draws=draws), 5
    if family == "rooms_hansen8":
        return oracle.RoomsOracle(b, "4", obs_type="hansen8", draws=draws), 8
    if family == "rooms_grid5":
        return oracle.RoomsOracle(b, "4", obs_type="grid", obs_n=5, draws=draws), 8
    if family == "rooms_grid9":
        return oracle.RoomsOracle(b, "4", obs_type="grid", obs_n=9, draws=draws), 8
    if family == "crooms":
        return oracle.CRoomsOracle(b, "4", obs_type="vector_mdp", draws=draws), 0
    if family == "msrooms":
        return oracle.MSRoomsOracle(b, grid_z=3, draws=draws), 4
    if family == "car":
        return oracle.CarOracle(b, draws=draws), -1
    if family == "tag":
        return oracle.TagOracle(b, draws=draws), 0
    raise KeyError(family)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="taxi")
    ap.add_argument("--envs", type=int, default=1 << 16)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--start-at", type=float, default=0.0)
    a = ap.parse_args()
    env, n_act = make_env(a.family, a.envs, a.seed)
    env.reset()
    rng = np.random.default_rng(1234 + a.seed)
    # de-synchronise episode phases exactly like the GPU arm (otherwise every env truncates on the same step)
    env.elapsed[:] = rng.integers(0, env.time_limit + 1, size=a.envs)
    if n_act > 0:
        acts = rng.integers(n_act, size=(8, a.envs))
    else:
        acts = rng.uniform(-1, 1, size=(8, a.envs, 2 if n_act == 0 else 1)).astype(np.float32)
    for t in range(a.warmup):
        env.step(acts[t % 8])
    while time.monotonic() < a.start_at:
        time.sleep(0.001)
    t0 = time.monotonic()
    for t in range(a.steps):
        env.step(acts[t % 8])
    t1 = time.monotonic()
    print(json.dumps({"envs": a.envs, "steps": a.steps, "t0": t0, "t1": t1}))


if __name__ == "__main__":
    main()

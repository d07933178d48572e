"""TEST/BENCH INFRASTRUCTURE — times the reference's CPU vectorized step on host cores.

    python -m oracle.cpu_bench --impl reference --family taxi --envs 262144 --steps 25 --warmup 5 --start-at <monotonic>

``--impl reference`` steps the UNMODIFIED reference package (``baseline/_ref/gym_po``, installed there by
``__graft_entry__.build()`` with ``pip install --target``; loaded through ``oracle.ref_loader``, i.e. with the
third-party stand-ins and the one documented in-memory signature repair).  ``--impl port`` steps the numpy oracle
port (``oracle/``) — the fallback when ``baseline/_ref`` is absent, and the only implementation of the point-mass
Tag env (the reference's AntTag is a scalar MuJoCo env).  Used only by bench.py's ``cpu_baseline`` leg and
``--impl reference``.

One process = one core (numpy elementwise ops are single-threaded).  bench.py launches one worker per available
core, each with its share of the batch, and aggregates: total env-steps / (last end - first start).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")

# families the reference implements as an internally vectorized env (everything but the point-mass Tag)
REFERENCE_FAMILIES = ("taxi", "rooms_hansen8", "rooms_grid5", "rooms_grid9", "crooms", "msrooms", "car")


def reference_installed() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "gym_po", "__init__.py"))


def make_port(family, b, seed):
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


def make_reference(family, b, seed):
    """The reference's own classes with the constructor arguments of bench.py's make_env()."""
    os.environ["GYM_PO_REFERENCE_ROOT"] = REF_ROOT
    from oracle import ref_loader
    ref_loader.REFERENCE_ROOT = REF_ROOT
    envs = ref_loader.load_reference()
    import importlib
    rooms = importlib.import_module(ref_loader.ALIAS + ".envs.rooms")
    if family == "taxi":
        return envs.TaxiVecEnv(b), 5
    if family == "rooms_hansen8":
        return rooms.RoomsEnv(b, "4", obs_type="hansen8"), 8
    if family == "rooms_grid5":
        return rooms.RoomsEnv(b, "4", obs_type="grid", obs_n=5), 8
    if family == "rooms_grid9":
        return rooms.RoomsEnv(b, "4", obs_type="grid", obs_n=9), 8
    if family == "crooms":
        return rooms.CRoomsEnv(b, "4", obs_type="vector_mdp"), 0
    if family == "msrooms":
        ms = importlib.import_module(ref_loader.ALIAS + ".envs.rooms.msrooms")
        return ms.MultistoryFourRoomsEnv(b, grid_z=3), 4
    if family == "car":
        return envs.CarVecEnv(b), -1
    raise KeyError(family)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="port", choices=["port", "reference"])
    ap.add_argument("--family", default="taxi")
    ap.add_argument("--envs", type=int, default=1 << 16)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--start-at", type=float, default=0.0)
    ap.add_argument("--sync-stdin", action="store_true",
                    help="after the warm-up print READY <s per warm-up step>, then read 'GO <start_at> <steps>' from stdin")
    a = ap.parse_args()
    if a.impl == "reference":
        env, n_act = make_reference(a.family, a.envs, a.seed)
        env.reset(seed=a.seed)
    else:
        env, n_act = make_port(a.family, a.envs, a.seed)
        env.reset()
    rng = np.random.default_rng(1234 + a.seed)
    # de-synchronise episode phases exactly like the GPU arm (otherwise every env truncates on the same step)
    env.elapsed[:] = rng.integers(0, env.time_limit + 1, size=a.envs)
    if n_act > 0:
        acts = rng.integers(n_act, size=(8, a.envs))
    else:
        acts = rng.uniform(-1, 1, size=(8, a.envs, 2 if n_act == 0 else 1)).astype(np.float32)
    tw = time.monotonic()
    for t in range(a.warmup):
        env.step(acts[t % 8])
    tw = (time.monotonic() - tw) / max(a.warmup, 1)
    if a.sync_stdin:   # all workers of a run start together, on the parent's word
        print(f"READY {tw!r}", flush=True)
        go = sys.stdin.readline().split()
        a.start_at, a.steps = float(go[1]), int(go[2])
    while time.monotonic() < a.start_at:
        time.sleep(0.001)
    t0 = time.monotonic()
    for t in range(a.steps):
        env.step(acts[t % 8])
    t1 = time.monotonic()
    print(json.dumps({"envs": a.envs, "steps": a.steps, "t0": t0, "t1": t1, "impl": a.impl}))


if __name__ == "__main__":
    main()

"""Generates render_taxi_*.npz by EXECUTING THE REAL REFERENCE's TaxiVecEnv.render (build container only:
needs /root/reference and cv2).  Each fixture holds, for a handful of moments of a seeded rollout, the encoded
states, the name of env 0's last action (or '') and the RGB frame the reference drew for idx = arange(k)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.draws import make_generator  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

CASES = [("render_taxi_5x5", "TaxiVecEnv", 5), ("render_taxi_5x5_hansen", "HansenTaxiVecEnv", 3),
         ("render_taxi_8x8", "ExtendedTaxiVecEnv", 4), ("render_taxi_8x8_hansen", "ExtendedHansenTaxiVecEnv", 1)]


def main():
    E = load_reference()
    for name, cls, k in CASES:
        env = getattr(E, cls)(8, time_limit=15)
        env._np_random = make_generator(7)
        env.reset()
        rng = np.random.default_rng(3)
        states, names, frames = [], [], []
        for t in range(60):
            if t % 7 == 0:
                states.append(env.s[:k].copy())
                names.append("" if env.lastaction is None else env.ACTION_NAMES[env.lastaction])
                frames.append(env.render(idx=np.arange(k)).copy())
            a = rng.integers(5, size=8)
            if t % 2:
                a[:] = 4          # plenty of pickups so that full taxis / passengers under the taxi show up
            env.step(a)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), states=np.array(states), names=np.array(names),
                            frames=np.array(frames), cls=cls)
        print(name, np.array(frames).shape, sorted(set(names)))
    # car-flag: env 0's number-line frame along a rollout that drives towards the priest, then a flag
    env = E.CarVecEnv(4, time_limit=400, render_mode="rgb_array")
    env._np_random = make_generator(11)
    env.reset()
    rows = []
    for t in range(330):
        target = env.priests[0] if t < 120 else env.heavens[0]
        env.step(np.full((4, 1), np.sign(target - env.s[0, 0]), dtype=np.float32))
        if t % 15 == 0:
            rows.append((env.s[0].copy(), env.heavens[0], env.priests[0], env.render().copy()))
    np.savez_compressed(os.path.join(HERE, "render_car.npz"), s=np.array([r[0] for r in rows]),
                        heavens=np.array([r[1] for r in rows]), priests=np.array([r[2] for r in rows]),
                        frames=np.array([r[3] for r in rows]))
    print("render_car", len(rows), sorted({float(r[0][2]) for r in rows}), sorted({float(r[1]) for r in rows}))


if __name__ == "__main__":
    main()

"""Generates the golden fixtures in this directory by EXECUTING THE REAL REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the reference env is constructed with the recorded kwargs, its random
generator is wrapped in a recording proxy (oracle.ref_loader.RecordingGenerator), and a
seeded action stream is stepped.  Stored per case (one ``<name>.npz``):

* ``meta``            json: env class, kwargs, seed, B, T, sha256 (SURVEY.md Appendix D recipe)
* ``actions``         [T,B] (int8) or [T,B,2] (float64)
* ``obs0`` / ``obs``  reset obs, per-step obs [T,...]  (narrowest exact integer dtype, or float64)
* ``rew`` float32 [T,B], ``term`` / ``trunc`` bool [T,B]
* ``state_*``         the env's public state arrays after the last step
* ``log_kind`` / ``log_size`` / ``log_val``  every Generator call's result, in call order
  (multinomial results are stored as their argmax — the only thing the reference uses)

``tag_move_target.npz`` holds input/output vectors of ``AntTagEnv._move_target`` called unbound
on a stand-in object (works without MuJoCo).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.draws import make_generator  # noqa: E402
from oracle.ref_loader import RecordingGenerator, load_reference  # noqa: E402

KINDS = ["multinomial_argmax", "integers", "random", "choice", "normal", "uniform"]

B = 64
CASES = [
    # name, class, kwargs, n_act (0 = continuous yx), T
    ("taxi", "TaxiVecEnv", {}, 5, 500),
    ("taxi_hansen", "HansenTaxiVecEnv", {}, 5, 500),
    ("taxi_ext", "ExtendedTaxiVecEnv", {}, 5, 500),
    ("taxi_ext_hansen", "ExtendedHansenTaxiVecEnv", {}, 5, 500),
    ("taxi_multi", "TaxiVecEnv", {"num_passengers": 3, "time_limit": 2000}, 5, 3000),
    ("taxi_rewards", "TaxiVecEnv", {"num_passengers": 2, "time_limit": 50, "reward_goal": 20.0,
                                    "reward_bad": -10.0, "reward_any": -1.0}, 5, 300),
    ("rooms_hansen8", "RoomsEnv", {"layout": "4", "obs_type": "hansen8"}, 8, 500),
    ("rooms_hansen4_card", "RoomsEnv", {"layout": "4", "obs_type": "hansen", "action_type": "cardinal"}, 4, 500),
    ("rooms_vhansen8", "RoomsEnv", {"layout": "4", "obs_type": "vector_hansen8"}, 8, 500),
    ("rooms_vghansen8", "RoomsEnv", {"layout": "4", "obs_type": "vector_goal_hansen8"}, 8, 500),
    ("rooms_grid5", "RoomsEnv", {"layout": "4", "obs_type": "grid", "obs_n": 5}, 8, 500),
    ("rooms_grid9", "RoomsEnv", {"layout": "4", "obs_type": "grid", "obs_n": 9}, 8, 500),
    ("rooms32_grid9_rgoal", "RoomsEnv", {"layout": "32", "obs_type": "grid", "obs_n": 9, "goal_xy": None}, 8, 500),
    ("rooms_mdp", "RoomsEnv", {"layout": "4", "obs_type": "mdp"}, 8, 500),
    # extra coverage beyond SURVEY Appendix D
    ("rooms_room", "RoomsEnv", {"layout": "8", "obs_type": "room", "time_limit": 60}, 8, 200),
    ("rooms_room_goal_rg", "RoomsEnv", {"layout": "10b", "obs_type": "room_goal", "goal_xy": None, "time_limit": 40}, 8, 200),
    ("rooms_mdp_goal_rg", "RoomsEnv", {"layout": "16", "obs_type": "mdp_goal", "goal_xy": None, "time_limit": 40}, 8, 200),
    ("rooms_vmdp", "RoomsEnv", {"layout": "2", "obs_type": "vector_mdp", "time_limit": 30}, 8, 200),
    ("rooms_vmdp_goal_rg", "RoomsEnv", {"layout": "8b", "obs_type": "vector_mdp_goal", "goal_xy": None, "time_limit": 30}, 8, 200),
    ("rooms_hansen8_rg_rewards", "RoomsEnv", {"layout": "16b", "obs_type": "hansen8", "goal_xy": None, "time_limit": 25,
                                              "action_failure_probability": 0.35, "step_reward": -0.1,
                                              "wall_reward": -0.5, "goal_reward": 3.0}, 8, 300),
    ("rooms_vghansen4_rg", "RoomsEnv", {"layout": "1", "obs_type": "vector_goal_hansen", "action_type": "cardinal",
                                        "goal_xy": None, "time_limit": 20}, 4, 200),
    ("rooms_grid3_rg", "RoomsEnv", {"layout": "32b", "obs_type": "grid", "obs_n": 3, "goal_xy": None, "time_limit": 30}, 8, 200),
    ("rooms_grid7_goalxy", "RoomsEnv", {"layout": "10", "obs_type": "grid", "obs_n": 7, "goal_xy": (3, 2), "time_limit": 50}, 8, 200),
    # the reference's default goal for '32'/'32b' (ENDS = x 47, y 32) lies OUTSIDE the 25 x 49 grid: unreachable goal
    ("rooms32_hansen8_defgoal", "RoomsEnv", {"layout": "32", "obs_type": "hansen8", "time_limit": 30}, 8, 150),
    ("rooms32b_grid5_defgoal", "RoomsEnv", {"layout": "32b", "obs_type": "grid", "obs_n": 5, "time_limit": 30}, 8, 150),
    ("rooms32_vghansen8_defgoal", "RoomsEnv", {"layout": "32", "obs_type": "vector_goal_hansen8", "time_limit": 30}, 8, 150),
    ("crooms32_vghansen8_defgoal", "CRoomsEnv", {"layout": "32", "obs_type": "vector_goal_hansen8", "time_limit": 30}, 0, 150),
    # multistory FourRooms (needs the Appendix-C signature repair to import at all)
    ("msrooms_mdp_1floor", "MultistoryFourRoomsEnv", {"grid_z": 1}, 4, 300),
    ("msrooms_hansen8_3floors", "MultistoryFourRoomsEnv", {"grid_z": 3, "obs_type": "hansen8", "action_type": "ordinal", "time_limit": 150}, 8, 400),
    ("msrooms_vghansen_rg", "MultistoryFourRoomsEnv", {"grid_z": 2, "obs_type": "vector_goal_hansen", "goal_xyz": None, "time_limit": 40}, 4, 300),
    ("msrooms_vmdp_goal_rg", "MultistoryFourRoomsEnv", {"grid_z": 2, "obs_type": "vector_mdp_goal", "goal_xyz": None, "time_limit": 30,
                                                        "step_reward": -0.1, "wall_reward": -1.0}, 4, 300),
    ("msrooms_room_goal_rg", "MultistoryFourRoomsEnv", {"grid_z": 2, "obs_type": "room_goal", "goal_xyz": None, "time_limit": 30}, 4, 200),
    ("msrooms_vhansen8_ord", "MultistoryFourRoomsEnv", {"grid_z": 3, "obs_type": "vector_hansen8", "action_type": "ordinal", "time_limit": 60,
                                                         "action_failure_probability": 0.1}, 8, 250),
    # car-flag.  n_act -1 = float32 forces [B,1] steered towards the flags, -2 = float64 forces
    ("car_f32", "CarVecEnv", {"time_limit": 60}, -1, 400),
    ("car_f64", "CarVecEnv", {"time_limit": 45}, -2, 300),
    ("car_discrete5", "DiscreteActionCarVecEnv", {"num_actions": 5, "time_limit": 50}, 5, 300),
    ("crooms_vmdp", "CRoomsEnv", {"layout": "4", "obs_type": "vector_mdp"}, 0, 500),
    ("crooms_hansen8_ord", "CRoomsEnv", {"layout": "4", "obs_type": "hansen8", "action_type": "ordinal"}, 8, 500),
    ("crooms_vel_rg", "CRoomsEnv", {"layout": "8", "obs_type": "vector_mdp_goal", "use_velocity": True, "goal_xy": None,
                                    "time_limit": 60, "wall_reward": -0.25, "step_reward": -0.01}, 0, 300),
    ("crooms_card_nostd", "CRoomsEnv", {"layout": "2", "obs_type": "mdp", "action_type": "cardinal", "action_std": 0.0,
                                        "time_limit": 40}, 4, 200),
    ("crooms_grid5_power", "CRoomsEnv", {"layout": "16", "obs_type": "grid", "obs_m": 5, "action_power": 1.7,
                                         "goal_threshold": 1.25, "time_limit": 80}, 0, 200),
]


def narrow(x):
    x = np.asarray(x)
    if x.dtype.kind == "f":
        if np.array_equal(x, np.round(x)) and np.abs(x).max(initial=0) < 32000:
            return x.astype(np.int16)
        return x.astype(np.float64)
    if x.dtype.kind in "iu":
        m = np.abs(x).max(initial=0)
        return x.astype(np.int8 if m < 127 else np.int16 if m < 32000 else np.int32)
    return x


def pack_log(log):
    kinds, sizes, vals = [], [], []
    for k, v in log:
        v = np.asarray(v)
        if k == "multinomial":
            k, v = "multinomial_argmax", v.argmax(-1)
        kinds.append(KINDS.index(k))
        sizes.append(v.size)
        vals.append(v.astype(np.float64).ravel())
    return (np.array(kinds, np.int8), np.array(sizes, np.int64),
            np.concatenate(vals) if vals else np.zeros(0))


def run_case(E, name, cls, kwargs, n_act, T, seed=0):
    if cls == "DiscreteActionCarVecEnv":
        kw = dict(kwargs)
        env = E.DiscreteActionCarVecEnv(kw.pop("num_actions"), B, **kw)
    else:
        env = getattr(E, cls)(B, **kwargs)
    rec = RecordingGenerator(make_generator(seed))
    if cls == "CRoomsEnv":
        env.rng = rec            # crooms.py:168, :246-249 — own generator
    else:
        env._np_random = rec     # gymnasium.Env.np_random
    out = env.reset()
    obs0 = out[0] if isinstance(out, tuple) else out
    h = hashlib.sha256()
    feed = lambda x: h.update(np.ascontiguousarray(x).astype(np.float64).tobytes())
    feed(obs0)
    obs0 = np.array(obs0, copy=True)
    arng = np.random.default_rng(1)
    A, O, R, D, TR = [], [], [], [], []
    for _ in range(T):
        if n_act < 0:      # car: random force magnitudes (some beyond the clip range), mostly pushing outwards
            a = arng.uniform(-1.4, 1.4, (B, 1))
            if len(A) % 3:
                a = np.sign(env.s[:, :1] + 1e-3) * np.abs(a)
            a = a.astype(np.float32 if n_act == -1 else np.float64)
        elif cls == "DiscreteActionCarVecEnv":   # outermost force towards the nearer flag 2 steps out of 3
            a = arng.integers(n_act, size=B)
            if len(A) % 3:
                a = np.where(env.s[:, 0] >= 0, n_act - 1, 0)
        else:
            a = arng.uniform(-1, 1, (B, 2)) if n_act == 0 else arng.integers(n_act, size=B)
        o, r, d, tr, _ = env.step(a.copy())
        feed(o); feed(r); feed(d); feed(tr)
        A.append(a); O.append(np.array(o, copy=True)); R.append(r.copy()); D.append(d.copy()); TR.append(tr.copy())
    state = {}
    if cls == "CRoomsEnv":
        state = {"state_agent": env.agent_yx, "state_goal": env.goal_yx, "state_velocity": env.agent_yx_velocity,
                 "state_elapsed": env.elapsed}
    elif cls == "RoomsEnv":
        state = {"state_agent": env.agent_yx, "state_goal": env.goal_yx, "state_elapsed": env.elapsed}
    elif cls == "MultistoryFourRoomsEnv":
        state = {"state_agent": env.agent_zyx, "state_goal": env.goal_zyx, "state_elapsed": env.elapsed}
    elif "Car" in cls:
        state = {"state_s": env.s, "state_elapsed": env.elapsed, "state_heavens": env.heavens, "state_priests": env.priests}
    else:
        state = {"state_s": env.s, "state_elapsed": env.elapsed, "state_ndrop": env.n_dropoffs_completed}
    kinds, sizes, vals = pack_log(rec.log)
    meta = {"cls": cls, "kwargs": kwargs, "seed": seed, "B": B, "T": T, "n_act": n_act,
            "sha256": h.hexdigest(), "numpy": np.__version__,
            "sum_term": int(np.sum(D)), "sum_trunc": int(np.sum(TR)), "sum_rew": float(np.sum(R, dtype=np.float64))}
    A = np.array(A)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), meta=json.dumps(meta),
        actions=A.astype(np.int8) if n_act > 0 else A, obs0=obs0 if "Car" in cls else narrow(obs0),
        obs=np.array(O) if "Car" in cls else narrow(np.array(O)),
        rew=np.array(R, np.float32), term=np.array(D, bool), trunc=np.array(TR, bool),
        log_kind=kinds, log_size=sizes, log_val=vals, **{k: np.asarray(v) for k, v in state.items()})
    return meta


def tag_vectors(E):
    import importlib
    mod = importlib.import_module("_gym_po_reference.envs.ant_tag")

    class Fake:
        pass

    rng = np.random.default_rng(7)
    n = 4096
    ant = rng.uniform(-5, 5, (n, 2))
    tgt = rng.uniform(-4.5, 4.5, (n, 2))
    choice = rng.integers(4, size=n)
    out = np.zeros((n, 2))
    for i in range(n):
        f = Fake()
        f.cage_max_xy = np.full((2,), 4.5)
        f.target_step = 0.5
        f.np_random = type("R", (), {"integers": staticmethod(lambda k, c=int(choice[i]): c)})()
        f.data = type("D", (), {})()
        f.data.mocap_pos = np.zeros((3, 3))
        mod.AntTagEnv._move_target(f, ant[i].copy(), tgt[i].copy())
        out[i] = f.data.mocap_pos[0, :2]
    np.savez_compressed(os.path.join(HERE, "tag_move_target.npz"), ant=ant, target=tgt, choice=choice.astype(np.int8),
                        new_target=out)


def layout_grids(E):
    """the reference's parsed integer grids for all 12 layouts (pins the shared text asset)"""
    import importlib
    L = importlib.import_module("_gym_po_reference.envs.rooms.layouts")
    grids = {"grid_" + k: L.np_to_grid(L.layout_to_np(v)).astype(np.int8) for k, v in L.LAYOUTS.items()}
    np.savez_compressed(os.path.join(HERE, "layout_grids.npz"), **grids)


def main():
    E = load_reference()
    for name, cls, kwargs, n_act, T in CASES:
        m = run_case(E, name, cls, kwargs, n_act, T)
        print(f"{name:28s} term={m['sum_term']:4d} trunc={m['sum_trunc']:4d} rew={m['sum_rew']:12.4f} sha={m['sha256'][:16]}")
    tag_vectors(E)
    layout_grids(E)


if __name__ == "__main__":
    main()

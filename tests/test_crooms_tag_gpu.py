"""GPU: continuous ROOMS and point-mass Tag vs the oracle.  The kernels compute in float64 with numpy's
operation order and no FMA, so on replayed draws the comparison is BIT-EXACT (stricter than the 1e-5
relative tolerance BASELINE.json's north_star allows for the continuous variants)."""
import numpy as np
import pytest
import torch

import oracle
from helpers import golden_names, load_golden, make_oracle, recorded_draws

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cmp(g, o, t):
    for name, x, y in zip(("obs", "reward", "terminated", "truncated"), g, o):
        np.testing.assert_array_equal(x.cpu().numpy().astype(np.float64), np.asarray(y, dtype=np.float64),
                                      err_msg=f"{name} at step {t}")


def _kw(meta):
    kw = dict(meta["kwargs"])
    if kw.get("goal_xy", 0) is not None and "goal_xy" in kw:
        kw["goal_xy"] = tuple(kw["goal_xy"])
    return kw


def _actions(a, n_act):
    if n_act:
        return torch.as_tensor(a, dtype=torch.int8, device=DEV)
    return torch.as_tensor(a, dtype=torch.float64, device=DEV)


@pytest.mark.parametrize("name", golden_names("crooms"))
def test_crooms_golden_trajectory_free_running(name):
    from gym_po.envs import CRoomsEnv
    fx = load_golden(name)
    meta = fx["meta"]
    orc = make_oracle(meta, recorded_draws(fx))
    env = CRoomsEnv(meta["B"], device=DEV, rng_mode="replay", action_dtype=torch.float64, **_kw(meta))
    orc.reset()
    env.set_replay(**orc.draws)
    obs = env.reset()
    np.testing.assert_array_equal(obs.cpu().numpy(), fx["obs0"])
    for t in range(meta["T"]):
        a = fx["actions"][t]
        orc.step(a)
        env.set_replay(**orc.draws)
        g = env.step(_actions(a, meta["n_act"]))
        _cmp(g[:4], (fx["obs"][t], fx["rew"][t], fx["term"][t], fx["trunc"][t]), t)
    st = env.get_state()
    np.testing.assert_array_equal(st["agent"].cpu().numpy(), fx["state_agent"])
    np.testing.assert_array_equal(st["goal"].cpu().numpy(), fx["state_goal"])
    np.testing.assert_array_equal(st["velocity"].cpu().numpy(), fx["state_velocity"])
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), fx["state_elapsed"])


@pytest.mark.parametrize("kw", [
    dict(obs_type="vector_mdp", action_type="yx"),
    dict(obs_type="vector_mdp_goal", action_type="yx", goal_xy=None, use_velocity=True, cell_size=1.0),
    dict(obs_type="hansen8", action_type="ordinal", goal_xy=None),
    dict(obs_type="vector_goal_hansen", action_type="cardinal", action_std=0.0),
    dict(obs_type="grid", obs_m=7, action_type="yx", goal_xy=None, action_power=2.5),
    dict(obs_type="room_goal", action_type="yx", goal_xy=None, use_velocity=True),
    dict(obs_type="mdp_goal", action_type="ordinal", action_std=0.5, goal_threshold=1.5),
])
def test_crooms_lockstep(kw):
    from gym_po.envs import CRoomsEnv
    b = 5000
    kw = dict(layout="8", time_limit=40, step_reward=-0.01, wall_reward=-0.3, **kw)
    orc = oracle.CRoomsOracle(b, draws=oracle.GeneratorDraws(seed=8), **kw)
    env = CRoomsEnv(b, device=DEV, rng_mode="replay", action_dtype=torch.float64, **kw)
    o = orc.reset()
    env.set_replay(**orc.draws)
    np.testing.assert_array_equal(env.reset().cpu().numpy(), o)
    rng = np.random.default_rng(4)
    n_act = 0 if kw["action_type"] == "yx" else (4 if kw["action_type"] == "cardinal" else 8)
    for t in range(160):
        a = rng.uniform(-1, 1, (b, 2)) if n_act == 0 else rng.integers(n_act, size=b)
        o = orc.step(a)
        env.set_replay(**orc.draws)
        _cmp(env.step(_actions(a, n_act))[:4], o[:4], t)
    st = env.get_state()
    np.testing.assert_array_equal(st["agent"].cpu().numpy(), orc.agent)
    np.testing.assert_array_equal(st["velocity"].cpu().numpy(), orc.velocity)


def test_crooms_float32_actions_and_philox_invariants():
    """float32 action tensors (the reference's Box dtype) are widened exactly; Philox mode keeps agents
    inside walkable cells and noise statistics match N(0, action_std^2)."""
    from gym_po.envs import CRoomsEnv
    b = 1 << 20
    env = CRoomsEnv(b, "4", obs_type="vector_mdp", device=DEV, seed=5)
    obs = env.reset(seed=5)
    grid = torch.as_tensor(env.grid, device=DEV)
    cells = obs.floor().long()
    assert bool((grid[cells[:, 0], cells[:, 1]] >= 0).all())
    assert bool(((obs - obs.floor()) == 0.5).all())          # spawn at cell centres
    # one step with zero action from the centre of an open area: displacement = noise ~ N(0, 0.2^2)
    start = np.tile([4.5, 4.5], (b, 1))
    env.set_state(agent=start, goal=None, velocity=None, elapsed=np.zeros(b, dtype=int))
    obs, rew, term, trunc, _ = env.step(torch.zeros((env.capacity, 2), dtype=torch.float32, device=DEV))
    d = (obs - 4.5).cpu().numpy()
    assert abs(d.mean()) < 1e-3 and abs(d.std() - 0.2) < 1e-3
    assert abs(np.mean(d[:, 0] * d[:, 1])) < 1e-3             # independent components
    assert abs((np.abs(d) > 0.4).mean() - 0.0455) < 2e-3     # 2-sigma tail
    gen = torch.Generator(device=DEV).manual_seed(0)
    for t in range(50):
        a = torch.rand((env.capacity, 2), device=DEV, generator=gen) * 2 - 1
        obs, rew, term, trunc, _ = env.step(a)
        cells = obs.floor().long()
        assert bool((grid[cells[:, 0], cells[:, 1]] >= 0).all())
        assert bool((rew[term] == 1.0).all())


def test_tag_lockstep_vs_oracle():
    from gym_po.envs import TagVecEnv
    b = 20_000
    orc = oracle.TagOracle(b, time_limit=60, draws=oracle.GeneratorDraws(seed=2))
    env = TagVecEnv(b, time_limit=60, device=DEV, rng_mode="replay", action_dtype=torch.float64)
    o, _ = orc.reset()
    env.set_replay(**orc.draws)
    g, info = env.reset()
    np.testing.assert_array_equal(g.cpu().numpy(), o)
    rng = np.random.default_rng(9)
    n_tag = 0
    for t in range(200):
        # chase the target when visible, else random: produces tags, cage hits and truncations
        a = np.clip(orc.target - orc.agent, -1, 1) * (rng.random((b, 1)) < 0.7) + rng.uniform(-1, 1, (b, 2)) * 0.3
        a = np.clip(a, -1, 1)
        o = orc.step(a)
        env.set_replay(**orc.draws)
        gg = env.step(torch.as_tensor(a, dtype=torch.float64, device=DEV))
        _cmp(gg[:4], o[:4], t)
        n_tag += int(o[2].sum())
    assert n_tag > 100
    st = env.get_state()
    np.testing.assert_array_equal(st["agent"].cpu().numpy(), orc.agent)
    np.testing.assert_array_equal(st["target"].cpu().numpy(), orc.target)
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed)


def test_tag_philox_invariants():
    from gym_po.envs import TagVecEnv
    b = 1 << 20
    env = TagVecEnv(b, device=DEV, seed=3)
    obs, _ = env.reset(seed=3)
    d0 = (env.agent_xy - env.target_xy).norm(dim=-1)
    assert bool((d0 > 5.0).all()) and bool((env.agent_xy.abs() <= 4.5).all()) and bool((env.target_xy.abs() <= 4.5).all())
    assert bool((obs == 0).all())                              # nothing visible at spawn distance > 5
    gen = torch.Generator(device=DEV).manual_seed(1)
    for t in range(40):
        a = torch.rand((env.capacity, 2), device=DEV, generator=gen) * 2 - 1
        obs, rew, term, trunc, _ = env.step(a)
        assert bool((env.agent_xy.abs() <= 5.0).all()) and bool((env.target_xy.abs() <= 4.5).all())
        d = (env.agent_xy - env.target_xy).norm(dim=-1)
        vis = d < 3.0
        assert bool((obs[~vis] == 0).all()) and bool((obs[vis] == env.target_xy[vis]).all())
        assert bool((rew[term] == 1.0).all()) and bool((rew[~term] == 0.0).all())


# ---- float32 fast mode: tolerance-checked against the float64 oracle ------------------------------
# Stated tolerance: |gpu - oracle| <= 1e-5 * max(1, |oracle|) per coordinate after ONE step from an injected
# (float32-representable) state with replayed draws, on every env whose discrete decisions (reward,
# terminated, truncated) agree; decisions may legitimately flip when a float32 rounding moves a position
# across a cell boundary / threshold — those envs are counted and must be rarer than 2e-4.
F32_RTOL = 1e-5
F32_FLIP_RATE = 2e-4


def test_crooms_float32_philox_stays_in_walkable_cells():
    """float32 fast mode, Philox (the benchmarked CRooms configuration): agents pushed against walls for many steps
    never end up in a wall cell (the jitter's upper clip must stay strictly inside the cell: `centre + half - 1e-8` is
    not representable in float32), spawn at cell centres, noise law N(0, action_std^2)."""
    from gym_po.envs import CRoomsEnv
    b = 1 << 20
    env = CRoomsEnv(b, "4", obs_type="vector_mdp", device=DEV, seed=9, precision="float32")
    obs = env.reset(seed=9)
    assert obs.dtype == torch.float32
    grid = torch.as_tensor(env.grid, device=DEV)
    cells = obs.floor().long()
    assert bool((grid[cells[:, 0], cells[:, 1]] >= 0).all())
    assert bool(((obs - obs.floor()) == 0.5).all())
    gen = torch.Generator(device=DEV).manual_seed(0)
    blocked = 0
    for t in range(120):
        # a constant diagonal push keeps most agents pressed into a corner: lots of rejected moves
        a = torch.ones((env.capacity, 2), device=DEV) * (1.0 if t % 40 < 20 else -1.0)
        a += (torch.rand((env.capacity, 2), device=DEV, generator=gen) - 0.5) * 0.2
        before = env.agent_yx.clone()
        obs, rew, term, trunc, _ = env.step(a)
        cells = obs[:b].floor().long()
        assert bool((grid[cells[:, 0], cells[:, 1]] >= 0).all()), f"agent inside a wall cell at step {t}"
        stay = (before[:b].floor() == obs[:b].floor()).all(-1) & ~(term[:b] | trunc[:b])
        blocked += int(stay.sum())
    assert blocked > 20 * b     # the scenario really exercises the rejection branch
    start = np.tile([4.5, 4.5], (b, 1))
    env.set_state(agent=start, goal=None, velocity=None, elapsed=np.zeros(b, dtype=int))
    obs, *_ = env.step(torch.zeros((env.capacity, 2), dtype=torch.float32, device=DEV))
    d = (obs[:b] - 4.5).double().cpu().numpy()
    assert abs(d.mean()) < 1e-3 and abs(d.std() - 0.2) < 1e-3
    assert abs((np.abs(d) > 0.4).mean() - 0.0455) < 2e-3


def _close(g, o):
    o = np.asarray(o, dtype=np.float64)
    return np.abs(g.astype(np.float64) - o) <= F32_RTOL * np.maximum(1.0, np.abs(o))


def test_crooms_float32_mode_within_tolerance():
    from gym_po.envs import CRoomsEnv
    b = 100_000
    kw = dict(layout="8", obs_type="vector_mdp", action_type="yx", use_velocity=True, goal_xy=None, time_limit=12,
              wall_reward=-0.3, step_reward=-0.01)
    orc = oracle.CRoomsOracle(b, draws=oracle.GeneratorDraws(seed=3), **kw)
    env = CRoomsEnv(b, device=DEV, rng_mode="replay", precision="float32", **kw)
    assert env.agent_yx.dtype == torch.float32
    o = orc.reset()
    env.set_replay(**orc.draws)
    np.testing.assert_array_equal(env.reset().cpu().numpy(), o.astype(np.float32))
    rng = np.random.default_rng(1)
    flips = total = 0
    for t in range(40):
        st = orc.state      # make the oracle start from the float32-rounded state the GPU holds
        st = {k: (v.astype(np.float32).astype(np.float64) if v.dtype.kind == "f" else v) for k, v in st.items()}
        orc.set_state(**st)
        env.set_state(**st)
        a = rng.uniform(-1, 1, (b, 2)).astype(np.float32)
        oo, orew, oterm, otrunc, _ = orc.step(a.astype(np.float64))
        env.set_replay(**orc.draws)
        go, grew, gterm, gtrunc, _ = env.step(torch.as_tensor(a, device=DEV))
        same = (grew.cpu().numpy() == orew) & (gterm.cpu().numpy() == oterm) & (gtrunc.cpu().numpy() == otrunc)
        flips += int((~same).sum())
        total += b
        assert _close(go.cpu().numpy()[same], oo[same]).all(), f"step {t}"
        assert _close(env.agent_yx_velocity.cpu().numpy()[same], orc.velocity[same]).all()
    assert flips / total < F32_FLIP_RATE, flips / total


def test_tag_float32_mode_within_tolerance_and_philox():
    from gym_po.envs import TagVecEnv
    b = 100_000
    orc = oracle.TagOracle(b, time_limit=50, draws=oracle.GeneratorDraws(seed=5))
    env = TagVecEnv(b, time_limit=50, device=DEV, rng_mode="replay", precision="float32")
    orc.reset()
    env.set_replay(**orc.draws)
    env.reset()
    rng = np.random.default_rng(2)
    flips = total = 0
    for t in range(60):
        st = orc.state
        st = {k: (v.astype(np.float32).astype(np.float64) if v.dtype.kind == "f" else v) for k, v in st.items()}
        orc.set_state(**st)
        env.set_state(**st)
        a = np.clip(orc.target - orc.agent, -1, 1).astype(np.float32)
        oo, orew, oterm, otrunc, _ = orc.step(a.astype(np.float64))
        env.set_replay(**orc.draws)
        go, grew, gterm, gtrunc, _ = env.step(torch.as_tensor(a, device=DEV))
        # discrete decisions: tagged, truncated, target visible, target moved (vs stayed at the cage edge / choice 3)
        vis_same = ((go.cpu().numpy() != 0).any(-1)) == ((oo != 0).any(-1))
        done = oterm | otrunc
        moved_g = (env.target_xy.cpu().numpy() != st["target"].astype(np.float32)).any(-1)
        moved_o = (orc.target != st["target"]).any(-1)
        same = (gterm.cpu().numpy() == oterm) & (gtrunc.cpu().numpy() == otrunc) & vis_same & ((moved_g == moved_o) | done)
        flips += int((~same).sum())
        total += b
        assert _close(go.cpu().numpy()[same], oo[same]).all(), f"step {t}"
        assert _close(env.agent_xy.cpu().numpy()[same], orc.agent[same]).all()
        assert _close(env.target_xy.cpu().numpy()[same], orc.target[same]).all()
    assert flips / total < F32_FLIP_RATE, flips / total
    # Philox mode, float32: invariants
    env = TagVecEnv(1 << 20, device=DEV, seed=1, precision="float32")
    obs, _ = env.reset(seed=1)
    assert obs.dtype == torch.float32
    assert bool(((env.agent_xy - env.target_xy).norm(dim=-1) > 5.0).all())
    for _ in range(30):
        obs, rew, term, trunc, _ = env.step(torch.rand((env.capacity, 2), device=DEV) * 2 - 1)
        assert bool((env.agent_xy.abs() <= 5.0).all()) and bool((env.target_xy.abs() <= 4.5).all())

"""GPU: fused multistory FourRooms step vs the oracle / golden fixtures, bit-exact on
replayed draws; Philox-mode invariants at full size."""
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
    if kw.get("goal_xyz", 0) is not None and "goal_xyz" in kw:
        kw["goal_xyz"] = tuple(kw["goal_xyz"])
    return kw


@pytest.mark.parametrize("name", golden_names("msrooms"))
def test_golden_trajectory_free_running(name):
    """Fed the reference's recorded draws, the GPU env reproduces the reference trajectory exactly."""
    from gym_po.envs import MultistoryFourRoomsEnv
    fx = load_golden(name)
    meta = fx["meta"]
    orc = make_oracle(meta, recorded_draws(fx))
    env = MultistoryFourRoomsEnv(meta["B"], device=DEV, rng_mode="replay", **_kw(meta))
    orc.reset()
    env.set_replay(**orc.draws)
    obs, info = env.reset()                          # (obs, {}) like the reference (msrooms.py:383)
    assert info == {}
    np.testing.assert_array_equal(obs.cpu().numpy(), fx["obs0"])
    for t in range(meta["T"]):
        a = fx["actions"][t]
        orc.step(a)
        env.set_replay(**orc.draws)
        g = env.step(torch.as_tensor(a, device=DEV))
        _cmp(g[:4], (fx["obs"][t], fx["rew"][t], fx["term"][t], fx["trunc"][t]), t)
    st = env.get_state()
    np.testing.assert_array_equal(st["agent"].cpu().numpy(), fx["state_agent"])
    np.testing.assert_array_equal(st["goal"].cpu().numpy(), fx["state_goal"])
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), fx["state_elapsed"])


OBS_TYPES = ["room", "room_goal", "mdp", "mdp_goal", "vector_mdp", "vector_mdp_goal", "hansen", "hansen8", "vector_hansen",
             "vector_hansen8", "vector_goal_hansen", "vector_goal_hansen8"]


@pytest.mark.parametrize("obs_type", OBS_TYPES)
@pytest.mark.parametrize("goal_xyz", [(9, 7, -1), None])
@pytest.mark.parametrize("floors", [1, 3])
def test_lockstep_all_obs_variants(obs_type, goal_xyz, floors):
    """All 12 obs variants x fixed/random goal x 1/3 floors, ragged batch, short time limit (many resets)."""
    from gym_po.envs import MultistoryFourRoomsEnv
    b = 3000
    action_type = "cardinal" if obs_type in ("hansen", "vector_hansen", "mdp", "room") else "ordinal"
    kw = dict(grid_z=floors, obs_type=obs_type, goal_xyz=goal_xyz, time_limit=37, action_type=action_type,
              action_failure_probability=0.3, step_reward=-0.01, wall_reward=-0.2, goal_reward=2.0)
    orc = oracle.MSRoomsOracle(b, draws=oracle.GeneratorDraws(seed=5), **kw)
    env = MultistoryFourRoomsEnv(b, device=DEV, rng_mode="replay", **kw)
    o, _ = orc.reset()
    env.set_replay(**orc.draws)
    np.testing.assert_array_equal(env.reset()[0].cpu().numpy(), o)
    rng = np.random.default_rng(3)
    for t in range(150):
        a = rng.integers(orc.n_actions, size=b)
        o = orc.step(a)
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(a, dtype=torch.int8, device=DEV))[:4], o[:4], t)
    st = env.get_state()
    np.testing.assert_array_equal(st["agent"].cpu().numpy(), orc.agent)
    np.testing.assert_array_equal(st["goal"].cpu().numpy(), orc.goal)
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed)


@pytest.mark.parametrize("action_type,n_act", [("ordinal", 8), ("cardinal", 4)])
@pytest.mark.parametrize("obs_type,goal_xyz", [("hansen8", (9, 7, -1)), ("vector_goal_hansen8", None), ("mdp_goal", None)])
def test_exhaustive_cell_x_slipped_action(action_type, n_act, obs_type, goal_xyz):
    """Every walkable cell of a 4-floor building x intended action x every slip outcome (u placed just inside
    each threshold band): wall collisions, both stair teleports, goal hits, respawn."""
    from gym_po.envs import MultistoryFourRoomsEnv
    floors = 4
    grid = oracle.msrooms.multistory_grid(oracle.msrooms.FR_MAP, floors)
    cells = np.stack(np.nonzero(grid > 0), -1)
    P = oracle.rooms.slip_matrix(n_act, 1.0 / 3).cumsum(axis=1)
    cases = [(c, a, j) for c in range(len(cells)) for a in range(n_act) for j in range(n_act)]
    ci, ai, ji = (np.array(v) for v in zip(*cases))
    lo = np.where(ji > 0, P[ai, np.maximum(ji - 1, 0)], 0.0)
    u = np.where(ji == 0, 0.0, np.nextafter(lo, 2.0))
    b = len(cases)

    class Fixed(oracle.GeneratorDraws):
        def random(self, n, **ctx):
            return u.copy()

    kw = dict(grid_z=floors, obs_type=obs_type, action_type=action_type, goal_xyz=goal_xyz)
    orc = oracle.MSRoomsOracle(b, draws=Fixed(seed=0), **kw)
    env = MultistoryFourRoomsEnv(b, device=DEV, rng_mode="replay", **kw)
    orc.reset(); env.set_replay(**orc.draws); env.reset()
    top = cells[cells[:, 0] == floors - 1]
    goals = orc.goal if goal_xyz is not None else top[np.random.default_rng(1).integers(len(top), size=b)]
    st = dict(agent=cells[ci], goal=goals, elapsed=np.zeros(b, dtype=int))
    orc.set_state(**st); env.set_state(**st)
    o = orc.step(ai)
    assert o[2].sum() > 0                                          # some cases do reach the goal
    env.set_replay(**orc.draws)
    _cmp(env.step(torch.as_tensor(ai, dtype=torch.int8, device=DEV))[:4], o[:4], 0)
    np.testing.assert_array_equal(env.agent_zyx.cpu().numpy(), orc.agent)
    np.testing.assert_array_equal(env.goal_zyx.cpu().numpy(), orc.goal)


def test_philox_statistics_and_invariants_full_size():
    """2^22 envs, 3 floors: uniform bottom-floor spawn, slip law 2/3 + 1/9 x 3, never inside a wall, stairs reach the
    upper floors, goal hits pay the goal reward, truncation exactly at elapsed > time_limit."""
    from gym_po.envs import MultistoryFourRoomsEnv
    b = 1 << 22
    env = MultistoryFourRoomsEnv(b, grid_z=3, obs_type="hansen", time_limit=200, device=DEV, seed=3)
    obs, _ = env.reset(seed=3)
    grid = torch.as_tensor(env.grid, device=DEV)
    a0 = env.agent_zyx
    assert bool((a0[:, 0] == 0).all()) and bool((grid[a0[:, 0], a0[:, 1], a0[:, 2]] > 0).all())
    flat = (a0[:, 1] * 13 + a0[:, 2])
    cnt = torch.bincount(flat, minlength=169)[torch.as_tensor(env.valid_agent_states, device=DEV)].cpu().numpy()
    exp = b / len(env.valid_agent_states)
    chi2 = float(((cnt - exp) ** 2 / exp).sum())
    dof = len(env.valid_agent_states) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), chi2
    # slip from an open cell, intended N (cardinal): N w.p. 2/3, E/S/W w.p. 1/9 each
    st = dict(agent=np.tile([0, 3, 3], (b, 1)), goal=None, elapsed=np.zeros(b, dtype=int))
    env.set_state(**st)
    env.step(torch.zeros(env.capacity, dtype=torch.int8, device=DEV))
    d = env.agent_zyx - torch.tensor([0, 3, 3], device=DEV)
    key = ((d[:, 1] + 1) * 3 + (d[:, 2] + 1)).cpu().numpy()
    frac = np.bincount(key, minlength=9) / b
    assert abs(frac[1] - 2 / 3) < 2e-3
    for k in (3, 5, 7):
        assert abs(frac[k] - 1 / 9) < 1e-3, (k, frac[k])
    assert frac[4] == 0.0 and frac[0] == frac[2] == frac[6] == frac[8] == 0.0
    gen = torch.Generator(device=DEV).manual_seed(0)
    seen_floor = torch.zeros(3, dtype=torch.bool, device=DEV)
    n_term = 0
    for t in range(260):
        a = torch.randint(0, 4, (env.capacity,), dtype=torch.int8, device=DEV, generator=gen)
        prev = env.elapsed.clone()
        obs, rew, term, trunc, _ = env.step(a)
        zyx = env.agent_zyx
        assert bool((grid[zyx[:, 0], zyx[:, 1], zyx[:, 2]] > 0).all())
        assert bool((rew[term] == 1.0).all()) and bool((rew[~term] == 0.0).all())
        done = term | trunc
        assert bool((env.elapsed[done] == 0).all()) and bool((zyx[done][:, 0] == 0).all())
        assert bool((env.elapsed[~done] == prev[~done] + 1).all())
        assert bool((trunc == (prev + 1 > 200)).all())
        seen_floor |= torch.bincount(zyx[:, 0], minlength=3) > 0
        n_term += int(term.sum())
        assert int(obs.min()) >= 0 and int(obs.max()) <= 80 * 4
    assert bool(seen_floor.all()) and n_term > 0


def test_gpu_count_independence_and_host_path():
    from gym_po.envs import MultistoryFourRoomsEnv
    b = 1 << 14
    kw = dict(grid_z=2, obs_type="vector_mdp_goal", goal_xyz=None, time_limit=30)
    whole = MultistoryFourRoomsEnv(b, device=DEV, seed=11, **kw)
    lo = MultistoryFourRoomsEnv(b // 2, device=DEV, seed=11, env_offset=0, **kw)
    hi = MultistoryFourRoomsEnv(b // 2, device=DEV, seed=11, env_offset=b // 2, **kw)
    host = MultistoryFourRoomsEnv(b, device=DEV, seed=11, **kw)
    for e in (whole, lo, hi, host):
        e.reset(seed=11)
    rng = np.random.default_rng(0)
    for t in range(80):
        a_np = rng.integers(4, size=b).astype(np.int8)
        a = torch.as_tensor(a_np, device=DEV)
        w = whole.step(a)
        l = lo.step(a[: b // 2].contiguous())
        h = hi.step(a[b // 2:].contiguous())
        hp = host.step_host(a_np)
        for k in range(4):
            assert torch.equal(w[k][: b // 2], l[k]) and torch.equal(w[k][b // 2:], h[k]), (k, t)
            np.testing.assert_array_equal(w[k].cpu().numpy(), hp[k])


def test_constructor_errors_like_the_reference():
    from gym_po.envs import MultistoryFourRoomsEnv
    with pytest.raises(ValueError):
        MultistoryFourRoomsEnv(8, agent_xyz=(1, 1, 0), device=DEV)
    with pytest.raises(NotImplementedError):
        MultistoryFourRoomsEnv(8, obs_type="grid", device=DEV)


@pytest.mark.parametrize("obs_type,goal_xyz,floors", [("mdp", (9, 7, -1), 3), ("hansen8", None, 2), ("vector_mdp_goal", None, 4),
                                                      ("vector_goal_hansen8", (9, 7, -1), 2), ("vector_mdp", (9, 7, -1), 1)])
def test_fused_multi_step_launch_equals_single_steps(obs_type, goal_xyz, floors):
    """gpt_step_many runs T steps in ONE launch; outputs of every step and the final state must be bit-identical to T
    single-step launches."""
    from gym_po.envs import MultistoryFourRoomsEnv
    b, T = 3000, 31
    kw = dict(grid_z=floors, obs_type=obs_type, goal_xyz=goal_xyz, time_limit=19, step_reward=-0.1, wall_reward=-0.5)
    a = MultistoryFourRoomsEnv(b, device=DEV, seed=9, **kw)
    c = MultistoryFourRoomsEnv(b, device=DEV, seed=9, **kw)
    a.reset(seed=9); c.reset(seed=9)
    gen = torch.Generator(device=DEV).manual_seed(4)
    for rep in range(3):
        acts = torch.randint(0, 4, (T, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
        out = {n: torch.zeros((T,) + tuple(a._arrays[n].shape), dtype=a._arrays[n].dtype, device=DEV)
               for n in ("obs", "reward", "terminated", "truncated")}
        l0 = a.launch_count
        a.step_many(acts, out)
        assert a.launch_count == l0 + 1
        for t in range(T):
            o = c.step(acts[t])
            for n, x in zip(("obs", "reward", "terminated", "truncated"), o[:4]):
                assert torch.equal(out[n][t][:b].reshape(x.shape).view(x.dtype), x), (n, rep, t)
        sa, sc = a.get_state(), c.get_state()
        for k in sa:
            assert torch.equal(sa[k], sc[k]), k


def test_custom_floor_map_lockstep():
    """A caller-supplied floor map (the kwarg the reference takes): open hall with pillars, 3 floors, random goal."""
    from gym_po.envs import MultistoryFourRoomsEnv
    fm = np.zeros((13, 15), dtype=np.int64)
    fm[1:12, 1:14] = 1
    fm[3:10:3, 3:12:4] = 0                      # pillars
    fm[6, 1:6] = 0                              # a partial wall
    b = 2000
    kw = dict(grid_z=3, floor_map=fm, obs_type="vector_goal_hansen8", action_type="ordinal", goal_xyz=None, time_limit=120)
    orc = oracle.MSRoomsOracle(b, draws=oracle.GeneratorDraws(seed=8), **kw)
    env = MultistoryFourRoomsEnv(b, device=DEV, rng_mode="replay", **kw)
    o, _ = orc.reset()
    env.set_replay(**orc.draws)
    np.testing.assert_array_equal(env.reset()[0].cpu().numpy(), o)
    rng = np.random.default_rng(3)
    for t in range(200):
        a = rng.integers(8, size=b)
        o = orc.step(a)
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(a, dtype=torch.int8, device=DEV))[:4], o[:4], t)
    np.testing.assert_array_equal(env.agent_zyx.cpu().numpy(), orc.agent)
    assert len(np.unique(orc.agent[:, 0])) >= 2          # the stairs were used

"""GPU: fused ROOMS step vs the oracle / golden fixtures, bit-exact on replayed draws."""
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


@pytest.mark.parametrize("name", golden_names("rooms"))
def test_golden_trajectory_free_running(name):
    """Fed the reference's recorded draws, the GPU env reproduces the reference trajectory exactly."""
    from gym_po.envs import RoomsEnv
    fx = load_golden(name)
    meta = fx["meta"]
    orc = make_oracle(meta, recorded_draws(fx))
    env = RoomsEnv(meta["B"], device=DEV, rng_mode="replay", **_kw(meta))
    orc.reset()
    env.set_replay(**orc.draws)
    obs = env.reset()
    assert isinstance(obs, torch.Tensor)          # reset returns obs only (reference rooms.py:189)
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
             "vector_hansen8", "vector_goal_hansen", "vector_goal_hansen8", "grid"]


@pytest.mark.parametrize("obs_type", OBS_TYPES)
@pytest.mark.parametrize("goal_xy", [(0, 0), None])
def test_lockstep_all_obs_variants(obs_type, goal_xy):
    """All 13 obs variants x fixed/random goal, ragged batch, short time limit (many resets)."""
    from gym_po.envs import RoomsEnv
    b = 3000
    layout = "10b" if goal_xy is None else "4"
    action_type = "cardinal" if obs_type in ("hansen", "vector_hansen", "mdp") else "ordinal"
    kw = dict(layout=layout, obs_type=obs_type, obs_n=5, goal_xy=goal_xy, time_limit=23, action_type=action_type,
              action_failure_probability=0.3, step_reward=-0.01, wall_reward=-0.2, goal_reward=2.0)
    orc = oracle.RoomsOracle(b, draws=oracle.GeneratorDraws(seed=5), **kw)
    env = RoomsEnv(b, device=DEV, rng_mode="replay", **kw)
    o = orc.reset()
    env.set_replay(**orc.draws)
    np.testing.assert_array_equal(env.reset().cpu().numpy(), o)
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


@pytest.mark.parametrize("layout", list(oracle.LAYOUT_NAMES))
@pytest.mark.parametrize("n", [3, 4, 9, 11])
def test_grid_window_every_cell_every_layout(layout, n):
    """Every walkable cell of all 12 layouts as agent position, random goals: window obs incl. map edges."""
    from gym_po.envs import RoomsEnv
    grid = oracle.load_layout(layout)
    cells = np.stack(np.nonzero(grid >= 0), -1)
    b = len(cells)
    rng = np.random.default_rng(1)
    goals = cells[rng.integers(b, size=b)]
    kw = dict(layout=layout, obs_type="grid", obs_n=n, goal_xy=None)
    orc = oracle.RoomsOracle(b, draws=oracle.GeneratorDraws(seed=2), **kw)
    env = RoomsEnv(b, device=DEV, rng_mode="replay", **kw)
    orc.reset()
    env.set_replay(**orc.draws)
    env.reset()
    st = dict(agent=cells, goal=goals, elapsed=np.zeros(b, dtype=int))
    orc.set_state(**st)
    env.set_state(**st)
    for a in range(8):
        act = np.full(b, a)
        o = orc.step(act)
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(act, dtype=torch.int8, device=DEV))[:4], o[:4], a)


@pytest.mark.parametrize("action_type,n_act", [("ordinal", 8), ("cardinal", 4)])
def test_exhaustive_cell_x_slipped_action(action_type, n_act):
    """Every cell x intended action x every slip outcome (u placed just inside each threshold band)."""
    from gym_po.envs import RoomsEnv
    for layout in ("4", "32b"):
        grid = oracle.load_layout(layout)
        cells = np.stack(np.nonzero(grid >= 0), -1)
        P = oracle.rooms.slip_matrix(n_act, 0.2).cumsum(axis=1)
        cases = [(c, a, j) for c in range(len(cells)) for a in range(n_act) for j in range(n_act)]
        ci, ai, ji = (np.array(v) for v in zip(*cases))
        lo = np.where(ji > 0, P[ai, np.maximum(ji - 1, 0)], 0.0)
        u = np.nextafter(lo, 2.0)                       # smallest u with exactly j thresholds below it
        u = np.where(ji == 0, 0.0, u)
        b = len(cases)

        class Fixed(oracle.GeneratorDraws):
            def random(self, n, **ctx):
                return u.copy()

        kw = dict(layout=layout, obs_type="hansen8", action_type=action_type, goal_xy=(0, 0))
        orc = oracle.RoomsOracle(b, draws=Fixed(seed=0), **kw)
        env = RoomsEnv(b, device=DEV, rng_mode="replay", **kw)
        orc.reset(); env.set_replay(**orc.draws); env.reset()
        st = dict(agent=cells[ci], goal=orc.goal, elapsed=np.zeros(b, dtype=int))
        orc.set_state(**st); env.set_state(**st)
        o = orc.step(ai)
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(ai, dtype=torch.int8, device=DEV))[:4], o[:4], 0)
        np.testing.assert_array_equal(env.agent_yx.cpu().numpy(), orc.agent)


def test_philox_statistics_and_invariants_full_size():
    """2^22 envs (BASELINE configs 3/4 per-GPU size): slip rate, wall collisions, goal hits, resets."""
    from gym_po.envs import RoomsEnv
    b = 1 << 22
    env = RoomsEnv(b, "4", obs_type="hansen8", device=DEV, seed=3)
    obs = env.reset(seed=3)
    grid = torch.as_tensor(env.grid, device=DEV)
    a0 = env.agent_yx
    assert bool((grid[a0[:, 0], a0[:, 1]] >= 0).all())
    # spawn is uniform over the 200 valid cells
    cnt = torch.bincount(a0[:, 0] * 17 + a0[:, 1], minlength=289)[torch.as_tensor(env.valid_states, device=DEV)].cpu().numpy()
    exp = b / 200
    chi2 = float(((cnt - exp) ** 2 / exp).sum())
    assert chi2 < 199 + 6 * np.sqrt(2 * 199), chi2
    # slip: intended N from an open cell moves N w.p. 0.8, each other direction w.p. 0.2/7
    st = dict(agent=np.tile([3, 3], (b, 1)), goal=None, elapsed=np.zeros(b, dtype=int))
    env.set_state(**st)
    env.step(torch.zeros(env.capacity, dtype=torch.int8, device=DEV))
    d = (env.agent_yx - torch.tensor([3, 3], device=DEV))
    key = ((d[:, 0] + 1) * 3 + (d[:, 1] + 1)).cpu().numpy()
    frac = np.bincount(key, minlength=9) / b
    assert abs(frac[1] - 0.8) < 2e-3                               # (-1,0) = N
    for k in (0, 2, 3, 5, 6, 7, 8):
        assert abs(frac[k] - 0.2 / 7) < 1e-3, (k, frac[k])
    assert frac[4] == 0.0                                          # never stays in an open area
    gen = torch.Generator(device=DEV).manual_seed(0)
    for t in range(60):
        a = torch.randint(0, 8, (env.capacity,), dtype=torch.int8, device=DEV, generator=gen)
        prev = env.elapsed.clone()
        obs, rew, term, trunc, _ = env.step(a)
        ayx = env.agent_yx
        assert bool((grid[ayx[:, 0], ayx[:, 1]] >= 0).all())       # never inside a wall
        assert bool((rew[term] == 1.0).all()) and bool((rew[~term] == 0.0).all())
        assert bool((env.elapsed[term | trunc] == 0).all())
        assert bool((env.elapsed[~(term | trunc)] == prev[~(term | trunc)] + 1).all())
        assert int(obs.min()) >= 0 and int(obs.max()) <= 2040


def test_gpu_count_independence_and_host_path():
    from gym_po.envs import RoomsEnv
    b = 1 << 14
    kw = dict(layout="8", obs_type="vector_goal_hansen8", goal_xy=None, time_limit=30)
    whole = RoomsEnv(b, device=DEV, seed=11, **kw)
    lo = RoomsEnv(b // 2, device=DEV, seed=11, env_offset=0, **kw)
    hi = RoomsEnv(b // 2, device=DEV, seed=11, env_offset=b // 2, **kw)
    host = RoomsEnv(b, device=DEV, seed=11, **kw)
    for e in (whole, lo, hi, host):
        e.reset(seed=11)
    rng = np.random.default_rng(0)
    for t in range(80):
        a_np = rng.integers(8, size=b).astype(np.int8)
        a = torch.as_tensor(a_np, device=DEV)
        w = whole.step(a)
        l = lo.step(a[: b // 2].contiguous())
        h = hi.step(a[b // 2:].contiguous())
        hp = host.step_host(a_np)
        for k in range(4):
            assert torch.equal(w[k][: b // 2], l[k]) and torch.equal(w[k][b // 2:], h[k]), (k, t)
            np.testing.assert_array_equal(w[k].cpu().numpy(), hp[k])


@pytest.mark.parametrize("obs_type,goal_xy", [("hansen8", (0, 0)), ("grid", None)])
def test_episode_statistics_match_returned_rewards(obs_type, goal_xy):
    """track_stats=1 (Philox mode): the on-device statistics equal what the host accumulates from the
    returned reward / terminated / truncated tensors (padding rows excluded, reset() not counted)."""
    from gym_po.envs import RoomsEnv
    from gym_po.sharding import allreduce_stats
    b = 5000
    env = RoomsEnv(b, "2", obs_type=obs_type, obs_n=5, goal_xy=goal_xy, time_limit=17, step_reward=-0.25, wall_reward=-1.0,
                   goal_reward=4.0, device=DEV, seed=4, track_stats=True)
    env.reset(seed=4)
    ret = torch.zeros(b, device=DEV)
    length = torch.zeros(b, dtype=torch.int64, device=DEV)
    tot = torch.zeros(5, dtype=torch.float64, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(2)
    for t in range(120):
        a = torch.randint(0, 8, (env.capacity,), dtype=torch.int8, device=DEV, generator=gen)
        obs, rew, term, trunc, _ = env.step(a)
        ret += rew
        length += 1
        done = term | trunc
        r = ret[done].double()
        tot += torch.stack([done.sum().double(), r.sum(), length[done].sum().double(), (r * r).sum(),
                            torch.tensor(float(b), dtype=torch.float64, device=DEV)])
        ret[done] = 0
        length[done] = 0
    got = env.stats_tensor().cpu().numpy()
    tot = tot.cpu().numpy()
    assert tot[0] > 1000
    assert got[0] == tot[0] and got[2] == tot[2] and got[4] == tot[4]
    np.testing.assert_allclose(got[1], tot[1], rtol=1e-5)
    np.testing.assert_allclose(got[3], tot[3], rtol=1e-5)
    assert abs(allreduce_stats(env.stats_tensor().clone())["mean_return"] - tot[1] / tot[0]) < 1e-4
    with pytest.raises(ValueError):
        from gym_po.envs import CRoomsEnv
        CRoomsEnv(64, device=DEV, track_stats=True)


@pytest.mark.parametrize("obs_type,goal_xy,layout", [("hansen8", (0, 0), "4"), ("grid", None, "8"), ("vector_goal_hansen8", None, "10b"),
                                                     ("grid", (0, 0), "32"), ("mdp_goal", None, "2"), ("vector_mdp", (0, 0), "16")])
def test_fused_multi_step_launch_equals_single_steps(obs_type, goal_xy, layout):
    """gpt_step_many on ROOMS runs T steps in ONE launch (pos / goal / elapsed in registers); every step's outputs and the
    final state must be bit-identical to T single-step launches (Philox counters = (global env / quad id, step))."""
    from gym_po.envs import RoomsEnv
    b, T = 3000, 29
    kw = dict(layout=layout, obs_type=obs_type, obs_n=5, goal_xy=goal_xy, time_limit=11, step_reward=-0.1, wall_reward=-0.5)
    a = RoomsEnv(b, device=DEV, seed=9, **kw)
    c = RoomsEnv(b, device=DEV, seed=9, **kw)
    a.reset(seed=9); c.reset(seed=9)
    gen = torch.Generator(device=DEV).manual_seed(4)
    for rep in range(3):
        acts = torch.randint(0, 8, (T, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
        out = {n: torch.zeros((T,) + tuple(a._arrays[n].shape), dtype=a._arrays[n].dtype, device=DEV)
               for n in ("obs", "reward", "terminated", "truncated")}
        l0 = a.launch_count
        a.step_many(acts, out)
        assert a.launch_count == l0 + 1      # one fused launch
        for t in range(T):
            o = c.step(acts[t])
            for n, x in zip(("obs", "reward", "terminated", "truncated"), o[:4]):
                assert torch.equal(out[n][t][:b].reshape(x.shape).view(x.dtype), x), (n, rep, t)
        sa, sc = a.get_state(), c.get_state()
        for k in sa:
            assert torch.equal(sa[k], sc[k]), k
        assert a.rng_counter == c.rng_counter


@pytest.mark.parametrize("obs_type,goal_xy,layout,cardinal", [("hansen8", (0, 0), "4", False), ("vector_hansen8", (0, 0), "4", False),
                                                              ("vector_goal_hansen4", None, "8", True), ("vector_mdp", (0, 0), "16", False),
                                                              ("vector_mdp_goal", None, "2", True), ("room_goal", None, "10b", False),
                                                              ("grid", (0, 0), "4", False), ("grid", None, "8", True)])
def test_fused_launch_odd_tiles_padded_rows_in_place(obs_type, goal_xy, layout, cardinal):
    """Fused launches at an ODD number of 512-env tiles (half-filled last CTA), rollout storage with padded rows and
    in-place outputs — all equal to single-step launches, for every observation width (2, 4, 8 bytes per env) and the
    window observation (TMA tile store into rollout slots, per-lane stores for the in-place outputs)."""
    from gym_po.envs import RoomsEnv
    b, T = 1400, 23
    kw = dict(layout=layout, obs_type=obs_type, goal_xy=goal_xy, time_limit=9, step_reward=-0.1, wall_reward=-0.5)
    if cardinal:
        kw["action_type"] = "cardinal"
    a = RoomsEnv(b, device=DEV, seed=5, **kw)
    c = RoomsEnv(b, device=DEV, seed=5, **kw)
    assert (a.capacity // 512) % 2 == 1
    a.reset(seed=5); c.reset(seed=5)
    gen = torch.Generator(device=DEV).manual_seed(8)
    names = ("obs", "reward", "terminated", "truncated")
    n_act = 4 if cardinal else 8
    for pad in (0, 16 * 3):
        rows = a.capacity + pad
        acts = torch.randint(0, n_act, (T, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
        out = {n: torch.full((T, rows) + tuple(a._arrays[n].shape[1:]), 77, dtype=a._arrays[n].dtype, device=DEV) for n in names}
        l0 = a.launch_count
        a.step_many(acts, out)
        assert a.launch_count == l0 + 1
        for t in range(T):
            o = c.step(acts[t])
            for n, x in zip(names, o[:4]):
                assert torch.equal(out[n][t][:b].reshape(x.shape).view(x.dtype), x), (n, pad, t)
            for n in names:
                assert bool((out[n][t][a.capacity:] == 77).all()), (n, t)
        sa, sc = a.get_state(), c.get_state()
        for k in sa:
            assert torch.equal(sa[k], sc[k]), k
    acts = torch.randint(0, n_act, (7, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
    a.step_many(acts)                       # in place: the bound arrays hold the last step's results
    for t in range(7):
        o = c.step(acts[t])
    for n, x in zip(names, o[:4]):
        assert torch.equal(a._arrays[n][:b].reshape(x.shape).view(x.dtype), x), n


def test_step_many_with_a_huge_time_limit_steps_one_launch_at_a_time():
    """The fused kernels keep the elapsed counters as biased 16-bit pairs: a time limit beyond 32766 is not fused —
    `step_many` then runs T single-step launches with the same results."""
    from gym_po.envs import RoomsEnv
    b, T = 2000, 6
    a = RoomsEnv(b, device=DEV, seed=2, obs_type="hansen8", time_limit=40000)
    c = RoomsEnv(b, device=DEV, seed=2, obs_type="hansen8", time_limit=40000)
    a.reset(seed=2); c.reset(seed=2)
    el = torch.randint(32760, 40001, (b,), device=DEV, dtype=torch.int32)
    a.elapsed.copy_(el)
    c.elapsed.copy_(el)
    acts = torch.randint(0, 8, (T, a.capacity), dtype=torch.int8, device=DEV)
    out = {n: torch.zeros((T,) + tuple(a._arrays[n].shape), dtype=a._arrays[n].dtype, device=DEV)
           for n in ("obs", "reward", "terminated", "truncated")}
    l0 = a.launch_count
    a.step_many(acts, out)
    assert a.launch_count == l0 + T
    for t in range(T):
        o = c.step(acts[t])
        for n, x in zip(("obs", "reward", "terminated", "truncated"), o[:4]):
            assert torch.equal(out[n][t][:b].reshape(x.shape).view(x.dtype), x), (n, t)
    assert torch.equal(a.elapsed, c.elapsed)
    # at the largest fused limit the packed counters still truncate exactly at elapsed > time_limit
    a = RoomsEnv(b, device=DEV, seed=3, obs_type="hansen8", time_limit=32766)
    c = RoomsEnv(b, device=DEV, seed=3, obs_type="hansen8", time_limit=32766)
    a.reset(seed=3); c.reset(seed=3)
    el = torch.randint(32760, 32767, (b,), device=DEV, dtype=torch.int32)
    a.elapsed.copy_(el)
    c.elapsed.copy_(el)
    l0 = a.launch_count
    a.step_many(acts, out)
    assert a.launch_count == l0 + 1
    seen_trunc = False
    for t in range(T):
        o = c.step(acts[t])
        seen_trunc |= bool(o[3].any())
        for n, x in zip(("obs", "reward", "terminated", "truncated"), o[:4]):
            assert torch.equal(out[n][t][:b].reshape(x.shape).view(x.dtype), x), (n, t)
    assert seen_trunc
    assert torch.equal(a.elapsed, c.elapsed)

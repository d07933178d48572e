"""GPU: the kernels that are actually benchmarked — Philox mode, single-step AND fused multi-step launches —
compared DIRECTLY with the oracle.

The oracle (numpy restatement of the reference step) is driven by ``PhiloxDraws`` (tests/philox_host.py), which
recomputes on the host the draw every env consumes at every Philox step counter: reset state through the reset alias
table, slip direction through the slip alias rows, spawn cells through multiply-high.  Everything the step returns
(obs, reward, terminated, truncated) and the final state must match bit for bit.  Reference semantics:
extended_taxi.py:244-287, :344-364; rooms/rooms.py:191-222; rooms/action_utils.py:84-90; rooms/msrooms.py:385-428.
"""
import numpy as np
import pytest
import torch

import oracle
from philox_host import PhiloxDraws

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ("obs", "reward", "terminated", "truncated")


def _cmp(g, o, what):
    for name, x, y in zip(NAMES, g, o):
        x = x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        np.testing.assert_array_equal(x.astype(np.float64).reshape(np.shape(y)), np.asarray(y, dtype=np.float64),
                                      err_msg=f"{name} at {what}")


def _first(o):
    return o[0] if isinstance(o, tuple) else o


def _run(env, orc, draws, n_actions, b, *, single_steps, fused_T, fused_launches, seed, state_cmp):
    """reset, `single_steps` gpt_step launches, then `fused_launches` gpt_step_many launches of `fused_T` steps with
    rollout slots — the oracle stepped beside every one of them on host-rebuilt Philox draws."""
    rng = np.random.default_rng(seed)
    cap = env.capacity
    draws.counter = env.rng_counter
    o0 = _first(orc.reset())
    g0 = _first(env.reset())
    np.testing.assert_array_equal(g0.cpu().numpy().astype(np.float64).reshape(np.shape(o0)), np.asarray(o0, dtype=np.float64), err_msg="reset obs")
    for t in range(single_steps):
        a = rng.integers(n_actions, size=b)
        draws.counter = env.rng_counter
        o = orc.step(a)
        g = env.step(torch.as_tensor(a, dtype=torch.int8, device=DEV))
        _cmp(g[:4], o[:4], f"single step {t}")
    state_cmp("after the single steps")
    for L in range(fused_launches):
        acts = np.zeros((fused_T, cap), dtype=np.int8)
        acts[:, :b] = rng.integers(n_actions, size=(fused_T, b))
        out = {}
        for nm in NAMES:
            arr = env._arrays[nm]
            out[nm] = torch.zeros((fused_T,) + tuple(arr.shape), dtype=arr.dtype, device=DEV)
        c0, l0 = env.rng_counter, env.launch_count
        env.step_many(torch.as_tensor(acts, device=DEV), out)
        assert env.launch_count - l0 == 1, "step_many must run as ONE fused launch here"
        assert env.rng_counter == c0 + fused_T
        host = {nm: out[nm][:, :b].cpu().numpy() for nm in NAMES}
        for t in range(fused_T):
            draws.counter = c0 + t
            o = orc.step(acts[t, :b])
            _cmp([host[nm][t] for nm in NAMES], o[:4], f"fused launch {L} step {t}")
        state_cmp(f"after fused launch {L}")


# ------------------------------------------------------------------------------------------------ Taxi
@pytest.mark.parametrize("kwargs", [
    {}, {"hansen_obs": True}, {"num_passengers": 3, "time_limit": 60},
    {"map": oracle.EXTENDED_TAXI_MAP, "time_limit": 37}, {"map": oracle.EXTENDED_TAXI_MAP, "num_passengers": 3, "hansen_obs": True, "time_limit": 90},
])
def test_taxi_philox_kernels_vs_oracle(kwargs):
    from gym_po.envs import TaxiVecEnv
    b, seed, off = 20_011, 0x1234_5678_9ABC_DEF0, 512 * 12_345_679   # a global env id beyond 2^32 exercises the high counter word
    env = TaxiVecEnv(b, device=DEV, seed=seed, env_offset=off, **kwargs)
    draws = PhiloxDraws(env, seed, "taxi", env_offset=off)
    orc = oracle.TaxiOracle(b, draws=draws, **kwargs)

    def state_cmp(what):
        st = env.get_state()
        np.testing.assert_array_equal(st["s"].cpu().numpy(), orc.s, err_msg=f"s {what}")
        np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed, err_msg=f"elapsed {what}")
        np.testing.assert_array_equal(st["ndrop"].cpu().numpy(), orc.ndrop, err_msg=f"ndrop {what}")

    # time limits are short: the synchronised truncation of every env (a full-batch reset) falls inside both legs
    _run(env, orc, draws, 5, b, single_steps=100, fused_T=8, fused_launches=12, seed=4, state_cmp=state_cmp)


def test_taxi_respawn_is_uniform_over_other_locations():
    """The kernel's `d += d >= p` respawn (no rejection loop) never clashes and, through the host-rebuilt draws, gives
    every (p, d != p) pair — same law as the reference's redraw loop (extended_taxi.py:360-363)."""
    from gym_po.envs import TaxiVecEnv
    b, seed = 1 << 15, 99
    env = TaxiVecEnv(b, device=DEV, seed=seed, num_passengers=2)
    draws = PhiloxDraws(env, seed, "taxi")
    draws.counter = 5
    every = np.ones(b, dtype=bool)
    p = draws.integers(4, size=b, where=every, kind="new_p")
    d = draws.integers(4, size=b, where=every, kind="new_d")
    assert (p != d).all() and p.min() == 0 and p.max() == 3 and d.min() == 0 and d.max() == 3
    counts = np.bincount(p * 4 + d, minlength=16).reshape(4, 4)
    assert (np.diag(counts) == 0).all()
    off = counts[~np.eye(4, dtype=bool)]
    exp = b / 12
    assert ((off - exp) ** 2 / exp).sum() < 40.0    # chi-square, 11 dof: p < 1e-4 at 40


# ------------------------------------------------------------------------------------------------ ROOMS
@pytest.mark.parametrize("kw", [
    dict(layout="4", obs_type="hansen8", action_type="ordinal", goal_xy=(0, 0), time_limit=40),
    dict(layout="4", obs_type="hansen", action_type="cardinal", goal_xy=(0, 0), time_limit=40),
    dict(layout="4", obs_type="hansen8", action_type="ordinal", goal_xy=None, time_limit=25),
    dict(layout="10b", obs_type="vector_goal_hansen8", action_type="cardinal", goal_xy=None, time_limit=25),
    dict(layout="4", obs_type="grid", obs_n=5, action_type="ordinal", goal_xy=(0, 0), time_limit=30),
    dict(layout="4", obs_type="grid", obs_n=9, action_type="ordinal", goal_xy=None, time_limit=30),
    dict(layout="16", obs_type="mdp", action_type="cardinal", goal_xy=(0, 0), time_limit=50, action_failure_probability=0.35),
    dict(layout="2", obs_type="vector_mdp_goal", action_type="ordinal", goal_xy=None, time_limit=20),
])
def test_rooms_philox_kernels_vs_oracle(kw):
    from gym_po.envs import RoomsEnv
    b, seed, off = 6_007, 0xFEDC_BA98_7654_3210, 512 * 9_000_001
    kw = dict(kw, step_reward=-0.01, wall_reward=-0.2, goal_reward=2.0)
    env = RoomsEnv(b, device=DEV, seed=seed, env_offset=off, **kw)
    draws = PhiloxDraws(env, seed, "rooms", env_offset=off)
    orc = oracle.RoomsOracle(b, draws=draws, **kw)

    def state_cmp(what):
        st = env.get_state()
        np.testing.assert_array_equal(st["agent"].cpu().numpy(), orc.agent, err_msg=f"agent {what}")
        np.testing.assert_array_equal(st["goal"].cpu().numpy(), orc.goal, err_msg=f"goal {what}")
        np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed, err_msg=f"elapsed {what}")

    _run(env, orc, draws, orc.n_actions, b, single_steps=60, fused_T=8, fused_launches=8, seed=6, state_cmp=state_cmp)


# ------------------------------------------------------------------------------------------------ MSROOMS
@pytest.mark.parametrize("kw", [
    dict(grid_z=3, obs_type="mdp", action_type="cardinal", time_limit=60),
    dict(grid_z=2, obs_type="hansen8", action_type="ordinal", goal_xyz=None, time_limit=40),
    dict(grid_z=3, obs_type="vector_goal_hansen", action_type="cardinal", goal_xyz=None, time_limit=45),
    dict(grid_z=1, obs_type="vector_mdp", action_type="ordinal", time_limit=30),
])
def test_msrooms_philox_kernels_vs_oracle(kw):
    from gym_po.envs import MultistoryFourRoomsEnv
    b, seed, off = 5_003, 77, 512 * 8_388_609
    env = MultistoryFourRoomsEnv(b, device=DEV, seed=seed, env_offset=off, **kw)
    draws = PhiloxDraws(env, seed, "msrooms", env_offset=off)
    orc = oracle.MSRoomsOracle(b, draws=draws, **kw)

    def state_cmp(what):
        st = env.get_state()
        np.testing.assert_array_equal(st["agent"].cpu().numpy(), orc.agent, err_msg=f"agent {what}")
        np.testing.assert_array_equal(st["goal"].cpu().numpy(), orc.goal, err_msg=f"goal {what}")
        np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed, err_msg=f"elapsed {what}")

    _run(env, orc, draws, orc.n_actions, b, single_steps=70, fused_T=8, fused_launches=8, seed=8, state_cmp=state_cmp)


# ------------------------------------------------------------------------------------------------ BASELINE sizes
class _CheapFullBatchResets(oracle.GeneratorDraws):
    """Replay mode consumes whatever was recorded: for the two full-batch resets of the 2^22-env test the recorded
    reset states come from a uniform pick over the valid states instead of 4M x 500 multinomial counts (~40 s each on
    the host); the partial resets inside the 40 steps use the reference's multinomial call."""

    def multinomial_argmax(self, n, pvals, b, **ctx):
        if b < (1 << 18):
            return super().multinomial_argmax(n, pvals, b, **ctx)
        return self.gen.choice(np.flatnonzero(pvals), b)


def test_taxi_baseline_size_replay():
    """BASELINE.json configs[1]: Taxi at 2^22 envs, bit-exact vs the oracle on replayed draws — 40 steps from
    de-synchronised episode phases (so resets keep occurring) plus the step on which EVERY env truncates."""
    from gym_po.envs import TaxiVecEnv
    b = 1 << 22
    orc = oracle.TaxiOracle(b, draws=_CheapFullBatchResets(seed=21))
    env = TaxiVecEnv(b, device=DEV, rng_mode="replay")
    rng = np.random.default_rng(5)
    o_obs, _ = orc.reset()
    env.set_replay(**orc.draws)
    g_obs, _ = env.reset()
    np.testing.assert_array_equal(g_obs.cpu().numpy(), o_obs)
    el = rng.integers(0, 201, size=b)
    orc.elapsed[:] = el
    env.elapsed.copy_(torch.as_tensor(el, dtype=torch.int32))
    n_reset = 0
    for t in range(40):
        a = rng.integers(5, size=b).astype(np.int8)
        o = orc.step(a)
        n_reset += int((o[2] | o[3]).sum())
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(a, device=DEV))[:4], o[:4], f"step {t}")
    assert n_reset > 100_000
    orc.elapsed[:] = 200
    env.elapsed.fill_(200)
    a = rng.integers(5, size=b).astype(np.int8)
    o = orc.step(a)
    assert o[3].all()                                   # the mass-truncation step: a full-batch autoreset
    env.set_replay(**orc.draws)
    _cmp(env.step(torch.as_tensor(a, device=DEV))[:4], o[:4], "mass truncation")
    st = env.get_state()
    np.testing.assert_array_equal(st["s"].cpu().numpy(), orc.s)
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed)


def test_taxi_baseline_size_philox_fused():
    """The timed configuration itself: Taxi, 2^22 envs, Philox mode, fused launches (TMA I/O) of 20 steps — what
    bench.py times — and of 10 steps with rollout slots, de-synchronised phases — against the oracle on host-rebuilt draws."""
    from gym_po.envs import TaxiVecEnv
    b, seed, T = 1 << 22, 0, 20
    env = TaxiVecEnv(b, device=DEV, seed=seed)
    draws = PhiloxDraws(env, seed, "taxi")
    orc = oracle.TaxiOracle(b, draws=draws)
    rng = np.random.default_rng(7)
    draws.counter = env.rng_counter
    o_obs, _ = orc.reset()
    g_obs, _ = env.reset()
    np.testing.assert_array_equal(g_obs.cpu().numpy(), o_obs)
    el = rng.integers(0, 201, size=b)
    orc.elapsed[:] = el
    env.elapsed.copy_(torch.as_tensor(el, dtype=torch.int32))
    out = {nm: torch.zeros((T,) + tuple(env._arrays[nm].shape), dtype=env._arrays[nm].dtype, device=DEV) for nm in NAMES}
    for L, T in enumerate((T, T // 2)):
        acts = rng.integers(5, size=(T, b)).astype(np.int8)
        c0, l0 = env.rng_counter, env.launch_count
        env.step_many(torch.as_tensor(acts, device=DEV), out)
        assert env.launch_count - l0 == 1
        host = {nm: out[nm][:T].cpu().numpy() for nm in NAMES}
        for t in range(T):
            draws.counter = c0 + t
            o = orc.step(acts[t])
            _cmp([host[nm][t] for nm in NAMES], o[:4], f"launch {L} step {t}")
    st = env.get_state()
    np.testing.assert_array_equal(st["s"].cpu().numpy(), orc.s)
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed)


def test_rooms_hansen8_baseline_size_replay():
    """BASELINE.json configs[2] per-GPU size: FourRooms hansen8 + 0.2 slip at 2^21 envs, replayed draws, 30 steps
    from de-synchronised phases plus the step on which every env truncates."""
    from gym_po.envs import RoomsEnv
    b = 1 << 21
    kw = dict(layout="4", obs_type="hansen8")
    orc = oracle.RoomsOracle(b, draws=oracle.GeneratorDraws(seed=22), **kw)
    env = RoomsEnv(b, device=DEV, rng_mode="replay", **kw)
    rng = np.random.default_rng(9)
    o = orc.reset()
    env.set_replay(**orc.draws)
    np.testing.assert_array_equal(env.reset().cpu().numpy(), o)
    el = rng.integers(0, 501, size=b)
    orc.elapsed[:] = el
    env.elapsed.copy_(torch.as_tensor(el, dtype=torch.int32))
    for t in range(30):
        a = rng.integers(8, size=b).astype(np.int8)
        o = orc.step(a)
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(a, device=DEV))[:4], o[:4], f"step {t}")
    orc.elapsed[:] = 500
    env.elapsed.fill_(500)
    a = rng.integers(8, size=b).astype(np.int8)
    o = orc.step(a)
    assert o[3].all()
    env.set_replay(**orc.draws)
    _cmp(env.step(torch.as_tensor(a, device=DEV))[:4], o[:4], "mass truncation")
    st = env.get_state()
    np.testing.assert_array_equal(st["agent"].cpu().numpy(), orc.agent)
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed)


def test_rooms_hansen8_baseline_size_philox_fused():
    """FourRooms hansen8 at 2^21 envs, Philox mode, fused launches of 10 steps, against the oracle."""
    from gym_po.envs import RoomsEnv
    b, seed, T = 1 << 21, 3, 10
    kw = dict(layout="4", obs_type="hansen8")
    env = RoomsEnv(b, device=DEV, seed=seed, **kw)
    draws = PhiloxDraws(env, seed, "rooms")
    orc = oracle.RoomsOracle(b, draws=draws, **kw)
    rng = np.random.default_rng(11)
    draws.counter = env.rng_counter
    o = orc.reset()
    np.testing.assert_array_equal(env.reset().cpu().numpy(), o)
    el = rng.integers(0, 501, size=b)
    orc.elapsed[:] = el
    env.elapsed.copy_(torch.as_tensor(el, dtype=torch.int32))
    out = {nm: torch.zeros((T,) + tuple(env._arrays[nm].shape), dtype=env._arrays[nm].dtype, device=DEV) for nm in NAMES}
    for L in range(2):
        acts = rng.integers(8, size=(T, b)).astype(np.int8)
        c0 = env.rng_counter
        env.step_many(torch.as_tensor(acts, device=DEV), out)
        host = {nm: out[nm].cpu().numpy() for nm in NAMES}
        for t in range(T):
            draws.counter = c0 + t
            o = orc.step(acts[t])
            _cmp([host[nm][t] for nm in NAMES], o[:4], f"launch {L} step {t}")
    st = env.get_state()
    np.testing.assert_array_equal(st["agent"].cpu().numpy(), orc.agent)
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed)

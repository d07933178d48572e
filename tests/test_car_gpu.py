"""GPU: fused car-flag step vs the oracle / golden fixtures, bit-exact on replayed draws."""
import numpy as np
import pytest
import torch

import oracle
from helpers import golden_names, load_golden, make_oracle, recorded_draws

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _env(meta, b, **extra):
    from gym_po.envs import CarVecEnv, DiscreteActionCarVecEnv
    kw = dict(meta["kwargs"])
    if meta["cls"] == "DiscreteActionCarVecEnv":
        return DiscreteActionCarVecEnv(kw.pop("num_actions"), b, device=DEV, **kw, **extra)
    return CarVecEnv(b, device=DEV, **kw, **extra)


def _cmp(g, o, t):
    for name, x, y in zip(("obs", "reward", "terminated", "truncated"), g, o):
        np.testing.assert_array_equal(x.cpu().numpy(), np.asarray(y), err_msg=f"{name} at step {t}")


@pytest.mark.parametrize("name", golden_names("car"))
def test_golden_trajectory_free_running(name):
    fx = load_golden(name)
    meta = fx["meta"]
    adt = torch.float64 if meta["n_act"] == -2 else torch.float32
    orc = make_oracle(meta, recorded_draws(fx))
    env = _env(meta, meta["B"], rng_mode="replay", action_dtype=adt)
    orc.reset()
    env.set_replay(**orc.draws)
    obs, info = env.reset()
    assert info == {} and obs.dtype == torch.float32 and tuple(obs.shape) == (meta["B"], 3)
    np.testing.assert_array_equal(obs.cpu().numpy(), fx["obs0"])
    for t in range(meta["T"]):
        a = fx["actions"][t]
        orc.step(a)
        env.set_replay(**orc.draws)
        g = env.step(torch.as_tensor(a, device=DEV))
        _cmp(g[:4], (fx["obs"][t], fx["rew"][t], fx["term"][t], fx["trunc"][t]), t)
    st = env.get_state()
    np.testing.assert_array_equal(st["s"].cpu().numpy(), fx["state_s"])
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), fx["state_elapsed"])
    np.testing.assert_array_equal(st["heavens"].cpu().numpy(), fx["state_heavens"])
    np.testing.assert_array_equal(st["priests"].cpu().numpy(), fx["state_priests"])


@pytest.mark.parametrize("mode", ["f32", "f64", "discrete7"])
@pytest.mark.parametrize("b", [1, 777, 30_000])
def test_lockstep_vs_oracle(mode, b):
    from gym_po.envs import CarVecEnv, DiscreteActionCarVecEnv
    if mode == "discrete7":
        orc = oracle.CarOracle(b, time_limit=110, num_actions=7, draws=oracle.GeneratorDraws(seed=4))
        env = DiscreteActionCarVecEnv(7, b, time_limit=110, device=DEV, rng_mode="replay")
    else:
        dt = np.float32 if mode == "f32" else np.float64
        orc = oracle.CarOracle(b, time_limit=110, draws=oracle.GeneratorDraws(seed=4))
        env = CarVecEnv(b, time_limit=110, device=DEV, rng_mode="replay", action_dtype=torch.float32 if mode == "f32" else torch.float64)
    rng = np.random.default_rng(8)

    def act(t):
        if mode == "discrete7":
            a = rng.integers(7, size=b)
            return np.where(orc.s[:, 0] >= 0, 6, 0) if t % 3 else a
        a = rng.uniform(-1.5, 1.5, (b, 1))
        if t % 3:
            a = np.sign(orc.s[:, :1] + 1e-3) * np.abs(a)
        return a.astype(dt)

    a = act(0)                              # step before reset is legal (state zeros, heavens +1, priests +0.5)
    o = orc.step(a)
    env.set_replay(**orc.draws)
    _cmp(env.step(torch.as_tensor(a, device=DEV))[:4], o[:4], -1)
    o_obs, _ = orc.reset()
    env.set_replay(**orc.draws)
    g_obs, _ = env.reset()
    np.testing.assert_array_equal(g_obs.cpu().numpy(), o_obs)
    n_term = 0
    for t in range(260):
        a = act(t)
        o = orc.step(a)
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(a, device=DEV))[:4], o[:4], t)
        n_term += int(o[2].sum())
    if b > 100:
        assert n_term > 0
    np.testing.assert_array_equal(env.heavens.cpu().numpy(), orc.heavens)
    np.testing.assert_array_equal(env.priests.cpu().numpy(), orc.priests)
    np.testing.assert_array_equal(env.elapsed.cpu().numpy(), orc.elapsed)


def test_philox_invariants_and_host_path():
    from gym_po.envs import CarVecEnv
    b = 1 << 20
    env = CarVecEnv(b, device=DEV, seed=2)
    obs, _ = env.reset(seed=2)
    assert bool((obs[:, 0].abs() <= 0.2).all()) and bool((obs[:, 1:] == 0).all())
    assert abs(float(obs[:, 0].mean())) < 1e-3 and abs(float(obs[:, 0].std()) - 0.4 / np.sqrt(12)) < 1e-3
    assert abs(float((env.heavens > 0).float().mean()) - 0.5) < 3e-3
    assert abs(float((env.priests > 0).float().mean()) - 0.5) < 3e-3
    host = CarVecEnv(4096, device=DEV, seed=5)
    dev = CarVecEnv(4096, device=DEV, seed=5)
    host.reset(seed=5)
    dev.reset(seed=5)
    rng = np.random.default_rng(1)
    for t in range(200):
        a = rng.uniform(-1, 1, (env.capacity,)).astype(np.float32)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(a, device=DEV))
        assert bool((obs[:, 0].abs() <= 1.1).all()) and bool((obs[:, 1].abs() <= 0.07 + 1e-7).all())
        assert bool(((obs[:, 2] == 0) | (obs[:, 2].abs() == 1)).all())
        assert bool((rew[~term] == 0).all()) and bool((rew[term].abs() == 1).all())
        h = host.step_host(a[:4096])
        d = dev.step(torch.as_tensor(a[:4096], device=DEV))
        for k in range(4):
            np.testing.assert_array_equal(d[k].cpu().numpy(), h[k])

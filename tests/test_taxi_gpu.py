"""GPU: fused Taxi step (C ABI -> sm_100a kernel) vs the oracle, bit-exact on replayed draws."""
import numpy as np
import pytest
import torch

import oracle
from helpers import golden_names, load_golden, make_oracle, recorded_draws

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _taxi_env(meta, num_envs=None, **extra):
    from gym_po.envs import (ExtendedHansenTaxiVecEnv, ExtendedTaxiVecEnv, HansenTaxiVecEnv, TaxiVecEnv)
    cls = {"TaxiVecEnv": TaxiVecEnv, "HansenTaxiVecEnv": HansenTaxiVecEnv, "ExtendedTaxiVecEnv": ExtendedTaxiVecEnv,
           "ExtendedHansenTaxiVecEnv": ExtendedHansenTaxiVecEnv}[meta["cls"]]
    return cls(num_envs or meta["B"], device=DEV, **meta["kwargs"], **extra)


def _cmp(g, o, t, what=("obs", "reward", "terminated", "truncated")):
    for name, x, y in zip(what, g, o):
        np.testing.assert_array_equal(x.cpu().numpy(), np.asarray(y), err_msg=f"{name} at step {t}")


@pytest.mark.parametrize("name", golden_names("taxi"))
def test_golden_trajectory_free_running(name):
    """The GPU env, fed the reference's own recorded draws, reproduces the reference trajectory
    stored in tests/golden (obs, reward, terminated, truncated, final state) with no re-injection."""
    fx = load_golden(name)
    meta = fx["meta"]
    orc = make_oracle(meta, recorded_draws(fx))       # used only to scatter the draws per env
    env = _taxi_env(meta, rng_mode="replay")
    orc.reset()
    env.set_replay(**orc.draws)
    obs, info = env.reset()
    assert info == {}
    np.testing.assert_array_equal(obs.cpu().numpy(), fx["obs0"])
    for t in range(meta["T"]):
        a = fx["actions"][t]
        orc.step(a)
        env.set_replay(**orc.draws)
        g = env.step(torch.as_tensor(a, device=DEV))
        _cmp(g[:4], (fx["obs"][t], fx["rew"][t], fx["term"][t], fx["trunc"][t]), t)
        assert g[1].dtype == torch.float32 and g[2].dtype == torch.bool and g[3].dtype == torch.bool
    st = env.get_state()
    np.testing.assert_array_equal(st["s"].cpu().numpy(), fx["state_s"])
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), fx["state_elapsed"])
    np.testing.assert_array_equal(st["ndrop"].cpu().numpy(), fx["state_ndrop"].astype(np.int64))


@pytest.mark.parametrize("kwargs", [
    {}, {"hansen_obs": True}, {"map": oracle.EXTENDED_TAXI_MAP, "num_passengers": 4, "time_limit": 37},
    {"map": oracle.EXTENDED_TAXI_MAP, "hansen_obs": True, "time_limit": 11},
    {"num_passengers": 0, "time_limit": 5},
])
@pytest.mark.parametrize("b", [1, 513, 20_000])
def test_lockstep_vs_oracle(kwargs, b):
    """Ragged sizes (1, one-over-a-tile, many tiles), all maps, multi-passenger, mass truncation."""
    from gym_po.envs import TaxiVecEnv
    orc = oracle.TaxiOracle(b, draws=oracle.GeneratorDraws(seed=11), **kwargs)
    env = TaxiVecEnv(b, device=DEV, rng_mode="replay", **kwargs)
    rng = np.random.default_rng(2)
    # step before any reset is legal in the reference (state zeros)
    a = rng.integers(5, size=b)
    o = orc.step(a)
    env.set_replay(**orc.draws)
    _cmp(env.step(torch.as_tensor(a, dtype=torch.int8, device=DEV))[:4], o[:4], -1)
    o_obs, _ = orc.reset()
    env.set_replay(**orc.draws)
    g_obs, _ = env.reset()
    np.testing.assert_array_equal(g_obs.cpu().numpy(), o_obs)
    steps = 120 if b > 1000 else 450
    for t in range(steps):
        a = rng.integers(5, size=b)
        o = orc.step(a)
        env.set_replay(**orc.draws)
        _cmp(env.step(torch.as_tensor(a, dtype=torch.int8, device=DEV))[:4], o[:4], t)
    st = env.get_state()
    np.testing.assert_array_equal(st["s"].cpu().numpy(), orc.s)
    np.testing.assert_array_equal(st["elapsed"].cpu().numpy(), orc.elapsed)
    np.testing.assert_array_equal(st["ndrop"].cpu().numpy(), orc.ndrop)


@pytest.mark.parametrize("extended", [False, True])
def test_exhaustive_transition_table(extended):
    """Every reachable (state, action): teacher-force the oracle state, one step, compare everything.
    (States with the taxi standing ON a wall cell of the 8x8 map are unreachable — resets land on
    non-wall cells and moves never enter walls — and are excluded.)"""
    from gym_po.envs import TaxiVecEnv
    kw = {"map": oracle.EXTENDED_TAXI_MAP} if extended else {}
    probe = oracle.TaxiOracle(1, **kw)
    all_s = np.arange(probe.ns)
    r, c, _, _ = probe.decode(all_s)
    all_s = all_s[probe.tgrid[r, c] != "|"]
    ns = all_s.size
    states = np.repeat(all_s, 5)
    actions = np.tile(np.arange(5), ns)
    b = states.size
    for hansen in (False, True):
        for elapsed0, ndrop0, npass in ((0, 0, 1), (199, 0, 1), (200, 1, 3), (5, 2, 3)):
            orc = oracle.TaxiOracle(b, num_passengers=npass, hansen_obs=hansen, draws=oracle.GeneratorDraws(seed=1), **kw)
            env = TaxiVecEnv(b, num_passengers=npass, hansen_obs=hansen, device=DEV, rng_mode="replay", **kw)
            st = dict(s=states.copy(), elapsed=np.full(b, elapsed0), ndrop=np.full(b, ndrop0))
            orc.set_state(**st)
            env.set_state(**st)
            o = orc.step(actions)
            env.set_replay(**orc.draws)
            _cmp(env.step(torch.as_tensor(actions, dtype=torch.int8, device=DEV))[:4], o[:4], 0)
            np.testing.assert_array_equal(env.s.cpu().numpy(), orc.s)
            np.testing.assert_array_equal(env.elapsed.cpu().numpy(), orc.elapsed)
            np.testing.assert_array_equal(env.n_dropoffs_completed.cpu().numpy(), orc.ndrop)


def test_full_size_properties_philox():
    """BASELINE config 2 size (2^22 envs), Philox mode: size-independent invariants."""
    from gym_po.envs import TaxiVecEnv
    b = 1 << 22
    env = TaxiVecEnv(b, device=DEV, seed=7)
    obs, _ = env.reset()
    valid = torch.zeros(env.ns, dtype=torch.bool, device=DEV)
    valid[torch.as_tensor(env.valid_states, device=DEV).long()] = True
    assert bool(valid[obs.long()].all()), "reset must land on valid states"
    assert int(env.elapsed.max()) == 0
    gen = torch.Generator(device=DEV).manual_seed(0)
    total_term = 0
    for t in range(205):
        a = torch.randint(0, 5, (env.capacity,), dtype=torch.int8, device=DEV, generator=gen)
        prev_elapsed = env.elapsed.clone()
        obs, rew, term, trunc, _ = env.step(a)
        done = term | trunc
        # elapsed increments or resets; truncation exactly when elapsed would exceed the limit
        assert bool((env.elapsed[~done] == prev_elapsed[~done] + 1).all())
        assert bool((env.elapsed[done] == 0).all())
        assert bool((trunc == (prev_elapsed + 1 > env.time_limit)).all())
        assert bool(((rew == 1.0) | (rew == -0.5) | (rew == -0.05)).all())
        assert bool((rew[term] == 1.0).all())            # single passenger: terminated <=> delivered
        assert bool(valid[obs[done].long()].all())       # autoreset lands on valid states
        assert int(obs.min()) >= 0 and int(obs.max()) < env.ns
        total_term += int(term.sum())
    assert total_term > 0
    # step 201 after a synchronised reset truncates every env that did not terminate earlier
    assert int(env.elapsed.max()) <= 205


def test_philox_reset_law_and_gpu_count_independence():
    """(i) reset states follow the reference's argmax-of-multinomial law (chi-square);
    (ii) sharding the batch over 'ranks' with env_offset gives identical trajectories."""
    from gym_po.envs import TaxiVecEnv
    b = 1 << 21
    env = TaxiVecEnv(b, device=DEV, seed=123)
    obs, _ = env.reset()
    counts = torch.bincount(obs.long(), minlength=env.ns).cpu().numpy()[env.valid_states]
    expected = env.reset_law * b
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    dof = len(expected) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), f"chi2={chi2:.1f} dof={dof}"
    # uniform would be rejected decisively: the law is far from flat
    flat = np.full(len(expected), b / len(expected))
    assert float(((counts - flat) ** 2 / flat).sum()) > 50 * dof

    half = b // 2
    lo = TaxiVecEnv(half, device=DEV, seed=123, env_offset=0)
    hi = TaxiVecEnv(half, device=DEV, seed=123, env_offset=half)
    whole = TaxiVecEnv(b, device=DEV, seed=123)
    for e in (lo, hi, whole):
        e.reset(seed=123)
    gen = torch.Generator(device=DEV).manual_seed(5)
    for t in range(230):
        a = torch.randint(0, 5, (b,), dtype=torch.int8, device=DEV, generator=gen)
        w = whole.step(a)
        l = lo.step(a[:half].contiguous())
        h = hi.step(a[half:].contiguous())
        for k in range(4):
            assert torch.equal(w[k][:half], l[k]) and torch.equal(w[k][half:], h[k]), f"output {k} at step {t}"


def test_boundary_dlpack_and_errors():
    from gym_po.envs import TaxiVecEnv
    env = TaxiVecEnv(1024, device=DEV, seed=0)
    env.reset()
    a = torch.randint(0, 5, (env.capacity,), dtype=torch.int8, device=DEV)
    ref = TaxiVecEnv(1024, device=DEV, seed=0)
    ref.reset(seed=0)
    env.reset(seed=0)
    r1 = [x.clone() for x in ref.step(a)[:4]]
    r2 = env.step_dlpack(a)[:4]
    for x, y in zip(r1, r2):
        assert torch.equal(x, y)
    with pytest.raises(TypeError):      # wrong dtype through DLPack is rejected by the C library
        env.step_dlpack(a.to(torch.int32))
    with pytest.raises(TypeError):      # too few rows
        env.step_dlpack(a[:512])
    with pytest.raises(ValueError):
        env.step(torch.zeros(7, dtype=torch.int8, device=DEV))
    # int64 / numpy actions are accepted (converted / uploaded)
    env.step(torch.zeros(1024, dtype=torch.int64, device=DEV))
    env.step(np.zeros(1024, dtype=np.int64))
    with pytest.raises(ValueError):
        TaxiVecEnv(8, map=("A" + " " * 58 + "B",) + (" " * 60,) * 58 + ("C" + " " * 57 + "DE",), device=DEV)  # ns = 108000


def test_host_path_matches_device_path():
    from gym_po.envs import TaxiVecEnv
    b = 20_000
    dev_env = TaxiVecEnv(b, device=DEV, seed=9)
    host_env = TaxiVecEnv(b, device=DEV, seed=9)
    dev_env.reset(seed=9)
    host_env.reset(seed=9)
    rng = np.random.default_rng(0)
    for t in range(210):
        a = rng.integers(5, size=b).astype(np.int8)
        d = dev_env.step(torch.as_tensor(a, device=DEV))
        h = host_env.step_host(a)
        for k in range(4):
            np.testing.assert_array_equal(d[k].cpu().numpy(), h[k], err_msg=f"output {k} step {t}")


def test_episode_statistics_match_oracle_rollout():
    """track_stats=1: on-device {episodes, sum_return, sum_length, sum_return^2, env_steps} equal the values
    accumulated on the host from the oracle's rewards / done flags (padding rows excluded)."""
    from gym_po.envs import TaxiVecEnv
    b = 3000                       # ragged: capacity 3072
    kw = dict(time_limit=25, num_passengers=2)
    orc = oracle.TaxiOracle(b, draws=oracle.GeneratorDraws(seed=21), **kw)
    env = TaxiVecEnv(b, device=DEV, rng_mode="replay", track_stats=True, **kw)
    orc.reset()
    env.set_replay(**orc.draws)
    env.reset()
    rng = np.random.default_rng(6)
    ret = np.zeros(b, dtype=np.float32)
    length = np.zeros(b, dtype=np.int64)
    tot = np.zeros(5)
    for t in range(150):
        a = rng.integers(5, size=b)
        o, r, term, trunc, _ = orc.step(a)
        env.set_replay(**orc.draws)
        env.step(torch.as_tensor(a, dtype=torch.int8, device=DEV))
        ret += r
        length += 1
        done = term | trunc
        tot += [done.sum(), ret[done].astype(np.float64).sum(), length[done].sum(), (ret[done].astype(np.float64) ** 2).sum(), b]
        ret[done] = 0
        length[done] = 0
    got = env.stats_tensor().cpu().numpy()
    assert got[0] == tot[0] and got[2] == tot[2] and got[4] == tot[4]
    np.testing.assert_allclose(got[1], tot[1], rtol=1e-5)
    np.testing.assert_allclose(got[3], tot[3], rtol=1e-5)
    from gym_po.sharding import allreduce_stats
    out = allreduce_stats(env.stats_tensor().clone())
    assert abs(out["mean_length"] - tot[2] / tot[0]) < 1e-9
    env.stats_reset()
    torch.cuda.synchronize()
    assert float(env.stats_tensor().abs().sum()) == 0.0


def test_large_custom_map_uses_arithmetic_kernel():
    """ns = 15*15*7*6 = 9450 > 8192: the (state, action) table does not fit, the decode/wall-bit kernel runs."""
    from gym_po.envs import TaxiVecEnv
    big = ("A" + " " * 13 + "B",) + (" " * 6 + "|" + " " * 8,) * 6 + ("C" + " " * 13 + "D",) + (" " * 15,) * 6 + ("E" + " " * 13 + "F",)
    assert all(len(r) == 15 for r in big) and len(big) == 15
    b = 5000
    for hansen in (False, True):
        kw = dict(map=big, time_limit=40, num_passengers=2, hansen_obs=hansen)
        orc = oracle.TaxiOracle(b, draws=oracle.GeneratorDraws(seed=13), **kw)
        assert orc.ns == 9450
        env = TaxiVecEnv(b, device=DEV, rng_mode="replay", **kw)
        o_obs, _ = orc.reset()
        env.set_replay(**orc.draws)
        g_obs, _ = env.reset()
        np.testing.assert_array_equal(g_obs.cpu().numpy(), o_obs)
        rng = np.random.default_rng(2)
        for t in range(130):
            a = rng.integers(5, size=b)
            o = orc.step(a)
            env.set_replay(**orc.draws)
            _cmp(env.step(torch.as_tensor(a, dtype=torch.int8, device=DEV))[:4], o[:4], t)
        np.testing.assert_array_equal(env.s.cpu().numpy(), orc.s)
    with pytest.warns(RuntimeWarning):
        env = TaxiVecEnv(1024, map=big, device=DEV, seed=0)     # Philox mode: uniform reset law fallback
    obs, _ = env.reset()
    for _ in range(60):
        obs, *_ = env.step(torch.randint(0, 5, (1024,), dtype=torch.int8, device=DEV))
    assert int(obs.min()) >= 0 and int(obs.max()) < 9450


@pytest.mark.parametrize("hansen", [False, True])
def test_render_draws_device_state_like_the_reference(hansen):
    """render(idx) copies the selected envs' states to the host and reproduces the reference
    frames (fixtures drawn by the real reference for the same states)."""
    from gym_po.envs import TaxiVecEnv
    import os
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "render_taxi_5x5_hansen.npz" if hansen else "render_taxi_5x5.npz"))
    env = TaxiVecEnv(64, hansen_obs=hansen, device=DEV, seed=0)
    env.reset(seed=0)
    for s, nm, frame in zip(z["states"], z["names"], z["frames"]):
        k = len(s)
        st = env.get_state()
        st["s"][:k] = torch.as_tensor(s, device=DEV).to(torch.int32)
        env.set_state(st["s"].cpu().numpy(), st["elapsed"].cpu().numpy(), st["ndrop"].cpu().numpy())
        env._last_actions = None if not str(nm) else torch.full((64,), env.ACTION_NAMES.index(str(nm)), dtype=torch.int8)
        env._terminated[:] = False
        env._truncated[:] = False
        np.testing.assert_array_equal(env.render(idx=np.arange(k)), frame)
    # a real step records env 0's action; a finished episode clears the caption
    env.step(torch.full((env.capacity,), 2, dtype=torch.int8, device=DEV))
    a = env.render()
    env._terminated[0] = True
    b = env.render()
    assert a.shape == b.shape == (176, 132, 3) and (a != b).any() and (b[:, -20:] == 0).all()


@pytest.mark.parametrize("io", ["tma", "threads"])
@pytest.mark.parametrize("kw", [dict(), dict(hansen_obs=True, num_passengers=3, time_limit=17),
                                dict(map="ext", time_limit=9)])
def test_fused_multi_step_launch_equals_single_steps(kw, io):
    """gpt_step_many on Taxi runs T steps in ONE launch (state in registers); outputs of every step and the final
    state must be bit-identical to T single-step launches (Philox counters = (global env id, step))."""
    from gym_po.envs import EXTENDED_TAXI_MAP, TaxiVecEnv
    kw = dict(kw)
    if kw.get("map") == "ext":
        kw["map"] = EXTENDED_TAXI_MAP
    b, T = 3000, 37
    a = TaxiVecEnv(b, device=DEV, seed=9, **kw)
    c = TaxiVecEnv(b, device=DEV, seed=9, **kw)
    a.set_fused_steps(io)   # the fused kernel's I/O path: TMA bulk copies (default) or per-thread loads / stores
    a.reset(seed=9); c.reset(seed=9)
    gen = torch.Generator(device=DEV).manual_seed(4)
    for rep in range(3):
        acts = torch.randint(0, 5, (T, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
        if rep == 1:
            acts[:] = 4                      # deliveries / illegal pickups on every step
        out = {n: torch.zeros((T,) + tuple(a._arrays[n].shape), dtype=a._arrays[n].dtype, device=DEV)
               for n in ("obs", "reward", "terminated", "truncated")}
        l0 = a.launch_count
        a.step_many(acts, out)
        assert a.launch_count == l0 + 1      # one fused launch
        for t in range(T):
            o = c.step(acts[t])
            for n, x in zip(("obs", "reward", "terminated", "truncated"), o[:4]):
                assert torch.equal(out[n][t][:b].view(x.dtype), x), (n, rep, t)
        sa, sc = a.get_state(), c.get_state()
        for k in sa:
            assert torch.equal(sa[k], sc[k]), k
        assert a.rng_counter == c.rng_counter
    # in-place outputs (no rollout storage): the arrays hold the LAST step's results
    acts = torch.randint(0, 5, (5, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
    a.step_many(acts)
    for t in range(5):
        o = c.step(acts[t])
    for n, x in zip(("obs", "reward"), o[:2]):
        assert torch.equal(a._arrays[n][:b], x)


@pytest.mark.parametrize("b,kw", [(1, dict()), (1500, dict(time_limit=11)), (2500, dict(num_passengers=2, time_limit=13, hansen_obs=True))])
def test_fused_tma_launch_partial_cta(b, kw):
    """The default fused kernel moves its I/O with TMA bulk copies per CTA of 1024 envs: capacities with an ODD number of
    512-env tiles (half-filled last CTA), more steps than the action ring holds, rollout slots with a padded stride and
    in-place outputs — all equal to single-step launches."""
    from gym_po.envs import TaxiVecEnv
    T = 29
    a = TaxiVecEnv(b, device=DEV, seed=21, **kw)
    c = TaxiVecEnv(b, device=DEV, seed=21, **kw)
    a.set_fused_steps("tma")        # (the default for this family)
    assert (a.capacity // 512) % 2 == 1
    a.reset(seed=21); c.reset(seed=21)
    gen = torch.Generator(device=DEV).manual_seed(6)
    names = ("obs", "reward", "terminated", "truncated")
    for stride_pad in (0, 16 * 7):
        acts = torch.randint(0, 5, (T, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
        rows = a.capacity + stride_pad
        store = {n: torch.full((T * rows,), 77, dtype=a._arrays[n].dtype, device=DEV) for n in names}
        l0 = a.launch_count
        a.step_many(acts, {n: v.view(T, rows) for n, v in store.items()})
        assert a.launch_count == l0 + 1
        for t in range(T):
            o = c.step(acts[t])
            for n, x in zip(names, o[:4]):
                got = store[n][t * rows: t * rows + b]
                assert torch.equal(got.view(x.dtype), x), (n, stride_pad, t)
            if stride_pad:   # the padding between rollout slots is untouched
                for n in names:
                    assert bool((store[n][t * rows + a.capacity:(t + 1) * rows] == 77).all()), (n, t)
        sa, sc = a.get_state(), c.get_state()
        for k in sa:
            assert torch.equal(sa[k], sc[k]), k
    acts = torch.randint(0, 5, (11, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
    a.step_many(acts)                       # in place: the bound arrays hold the last step's results
    for t in range(11):
        o = c.step(acts[t])
    for n, x in zip(names, o[:4]):
        assert torch.equal(a._arrays[n][:b].view(x.dtype), x), n


@pytest.mark.parametrize("b", [1, 513, 4096])
def test_fused_launch_ragged_sizes_and_stats(b):
    """Fused launches at ragged batch sizes, with in-kernel statistics: equal to single steps, padding rows excluded."""
    from gym_po.envs import TaxiVecEnv
    T = 23
    a = TaxiVecEnv(b, time_limit=7, device=DEV, seed=3, track_stats=True)
    c = TaxiVecEnv(b, time_limit=7, device=DEV, seed=3, track_stats=True)
    a.reset(seed=3); c.reset(seed=3)
    gen = torch.Generator(device=DEV).manual_seed(4)
    acts = torch.randint(0, 5, (T, a.capacity), dtype=torch.int8, device=DEV, generator=gen)
    out = {n: torch.zeros((T,) + tuple(a._arrays[n].shape), dtype=a._arrays[n].dtype, device=DEV)
           for n in ("obs", "reward", "terminated", "truncated")}
    a.step_many(acts, out)
    for t in range(T):
        o = c.step(acts[t])
        for n, x in zip(("obs", "reward", "terminated", "truncated"), o[:4]):
            assert torch.equal(out[n][t][:b].view(x.dtype), x), (n, t)
    sa, sc = a.stats_tensor().cpu().numpy(), c.stats_tensor().cpu().numpy()
    assert sa[0] == sc[0] > 0 and sa[2] == sc[2] and sa[4] == sc[4] == b * T
    np.testing.assert_allclose(sa[[1, 3]], sc[[1, 3]], rtol=1e-6)
    # the switch: one launch per step on request
    a.set_fused_steps(False)
    l0 = a.launch_count
    a.step_many(acts[:5])
    assert a.launch_count == l0 + 5


@pytest.mark.parametrize("family", ["taxi", "rooms", "msrooms", "crooms", "tag", "car"])
def test_cuda_graph_capture_and_replay(family):
    """Graph mode: step() calls captured into a CUDA graph draw fresh random numbers on every replay (the Philox step
    counter lives in device memory) — the replayed trajectory equals an eagerly stepped twin env, bit for bit."""
    from gym_po.envs import CarVecEnv, CRoomsEnv, MultistoryFourRoomsEnv, RoomsEnv, TagVecEnv, TaxiVecEnv
    b, K = 3000, 4
    mk, n_act = {
        "taxi": (lambda: TaxiVecEnv(b, time_limit=9, hansen_obs=True, device=DEV, seed=3), 5),
        "rooms": (lambda: RoomsEnv(b, "8", obs_type="grid", obs_n=5, goal_xy=None, time_limit=9, device=DEV, seed=3), 8),
        "msrooms": (lambda: MultistoryFourRoomsEnv(b, grid_z=2, obs_type="vector_mdp_goal", goal_xyz=None, time_limit=9, device=DEV, seed=3), 4),
        "crooms": (lambda: CRoomsEnv(b, "4", obs_type="vector_mdp", time_limit=9, device=DEV, seed=3, precision="float32"), 0),
        "tag": (lambda: TagVecEnv(b, time_limit=9, device=DEV, seed=3), 0),
        "car": (lambda: CarVecEnv(b, time_limit=9, device=DEV, seed=3), -1),
    }[family]
    env, twin = mk(), mk()
    env.reset(seed=3); twin.reset(seed=3)
    env.set_graph_mode(True)
    if n_act > 0:
        static_a = torch.zeros((K, env.capacity), dtype=torch.int8, device=DEV)
        draw = lambda gen: torch.randint(0, n_act, static_a.shape, dtype=torch.int8, device=DEV, generator=gen)
    else:
        static_a = torch.zeros((K, env.capacity) + ((2,) if n_act == 0 else ()), dtype=torch.float32, device=DEV)
        draw = lambda gen: torch.rand(static_a.shape, device=DEV, generator=gen) * 2 - 1
    first = env.step(static_a[0])                                      # one eager step in graph mode
    rec = [[torch.zeros_like(x) for x in first[:4]] for _ in range(K)]
    twin.step(static_a[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(K):
            out = env.step(static_a[i])
            for dst, src in zip(rec[i], out[:4]):
                dst.copy_(src)
    gen = torch.Generator(device=DEV).manual_seed(5)
    for rep in range(6):
        static_a.copy_(draw(gen))
        g.replay()
        torch.cuda.synchronize()
        for i in range(K):
            o = twin.step(static_a[i])
            for name, x, y in zip(("obs", "reward", "terminated", "truncated"), rec[i], o[:4]):
                assert torch.equal(x, y), (name, rep, i)
    sa, sb = env.get_state(), twin.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert env.rng_counter == twin.rng_counter          # read back from the device
    env.set_graph_mode(True)                             # enabling twice must not rewind the device counter
    assert env.rng_counter == twin.rng_counter
    env.set_graph_mode(False)                            # and back to launch-parameter counters
    a = draw(gen)[0]
    for x, y in zip(env.step(a)[:4], twin.step(a)[:4]):
        assert torch.equal(x, y)
    with pytest.raises(ValueError):
        TaxiVecEnv(64, device=DEV, rng_mode="replay").set_graph_mode(True)


@pytest.mark.parametrize("family", ["taxi", "rooms", "msrooms"])
def test_cuda_graph_captures_fused_launches(family):
    """Graph mode keeps the fused multi-step launches (and PDL): a captured step_many() is ONE kernel whose Philox step
    counter advances by T on the device at every replay; trajectories equal an eagerly stepped twin, bit for bit."""
    from gym_po.envs import MultistoryFourRoomsEnv, RoomsEnv, TaxiVecEnv
    b, T = 5000, 6
    mk, n_act = {
        "taxi": (lambda: TaxiVecEnv(b, time_limit=11, num_passengers=2, device=DEV, seed=4), 5),
        "rooms": (lambda: RoomsEnv(b, "4", obs_type="hansen8", time_limit=11, device=DEV, seed=4), 8),
        "msrooms": (lambda: MultistoryFourRoomsEnv(b, grid_z=2, time_limit=11, device=DEV, seed=4), 4),
    }[family]
    env, twin = mk(), mk()
    env.reset(seed=4); twin.reset(seed=4)
    env.set_graph_mode(True)
    names = ("obs", "reward", "terminated", "truncated")
    static_a = torch.zeros((T, env.capacity), dtype=torch.int8, device=DEV)
    out = {n: torch.zeros((T,) + tuple(env._arrays[n].shape), dtype=env._arrays[n].dtype, device=DEV) for n in names}
    ref = {n: torch.zeros_like(out[n]) for n in names}
    l0 = env.launch_count
    env.step_many(static_a, out)                      # eager, in graph mode: one fused launch
    assert env.launch_count - l0 == 1
    twin.step_many(static_a, ref)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        env.step_many(static_a, out)
    gen = torch.Generator(device=DEV).manual_seed(6)
    for rep in range(5):
        static_a.copy_(torch.randint(0, n_act, static_a.shape, dtype=torch.int8, device=DEV, generator=gen))
        g.replay()
        twin.step_many(static_a, ref)
        torch.cuda.synchronize()
        for n in names:
            assert torch.equal(out[n], ref[n]), (n, rep)
    sa, sb = env.get_state(), twin.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert env.rng_counter == twin.rng_counter == 1 + T * 6   # reset + one eager launch + five replays (capturing does not execute)

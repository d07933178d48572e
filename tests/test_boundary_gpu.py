"""GPU: boundary hygiene of the C ABI (ADVICE r01 / VERDICT r01 next-round item 9): stream ordering of the host path,
reset on uninitialised state memory, the opt-in action range check, wrappers refusing to be bypassed."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _busy(ms=30):
    """Queue roughly `ms` of work on the current stream (so that anything issued next on ANOTHER stream without an
    ordering edge would overtake the kernels queued before it)."""
    torch.cuda._sleep(int(ms * 1.9e6))


@pytest.mark.parametrize("family", ["taxi", "rooms"])
def test_step_host_is_ordered_after_the_callers_stream(family):
    """reset() / step() on torch's stream followed by step_host() (internal non-blocking streams) and back: same
    results as the all-device path, even with a long kernel queued in front of the reset."""
    from gym_po.envs import RoomsEnv, TaxiVecEnv
    b = 40_000
    mk = (lambda: TaxiVecEnv(b, device=DEV, seed=5)) if family == "taxi" else (lambda: RoomsEnv(b, "4", obs_type="hansen8", device=DEV, seed=5))
    n_act = 5 if family == "taxi" else 8
    ref, env = mk(), mk()
    rng = np.random.default_rng(0)
    acts = rng.integers(n_act, size=(6, b)).astype(np.int8)
    ref.reset(seed=9)
    want = []
    for t in range(6):
        o = ref.step(torch.as_tensor(acts[t], device=DEV))
        want.append([x.cpu().numpy().copy() for x in o[:4]])
    _busy()
    env.reset(seed=9)                      # queued behind the sleep kernel
    got = [env.step_host(acts[0])]         # must wait for that reset
    got[0] = [np.array(x) for x in got[0][:4]]
    _busy()
    o = [x.clone() for x in env.step(torch.as_tensor(acts[1], device=DEV))[:4]]   # device step behind a sleep ...
    h = env.step_host(acts[2])                           # ... and the host step must see its state
    got.append([x.cpu().numpy() for x in o])
    got.append([np.array(x) for x in h[:4]])
    for t in (3, 4, 5):
        o = env.step(torch.as_tensor(acts[t], device=DEV))   # the caller's stream waits for the host path
        got.append([x.cpu().numpy().copy() for x in o[:4]])
    for t in range(6):
        for name, x, y in zip(("obs", "reward", "terminated", "truncated"), got[t], want[t]):
            np.testing.assert_array_equal(x, y, err_msg=f"{name} at step {t}")


@pytest.mark.parametrize("family", ["taxi", "rooms", "msrooms"])
def test_reset_on_garbage_state_memory(family):
    """A C-ABI caller may bind freshly allocated (uninitialised) memory: gpt_reset clears the state arrays before the
    kernel indexes its shared-memory tables with them."""
    from gym_po.envs import MultistoryFourRoomsEnv, RoomsEnv, TaxiVecEnv
    b = 10_000
    env = {"taxi": lambda: TaxiVecEnv(b, device=DEV, seed=1),
           "rooms": lambda: RoomsEnv(b, "4", obs_type="hansen8", goal_xy=None, device=DEV, seed=1),
           "msrooms": lambda: MultistoryFourRoomsEnv(b, grid_z=2, goal_xyz=None, device=DEV, seed=1)}[family]()
    for name, (i, role, dt, cols) in env._descs.items():
        if name in env._arrays and role == 0:   # STATE
            env._arrays[name].view(torch.uint8).fill_(0x7B)
    out = env.reset()
    obs = out[0] if isinstance(out, tuple) else out
    torch.cuda.synchronize()                    # an out-of-bounds shared-memory read would fault here
    assert int(obs.min()) >= 0
    a = torch.zeros(env.capacity, dtype=torch.int8, device=DEV)
    env.step(a)
    torch.cuda.synchronize()


def test_set_state_rejects_out_of_range_values():
    from gym_po.envs import RoomsEnv, TaxiVecEnv
    env = TaxiVecEnv(8, device=DEV, seed=0)
    with pytest.raises(ValueError):
        env.set_state(np.full(8, env.ns), np.zeros(8), np.zeros(8))
    r = RoomsEnv(8, "4", device=DEV, seed=0)
    with pytest.raises(ValueError):
        r.set_state(np.full((8, 2), 99), None, np.zeros(8))


def test_debug_action_range_check(monkeypatch):
    """GPT_DEBUG_ACTIONS=1: IndexError like the reference's numpy indexing (extended_taxi.py:248); off by default
    (the kernels mask the byte)."""
    import gym_po._device_env as de
    from gym_po.envs import TaxiVecEnv
    env = TaxiVecEnv(1000, device=DEV, seed=0)
    env.reset()
    bad = np.zeros(1000, dtype=np.int64)
    bad[17] = 5
    env.step(bad)                                   # default: no check, no exception
    a8 = torch.zeros(env.capacity, dtype=torch.int8, device=DEV)
    a8[3], a8[900] = 7, -1
    assert env.check_actions(a8) == 2               # gpt_check_actions counts bytes outside [0, 5)
    a8[1001] = 99                                   # padding rows are not checked
    assert env.check_actions(a8) == 2
    monkeypatch.setattr(de, "DEBUG_ACTIONS", True)
    with pytest.raises(IndexError):
        env.step(bad)
    with pytest.raises(IndexError):
        env.step(a8)
    with pytest.raises(IndexError):
        env.step_host(np.full(1000, 300, dtype=np.int64))   # would wrap to 44 in int8
    env.step(np.zeros(1000, dtype=np.int64))


def test_wrappers_refuse_to_be_bypassed():
    from gym_po.envs import TaxiVecEnv
    from gym_po.wrappers import NormalizeReward, RecordEpisodeStatistics
    base = TaxiVecEnv(600, device=DEV, seed=0)
    env = NormalizeReward(RecordEpisodeStatistics(base))
    env.reset()
    a = torch.zeros((2, base.capacity), dtype=torch.int8, device=DEV)
    with pytest.raises(NotImplementedError):
        env.step_many(a)
    with pytest.raises(NotImplementedError):
        env.step_host(np.zeros(600, dtype=np.int8))
    rec = RecordEpisodeStatistics(TaxiVecEnv(600, device=DEV, seed=0, time_limit=3))
    rec.reset()
    for _ in range(4):
        out = rec.step_dlpack(a[0])                 # goes through the wrapper kernel
    assert "episode" in out[4] and int(out[4]["_episode"].sum()) == 600
    assert float(rec.stats_tensor()[0]) == 600.0


def test_table_read_names_and_sizes():
    from gym_po.envs import RoomsEnv, TaxiVecEnv
    t = TaxiVecEnv(8, device=DEV, seed=0)
    al = t.read_table("reset_alias", np.uint32).reshape(-1, 2)
    assert al.shape[0] == len(t.valid_states)
    vals = np.concatenate((al[:, 1] & 0xFFFF, al[:, 1] >> 16))
    assert set(vals.tolist()) <= set(t.valid_states.tolist())
    r = RoomsEnv(8, "4", device=DEV, seed=0)
    cells = r.read_table("spawn_cells", np.uint16)
    grid = oracle.load_layout("4")
    np.testing.assert_array_equal(cells, np.flatnonzero(grid >= 0))
    with pytest.raises(ValueError):
        r.read_table("no_such_table", np.uint8)

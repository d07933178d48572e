"""CPU property tests (hypothesis) for the host logic and the oracle — size-independent invariants of the domain."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle


@settings(max_examples=60, deadline=None)
@given(total=st.integers(1, 5_000_000), world=st.integers(1, 16))
def test_shards_partition_the_batch(total, world):
    """Contiguous, tile-aligned, disjoint shards that cover [0, total) whatever the GPU count."""
    from gym_po.sharding import shard_envs
    end = 0
    for rank in range(world):
        n, off = shard_envs(total, rank, world)
        assert off % 512 == 0 and n >= 0
        if n:
            assert off == end
            end = off + n
    assert end == total


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), passengers=st.integers(1, 3), ext=st.booleans(), hansen=st.booleans())
def test_taxi_oracle_invariants(seed, passengers, ext, hansen):
    """Any action stream keeps the state valid, rewards in the three-value set, counters consistent."""
    b = 48
    env = oracle.TaxiOracle(b, num_passengers=passengers, time_limit=25, hansen_obs=hansen,
                            map=oracle.EXTENDED_TAXI_MAP if ext else oracle.TAXI_MAP, draws=oracle.GeneratorDraws(seed=seed))
    env.reset()
    rng = np.random.default_rng(seed)
    wall = np.array([[ch == "|" for ch in row[::(1 if ext else 2)]] for row in (oracle.EXTENDED_TAXI_MAP if ext else oracle.TAXI_MAP)])
    for _ in range(60):
        prev = env.elapsed.copy()
        obs, rew, term, trunc, _ = env.step(rng.integers(5, size=b))
        r, c, p, d = env.decode(env.s)
        assert ((0 <= r) & (r < env.rows) & (0 <= c) & (c < env.cols)).all() and not wall[r, c].any()
        assert (p <= env.nlocs).all() and (d < env.nlocs).all() and (p != d).all()
        assert np.isin(rew, np.float32([1.0, -0.5, -0.05])).all()
        done = term | trunc
        assert (env.elapsed[done] == 0).all() and (env.elapsed[~done] == prev[~done] + 1).all()
        assert (trunc == (prev + 1 > 25)).all()
        assert (rew[term] == 1.0).all()                      # the episode ends on a delivery
        assert (0 <= obs).all() and (obs < (320 if hansen else env.ns)).all()


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), layout=st.sampled_from(list(oracle.LAYOUT_NAMES)), rgoal=st.booleans(),
       obs_type=st.sampled_from(["hansen8", "vector_goal_hansen", "grid", "mdp"]))
def test_rooms_oracle_invariants(seed, layout, rgoal, obs_type):
    """Agents never stand in a wall, the goal reward is paid exactly on goal hits, the window obs marks agent and goal."""
    b = 32
    env = oracle.RoomsOracle(b, layout, obs_type=obs_type, obs_n=5, goal_xy=None if rgoal else (0, 0), time_limit=20,
                             wall_reward=-0.5, step_reward=-0.1, draws=oracle.GeneratorDraws(seed=seed))
    env.reset()
    grid = oracle.load_layout(layout)
    rng = np.random.default_rng(seed)
    for _ in range(40):
        obs, rew, term, trunc, _ = env.step(rng.integers(8, size=b))
        assert (grid[env.agent[:, 0], env.agent[:, 1]] >= 0).all()
        assert (rew[term] == 1.0).all() and np.isin(rew[~term], np.float32([-0.5, -0.1])).all()
        assert (env.elapsed[term | trunc] == 0).all()
        if obs_type == "grid":
            assert obs.shape == (b, 5, 5) and np.isin(obs, [0, 1, 2]).all() and (obs[:, 2, 2] >= 1).all()


@settings(max_examples=30, deadline=None)
@given(n=st.integers(1, 40), h=st.integers(1, 6), w=st.integers(1, 6))
def test_render_tiling_shape(n, h, w):
    from gym_po.envs.taxi_render import tile
    img = np.arange(n * h * w * 3, dtype=np.uint8).reshape(n, h, w, 3)
    sheet = tile(img)
    p = int(np.ceil(np.sqrt(n)))
    q = int(np.ceil(n / p))
    assert sheet.shape == (p * h, q * w, 3)
    assert (sheet[:h, :w] == img[0]).all()
    if n > 1:
        assert (sheet[:h, w:2 * w] == img[1]).all() if q > 1 else (sheet[h:2 * h, :w] == img[1]).all()

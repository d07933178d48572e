"""Wrapper layer: device-side RecordEpisodeStatistics / NormalizeReward vs the numpy
restatement of gymnasium's algorithm (oracle/wrappers.py; parity unpinned — gymnasium is not installed)."""
import numpy as np
import pytest
import torch

import oracle

DEV = "cuda:0"


def test_running_mean_std_merge_equals_whole_sample_moments():
    """CPU: the batch-merge recurrence reproduces the mean / variance of all samples seen (up to the 1e-4 prior)."""
    rng = np.random.default_rng(0)
    rms = oracle.RunningMeanStd()
    xs = []
    for _ in range(50):
        x = rng.normal(3.0, 2.0, size=977)
        xs.append(x)
        rms.update(x)
    allx = np.concatenate(xs)
    assert abs(rms.mean - allx.mean()) < 1e-6 and abs(rms.var - allx.var()) < 1e-5
    assert abs(rms.count - (len(allx) + 1e-4)) < 1e-9


def test_record_episode_statistics_oracle_bookkeeping():
    rec = oracle.RecordEpisodeStatisticsOracle(3)
    i1 = rec.step(np.array([1, 2, 3], np.float32), np.array([0, 0, 1], bool), np.array([0, 0, 0], bool))
    i2 = rec.step(np.array([1, 2, 3], np.float32), np.array([0, 0, 0], bool), np.array([1, 0, 0], bool))
    assert i1["r"].tolist() == [0, 0, 3] and i1["l"].tolist() == [0, 0, 1]
    assert i2["r"].tolist() == [2, 0, 0] and i2["l"].tolist() == [2, 0, 0]
    assert rec.totals.tolist() == [2, 5, 3, 13, 6] and rec.episode_returns.tolist() == [0, 4, 3]


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["taxi", "rooms", "tag"])
def test_wrappers_match_the_gymnasium_algorithm(family):
    from gym_po.envs import RoomsEnv, TagVecEnv, TaxiVecEnv
    from gym_po.wrappers import NormalizeReward, RecordEpisodeStatistics
    b = 5003
    if family == "taxi":
        base, n_act = TaxiVecEnv(b, time_limit=13, device=DEV, seed=2), 5
    elif family == "rooms":
        base, n_act = RoomsEnv(b, "2", obs_type="hansen8", goal_xy=None, time_limit=11, step_reward=-0.25, wall_reward=-1.0,
                               goal_reward=4.0, device=DEV, seed=2), 8
    else:
        base, n_act = TagVecEnv(b, time_limit=9, device=DEV, seed=2, precision="float32"), 0
    env = NormalizeReward(RecordEpisodeStatistics(base), gamma=0.95, epsilon=1e-8)
    assert env.num_envs == b and env.single_action_space is base.single_action_space   # attribute forwarding
    env.reset(seed=2)
    rec = oracle.RecordEpisodeStatisticsOracle(b)
    nrm = oracle.NormalizeRewardOracle(b, gamma=0.95, epsilon=1e-8)
    gen = torch.Generator(device=DEV).manual_seed(1)
    for t in range(150):
        if n_act:
            a = torch.randint(0, n_act, (base.capacity,), dtype=torch.int8, device=DEV, generator=gen)
        else:
            a = torch.rand((base.capacity, 2), device=DEV, generator=gen) * 2 - 1
        obs, nr, term, trunc, info = env.step(a)
        raw = base._reward.cpu().numpy()
        tm, tr = term.cpu().numpy(), trunc.cpu().numpy()
        oi = rec.step(raw, tm, tr)
        onr = nrm.step(raw, tm)
        np.testing.assert_array_equal(info["episode"]["r"].cpu().numpy(), oi["r"], err_msg=f"t={t}")   # same float32 sums
        np.testing.assert_array_equal(info["episode"]["l"].cpu().numpy(), oi["l"])
        np.testing.assert_array_equal(info["_episode"].cpu().numpy(), oi["_episode"])
        np.testing.assert_allclose(nr.cpu().numpy(), onr, rtol=1e-5, atol=1e-7, err_msg=f"t={t}")       # stated tolerance
    np.testing.assert_array_equal(env.env.episode_returns.cpu().numpy(), rec.episode_returns)
    np.testing.assert_array_equal(env.env.episode_lengths.cpu().numpy(), rec.episode_lengths)
    got = env.env.stats_tensor().cpu().numpy()
    assert rec.totals[0] > 1000
    assert got[0] == rec.totals[0] and got[2] == rec.totals[2] and got[4] == rec.totals[4]
    np.testing.assert_allclose(got[[1, 3]], rec.totals[[1, 3]], rtol=1e-5)
    rms = env.return_rms()
    np.testing.assert_allclose([rms["count"], rms["mean"], rms["var"]],
                               [nrm.return_rms.count, nrm.return_rms.mean, nrm.return_rms.var], rtol=1e-5)
    s = env.env.stats()
    assert abs(s["mean_return"] - rec.totals[1] / rec.totals[0]) < 1e-4
    assert env.launch_count == 300 and env.env.launch_count == 150


@pytest.mark.gpu
def test_fused_stats_agree_with_the_wrapper():
    """The in-kernel statistics of track_stats=True (Taxi) equal what the family-independent wrapper accumulates."""
    from gym_po.envs import TaxiVecEnv
    from gym_po.wrappers import RecordEpisodeStatistics
    b = 1 << 16
    base = TaxiVecEnv(b, time_limit=20, device=DEV, seed=5, track_stats=True)
    env = RecordEpisodeStatistics(base)
    env.reset(seed=5)
    gen = torch.Generator(device=DEV).manual_seed(1)
    for t in range(100):
        env.step(torch.randint(0, 5, (base.capacity,), dtype=torch.int8, device=DEV, generator=gen))
    a, w = base.stats_tensor().cpu().numpy(), env.stats_tensor().cpu().numpy()
    assert a[0] == w[0] > 0 and a[2] == w[2] and a[4] == w[4]
    np.testing.assert_allclose(a[[1, 3]], w[[1, 3]], rtol=1e-6)

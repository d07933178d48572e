"""TEST INFRASTRUCTURE — numpy restatement of the two gymnasium wrappers the reference's author stacks on the
vector envs (gym_po/tester.py:36-41: ``NormalizeReward(e, 0.95)``, ``RecordEpisodeStatistics(e, ...)``), SURVEY.md
§8f row 3.

PARITY UNPINNED: the wrappers live in gymnasium (third-party, un-vendored; setup.py:39 pins ``gymnasium>=0.26.0``;
not installed in the build image and not fetchable), so there is nothing to execute.  This restates the published
algorithm of gymnasium 0.27-0.29 (``gymnasium/wrappers/record_episode_statistics.py``, ``normalize.py``):

* RecordEpisodeStatistics: per-env running return / length; on ``terminated | truncated`` the finished episode's
  return and length are reported in ``info["episode"]["r"/"l"]`` (0 elsewhere, mask in ``info["_episode"]``) and
  the accumulators are cleared.
* NormalizeReward: ``returns = returns * gamma * (1 - terminated) + reward``; the running mean / variance of
  ``returns`` (RunningMeanStd, parallel-variance merge of one batch per step, count initialised to 1e-4) is updated
  and the reward is divided by ``sqrt(var + epsilon)``.
"""
from __future__ import annotations

import numpy as np


class RunningMeanStd:
    """gymnasium.wrappers.normalize.RunningMeanStd for a scalar statistic (shape ())."""

    def __init__(self, epsilon=1e-4):
        self.mean, self.var, self.count = 0.0, 1.0, epsilon

    def update(self, x):
        x = np.asarray(x, dtype=np.float64)
        batch_mean, batch_var, batch_count = x.mean(axis=0), x.var(axis=0), x.shape[0]
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / tot
        self.mean, self.var, self.count = float(new_mean), float(m2 / tot), float(tot)


class RecordEpisodeStatisticsOracle:
    def __init__(self, num_envs):
        self.episode_returns = np.zeros(num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(num_envs, dtype=np.int64)
        self.episode_count = 0
        self.totals = np.zeros(5)   # episodes, sum return, sum length, sum return^2, env-steps

    def step(self, reward, terminated, truncated):
        self.episode_returns += reward
        self.episode_lengths += 1
        done = np.logical_or(terminated, truncated)
        info = {"r": np.where(done, self.episode_returns, np.float32(0)), "l": np.where(done, self.episode_lengths, 0), "_episode": done}
        r = self.episode_returns[done].astype(np.float64)
        self.totals += [done.sum(), r.sum(), self.episode_lengths[done].sum(), (r * r).sum(), len(done)]
        self.episode_count += int(done.sum())
        self.episode_lengths[done] = 0
        self.episode_returns[done] = 0
        return info


class NormalizeRewardOracle:
    def __init__(self, num_envs, gamma=0.99, epsilon=1e-8):
        self.returns = np.zeros(num_envs)
        self.gamma, self.epsilon = gamma, epsilon
        self.return_rms = RunningMeanStd()

    def step(self, reward, terminated):
        self.returns = self.returns * self.gamma * (1 - np.asarray(terminated, dtype=np.float64)) + reward
        self.return_rms.update(self.returns)
        return reward / np.sqrt(self.return_rms.var + self.epsilon)

"""TEST INFRASTRUCTURE — numpy oracle for the ant-tag *pursuit rules* and the point-mass
Tag env built from them.

The reference ``AntTagEnv`` (gym_po/envs/ant_tag.py) is a single-env MuJoCo task; its
rigid-body physics (``do_simulation`` :139) is third-party and out of scope (parity
unpinned).  What *is* restated here, vectorized over B envs:

* ``_move_target`` :105-123 — target flees / side-steps / stays relative to the agent
* constants :69-73 — cage 4.5, visible radius 3.0 (strict), tag radius 1.5 (<=),
  minimum spawn distance 5.0, target step 0.5
* reset rejection :94-100, tag reward/termination :144-150, visibility gate :153
* time limit 500 through gymnasium's TimeLimit (envs/__init__.py:15-19):
  ``truncated = elapsed >= 500``

The agent body is replaced by the CRooms point-mass motion model
(rooms/crooms.py:175-178, :312-314): ``pos += a*power + N(0, action_std^2)`` clipped to
the arena's inner walls (+-5.0, assets/ant_tag_small.xml:72-83).  That combination has no
reference counterpart (DESIGN.md says so); the target rule is pinned against
``AntTagEnv._move_target`` called unbound on a stand-in ``self``.
"""
from __future__ import annotations

import numpy as np

from .draws import GeneratorDraws

CAGE = 4.5
VISIBLE_RADIUS = 3.0
TAG_RADIUS = 1.5
MIN_SPAWN_DISTANCE = 5.0
TARGET_STEP = 0.5
ARENA = 5.0


def tag_move_target(agent_xy, target_xy, choice, cage=CAGE, step=TARGET_STEP):
    """Vectorized ant_tag.py:105-123.  ``choice`` in {0 away, 1 (vy,-vx), 2 (-vy,vx), 3 stay}."""
    agent_xy = np.asarray(agent_xy, dtype=np.float64)
    target_xy = np.asarray(target_xy, dtype=np.float64)
    choice = np.asarray(choice)
    v = agent_xy - target_xy
    v = v / np.linalg.norm(v, axis=-1, keepdims=True)
    move = np.zeros_like(v)
    away, side1, side2 = choice == 0, choice == 1, choice == 2
    move[away] = -v[away]
    move[side1, 0], move[side1, 1] = v[side1, 1], -v[side1, 0]
    move[side2, 0], move[side2, 1] = -v[side2, 1], v[side2, 0]
    new = move * step
    new = new + target_xy
    outside = (np.abs(new) > cage).any(-1)
    new[outside] = target_xy[outside]
    return new


class TagOracle:
    """Point-mass Tag: B agents chase B evading targets (DESIGN.md 'Tag')."""

    def __init__(self, num_envs, time_limit=500, action_std=0.2, action_power=1.0, draws=None):
        self.num_envs = int(num_envs)
        self.time_limit = time_limit
        self.action_std, self.action_power = action_std, action_power
        self.rng = draws if draws is not None else GeneratorDraws()
        self.draws = {}

    def _blank_draws(self):
        b = self.num_envs
        return {"noise": np.zeros((b, 2)), "choice": np.zeros(b, np.int8),
                "spawn_agent": np.zeros((b, 2)), "spawn_target": np.zeros((b, 2))}

    def _spawn(self, mask):
        """ant_tag.py:88-103: agent uniform in the cage, target redrawn while within 5.0"""
        idx = np.flatnonzero(mask)
        if idx.size == 0:
            return
        agent = self.rng.uniform(-CAGE, CAGE, (idx.size, 2))
        target = self.rng.uniform(-CAGE, CAGE, (idx.size, 2))
        while True:
            close = np.linalg.norm(agent - target, axis=-1) <= MIN_SPAWN_DISTANCE
            if not close.any():
                break
            target[close] = self.rng.uniform(-CAGE, CAGE, (int(close.sum()), 2))
        self.agent[idx], self.target[idx] = agent, target
        self.elapsed[idx] = 0
        self.draws["spawn_agent"][idx] = agent
        self.draws["spawn_target"][idx] = target

    @property
    def state(self):
        return {"agent": self.agent.copy(), "target": self.target.copy(), "elapsed": self.elapsed.copy()}

    def set_state(self, agent, target, elapsed):
        self.agent = np.array(agent, dtype=np.float64)
        self.target = np.array(target, dtype=np.float64)
        self.elapsed = np.array(elapsed, dtype=np.int64)

    def _obs(self):
        d = np.linalg.norm(self.agent - self.target, axis=-1)
        return np.where((d < VISIBLE_RADIUS)[:, None], self.target, 0.0)

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.rng.reseed(seed)
        b = self.num_envs
        self.draws = self._blank_draws()
        self.agent, self.target = np.zeros((b, 2)), np.zeros((b, 2))
        self.elapsed = np.zeros(b, dtype=np.int64)
        self._spawn(np.ones(b, dtype=bool))
        return self._obs(), {}

    def step(self, action):
        action = np.asarray(action, dtype=np.float64)
        b = self.num_envs
        self.draws = self._blank_draws()
        self.elapsed += 1
        noise = self.rng.normal(self.action_std, action.shape)
        self.draws["noise"][:] = noise
        push = (action + noise) * self.action_power
        self.agent = (self.agent + push).clip(-ARENA, ARENA)
        choice = self.rng.integers(4, size=b)
        self.draws["choice"][:] = choice
        self.target = tag_move_target(self.agent, self.target, choice)
        d = np.linalg.norm(self.agent - self.target, axis=-1)
        tagged = d <= TAG_RADIUS
        rew = np.where(tagged, 1.0, 0.0).astype(np.float32)
        truncated = self.elapsed >= self.time_limit
        self._spawn(tagged | truncated)
        return self._obs(), rew, tagged, truncated, {}

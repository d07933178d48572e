"""TEST INFRASTRUCTURE — numpy oracle for the vectorized car-flag ("heaven/hell with a priest") env.

Restates ``CarVecEnv`` / ``DiscreteActionCarVecEnv`` of the reference (gym_po/envs/car_flag.py: constants
:25-34, ctor :49-85, ``reset`` :87-95, ``_reset_mask`` :97-112, ``step`` :114-141, ``_obs`` :143-144,
discrete wrapper :286-303).  dtype behaviour is part of the semantics: the state ``s`` is float32 [B,3]
(position, velocity, priest indicator) and IS the observation; position/velocity arithmetic happens in the
promotion of float32 with the action dtype (float32 actions -> float32 math, float64 actions -> float64
math, rounded to float32 when stored); the priest window is compared in float64.
Note ``truncated = elapsed >= time_limit`` here (the other envs use ``>``).
"""
from __future__ import annotations

import numpy as np

from .draws import GeneratorDraws

MAX_POS = 1.1
MIN_POS = -MAX_POS
MAX_SPEED = 0.07
MIN_ACT, MAX_ACT = -1.0, 1.0
PRIEST = 0.5
PRIEST_THRESHOLD = 0.2
POWER = 0.0015


class CarOracle:
    def __init__(self, num_envs, time_limit=160, num_actions=None, draws=None):
        self.num_envs = int(num_envs)
        self.time_limit = time_limit
        self.rng = draws if draws is not None else GeneratorDraws()
        self.s = np.zeros((self.num_envs, 3), dtype=np.float32)
        self.elapsed = np.zeros(self.num_envs, dtype=np.int64)
        self.heavens = np.ones(self.num_envs, dtype=np.float32)
        self.priests = np.full(self.num_envs, PRIEST)
        self.hells = -self.heavens
        # DiscreteActionCarVecEnv (:286-303): evenly spaced forces, float64
        self.action_table = None if num_actions is None else np.linspace(MIN_ACT, MAX_ACT, num_actions)
        self.draws = self._blank_draws()

    def _blank_draws(self):
        b = self.num_envs
        return {"reset_pos": np.zeros(b, np.float64), "heaven": np.zeros(b, np.int8), "priest": np.zeros(b, np.int8)}

    @property
    def state(self):
        return {"s": self.s.copy(), "elapsed": self.elapsed.copy(), "heavens": self.heavens.copy(),
                "priests": self.priests.copy()}

    def set_state(self, s, elapsed, heavens, priests):
        self.s = np.array(s, dtype=np.float32)
        self.elapsed = np.array(elapsed, dtype=np.int64)
        self.heavens = np.array(heavens, dtype=np.float32)
        self.hells = -self.heavens
        self.priests = np.array(priests, dtype=np.float64)

    def _respawn(self, mask):
        """car_flag.py:97-112 — draw order: position uniform(-0.2, 0.2), heaven side, priest side"""
        b = int(mask.sum())
        if not b:
            return
        pos = self.rng.uniform(-0.2, 0.2, (b, 1))
        self.s[mask] = np.concatenate((pos, np.zeros((b, 2), dtype=np.float32)), axis=-1)
        self.elapsed[mask] = 0
        heaven = self.rng.choice([-1, 1], b)
        self.heavens[mask] = heaven
        self.hells[mask] = -self.heavens[mask]
        priest = self.rng.choice([-PRIEST, PRIEST], b)
        self.priests[mask] = priest
        self.draws["reset_pos"][mask] = pos[:, 0]
        self.draws["heaven"][mask] = heaven
        self.draws["priest"][mask] = np.sign(priest)

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.rng.reseed(seed)
        self.draws = self._blank_draws()
        self._respawn(np.ones(self.num_envs, dtype=bool))
        return self.s, {}

    def step(self, actions):
        """car_flag.py:114-141"""
        actions = np.asarray(actions)
        if self.action_table is not None:
            actions = self.action_table[actions]
        self.draws = self._blank_draws()
        self.elapsed += 1
        force = np.clip(actions.flatten(), MIN_ACT, MAX_ACT)
        vel = np.clip(self.s[:, 1] + force * POWER, -MAX_SPEED, MAX_SPEED)
        pos = np.clip(self.s[:, 0] + vel, MIN_POS, MAX_POS)
        vel[(pos == MIN_POS) & (vel < 0)] = 0
        done = np.abs(pos) >= 1.0
        side = np.sign(pos)
        rew = np.zeros(self.num_envs, dtype=np.float32)
        rew[(side == self.heavens) & done] = 1.0
        rew[(side == self.hells) & done] = -1.0
        truncated = self.elapsed >= self.time_limit
        indicator = np.where((pos >= self.priests - PRIEST_THRESHOLD) & (pos <= self.priests + PRIEST_THRESHOLD),
                             self.heavens, 0.0)
        keep = ~done
        self.s[keep] = np.column_stack((pos[keep], vel[keep], indicator[keep]))
        self._respawn(done | truncated)
        return self.s, rew, done, truncated, {}

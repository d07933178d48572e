"""Vectorized point-mass Tag on B200 — host side.

The reference's ``AntTagEnv`` (gym_po/envs/ant_tag.py) is a single MuJoCo ant chasing an evading target.
MuJoCo is third-party physics outside the reference repository; what this env keeps are the reference's
*pursuit rules* — target motion relative to the pursuer (:105-123), tag radius 1.5 (:147-150), visibility
radius 3.0 (:153), cage 4.5, spawn at distance > 5.0 (:94-100), 500-step time limit through gymnasium's
TimeLimit (envs/__init__.py:15-19) — vectorized over ``num_envs`` point agents that move with the CROOMS
motion model ``pos += (a + N(0, action_std^2)) * action_power`` clipped to the arena's inner walls
(+-5.0, assets/ant_tag_small.xml:72-83).  Observation: the target's (x, y) when closer than 3.0, else zeros.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _native as N
from .._device_env import DeviceVecEnv
from ..spaces import Box, batch_space

__all__ = ["TagVecEnv"]


class TagVecEnv(DeviceVecEnv):
    metadata = {"name": "PointTag", "render_modes": []}
    cage_max_xy, visible_radius, tag_radius, min_distance, target_step = 4.5, 3.0, 1.5, 5.0, 0.5

    def __init__(self, num_envs: int, time_limit: int = 500, action_std: float = 0.2, action_power: float = 1.0,
                 render_mode: Optional[str] = None, *, device=None, rng_mode: str = "philox", seed: Optional[int] = None,
                 env_offset: int = 0, track_stats: bool = False, action_dtype=torch.float32, precision: str = "float64"):
        if precision not in ("float64", "float32"):
            raise ValueError("precision must be 'float64' or 'float32'")
        self._precision = precision
        self.num_envs = int(num_envs)
        self.time_limit = time_limit
        self.render_mode = render_mode
        self.single_action_space = Box(-1.0, 1.0, (2,))
        self.single_observation_space = Box(-np.inf, np.inf, (2,), dtype=np.float64)
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        cfg = N.GptConfig()
        cfg.family = N.FAMILY_TAG
        cfg.time_limit = int(time_limit)
        cfg.c_action_std, cfg.c_action_power = action_std, action_power
        cfg.c_action_f64 = int(action_dtype == torch.float64)
        cfg.c_state_f32 = int(self._precision == "float32")
        self._create(cfg, device=device, rng_mode=rng_mode, seed=seed, env_offset=env_offset, track_stats=track_stats)

    @property
    def agent_xy(self) -> torch.Tensor:
        return self._arrays["agent"][: self.num_envs]

    @property
    def target_xy(self) -> torch.Tensor:
        return self._arrays["target"][: self.num_envs]

    @property
    def elapsed(self) -> torch.Tensor:
        return self._arrays["elapsed"][: self.num_envs]

    def get_state(self):
        return {"agent": self.agent_xy.clone(), "target": self.target_xy.clone(), "elapsed": self.elapsed.clone()}

    def set_state(self, agent, target, elapsed):
        b = self.num_envs
        self._arrays["agent"][:b].copy_(torch.as_tensor(np.asarray(agent, dtype=np.float64)).to(self._arrays["agent"].dtype))
        self._arrays["target"][:b].copy_(torch.as_tensor(np.asarray(target, dtype=np.float64)).to(self._arrays["agent"].dtype))
        self._arrays["elapsed"][:b].copy_(torch.as_tensor(np.asarray(elapsed)).to(torch.int32))

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        return self._reset(seed), {}

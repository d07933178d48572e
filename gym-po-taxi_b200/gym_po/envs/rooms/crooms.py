"""Vectorized continuous-position ROOMS on B200 — host side.

Drop-in for the reference's ``CRoomsEnv`` (gym_po/envs/rooms/crooms.py:91-338): same kwargs, ``reset()``
returns the observation only (:266), ``step`` the 5-tuple with same-step autoreset.  Positions are
float64 (y, x) like the reference; the kernel (csrc/gpt_crooms.cu) uses the reference's operation order
without FMA contraction, so replayed trajectories are bit-identical.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from ... import _native as N
from ..._device_env import DeviceVecEnv
from ...spaces import Box, Discrete, batch_space
from .layouts import LAYOUTS, layout_to_np, np_to_grid
from .rooms import ACTIONS_CARDINAL, ACTIONS_ORDINAL, fixed_goal_yx, resolve_obs_kind, slip_cumsum

__all__ = ["CRoomsEnv"]


class CRoomsEnv(DeviceVecEnv):
    """Continuous ROOMS domain, vectorized, fused CUDA step."""

    metadata = {"name": "CRooms", "render.modes": ["human", "rgb_array"], "video.frames_per_second": 10}

    def __init__(self, num_envs: int, layout: str = "4", time_limit: int = 500, use_velocity: bool = False,
                 cell_size: float = 1.0, obs_type: str = "mdp", obs_m: int = 3,
                 action_failure_probability: float = 0.2, action_type: str = "yx", action_std: float = 0.2,
                 action_power: float = 1.0, agent_xy: Optional[Sequence[int]] = None,
                 goal_xy: Optional[Sequence[int]] = (0, 0), step_reward: float = 0.0, wall_reward: float = 0.0,
                 goal_reward: float = 1.0, goal_threshold: float = 0.5, render_mode: Optional[str] = None, *,
                 device=None, rng_mode: str = "philox", seed: Optional[int] = None, env_offset: int = 0,
                 track_stats: bool = False, action_dtype=torch.float32, precision: str = "float64", **kwargs):
        assert layout in LAYOUTS
        if agent_xy is not None:  # raises in the reference as well (crooms.py:232-235)
            raise ValueError("agent_xy is not supported (it raises in the reference as well)")
        self.metadata = dict(self.metadata, name=f"CRooms__{layout}__{action_type}__{obs_type}")
        if precision not in ("float64", "float32"):
            raise ValueError("precision must be 'float64' (bit-exact vs the reference) or 'float32' (fast)")
        self._precision = precision
        self.num_envs = int(num_envs)
        self.grid = np_to_grid(layout_to_np(LAYOUTS[layout]))
        self.gridshape = np.array(self.grid.shape)
        self.valid_states = np.flatnonzero(self.grid >= 0)
        kind, n, self.single_observation_space = resolve_obs_kind(obs_type, self.grid, obs_m, continuous=True)
        self._obs_kind, self._obs_n = kind, n
        self.max_velocity = 5.0
        cfg = N.GptConfig()
        keep = []
        if action_type == "yx":
            self.single_action_space = Box(-1.0, 1.0, (2,))
            cfg.rooms_n_actions = 0
            cfg.c_action_f64 = int(action_dtype == torch.float64)
        else:
            acts = ACTIONS_CARDINAL if action_type == "cardinal" else ACTIONS_ORDINAL
            self.single_action_space = Discrete(acts.shape[0])
            cfg.rooms_n_actions = acts.shape[0]
            thr = slip_cumsum(acts.shape[0], action_failure_probability)
            cfg.rooms_slip_cumsum = thr.ctypes.data_as(C.POINTER(C.c_double))
            keep.append(thr)
        cfg.c_state_f32 = int(precision == "float32")
        self.use_velocity = bool(use_velocity)
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        self.time_limit = time_limit
        self.step_reward, self.goal_reward, self.wall_reward = step_reward, goal_reward, wall_reward
        self.goal_threshold, self.cell_size, self.action_power = goal_threshold, cell_size, action_power
        self.render_mode = render_mode
        self.fixed_goal = None if goal_xy is None else fixed_goal_yx(self.grid, layout, goal_xy)
        if self.fixed_goal is not None and kind in (N.OBS_ROOM_GOAL, N.OBS_MDP_GOAL) and not (
                self.fixed_goal[0] < self.grid.shape[0] and self.fixed_goal[1] < self.grid.shape[1]):
            # the reference's default goal for layouts '32'/'32b' lies outside the grid; it indexes the
            # table with it at reset() and raises IndexError (rooms.py:27,42-45)
            raise IndexError(f"fixed goal {self.fixed_goal} is outside the {self.grid.shape} grid")

        cfg.family = N.FAMILY_CROOMS
        cfg.time_limit = int(time_limit)
        cfg.rooms_h, cfg.rooms_w = self.grid.shape
        grid8 = np.ascontiguousarray(self.grid, dtype=np.int8)
        keep.append(grid8)
        cfg.rooms_grid = grid8.ctypes.data_as(C.POINTER(C.c_int8))
        cfg.rooms_obs_kind, cfg.rooms_obs_n = kind, n
        cfg.rooms_goal_y, cfg.rooms_goal_x = self.fixed_goal if self.fixed_goal is not None else (-1, -1)
        cfg.rooms_step_reward, cfg.rooms_wall_reward, cfg.rooms_goal_reward = step_reward, wall_reward, goal_reward
        cfg.c_cell_size, cfg.c_action_std, cfg.c_action_power = cell_size, action_std, action_power
        cfg.c_goal_threshold = goal_threshold
        cfg.c_use_velocity = int(self.use_velocity)
        self._create(cfg, device=device, rng_mode=rng_mode, seed=seed, env_offset=env_offset, track_stats=track_stats,
                     keepalive=tuple(keep))

    def _shape_obs(self, obs):
        if self._obs_kind == N.OBS_GRID:
            return obs.reshape(obs.shape[0], self._obs_n, self._obs_n)
        return obs

    # ---- state (reference attributes agent_yx / goal_yx / agent_yx_velocity / elapsed) ----
    @property
    def agent_yx(self) -> torch.Tensor:
        return self._arrays["agent"][: self.num_envs]

    @property
    def goal_yx(self) -> torch.Tensor:
        if self.fixed_goal is not None:
            g = torch.tensor(self.fixed_goal, device=self.device, dtype=self._arrays["agent"].dtype) + 0.5
            return g.expand(self.num_envs, 2).clone()
        return self._arrays["goal"][: self.num_envs]

    @property
    def agent_yx_velocity(self) -> torch.Tensor:
        if not self.use_velocity:
            return torch.zeros((self.num_envs, 2), dtype=self._arrays["agent"].dtype, device=self.device)
        return self._arrays["velocity"][: self.num_envs]

    @property
    def elapsed(self) -> torch.Tensor:
        return self._arrays["elapsed"][: self.num_envs]

    def get_state(self):
        return {"agent": self.agent_yx.clone(), "goal": self.goal_yx.clone(), "velocity": self.agent_yx_velocity.clone(),
                "elapsed": self.elapsed.clone()}

    def set_state(self, agent, goal, velocity, elapsed):
        b = self.num_envs
        self._arrays["agent"][:b].copy_(torch.as_tensor(np.asarray(agent, dtype=np.float64)).to(self._arrays["agent"].dtype))
        if self.fixed_goal is None:
            self._arrays["goal"][:b].copy_(torch.as_tensor(np.asarray(goal, dtype=np.float64)).to(self._arrays["agent"].dtype))
        if self.use_velocity:
            self._arrays["velocity"][:b].copy_(torch.as_tensor(np.asarray(velocity, dtype=np.float64)).to(self._arrays["agent"].dtype))
        self._arrays["elapsed"][:b].copy_(torch.as_tensor(np.asarray(elapsed)).to(torch.int32))

    def seed(self, seed: Optional[int] = None):
        """Reference API (crooms.py:246-249): reseed; takes effect at the next ``reset(seed=...)``-less call."""
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1, np.uint64)[0])
        self._pending_seed = int(seed)
        return seed

    def reset(self, *, seed: Optional[int] = None, return_info: bool = False, options: Optional[dict] = None):
        if seed is None and getattr(self, "_pending_seed", None) is not None:
            seed, self._pending_seed = self._pending_seed, None
        return self._reset(seed)

"""Vectorized discrete ROOMS / FourRooms on B200 — host side.

Drop-in for the reference's ``RoomsEnv`` (gym_po/envs/rooms/rooms.py:71-226): same constructor kwargs,
``reset()`` returns the observation only (reference :189), ``step(action)`` returns the 5-tuple with
same-step autoreset.  The step runs in one fused CUDA kernel (csrc/gpt_rooms_kernel.cuh).

Observation dtypes are compact (the reference returns int64 / float64): scalar observations int32,
vector observations and the n x n window uint8.  Values are identical.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from ... import _native as N
from ..._device_env import DeviceVecEnv
from ...spaces import Box, Discrete, batch_space
from .layouts import ENDS, LAYOUTS, STARTS, layout_to_np, np_to_grid

__all__ = ["RoomsEnv", "resolve_obs_kind"]

# compass tables in the reference's order (rooms/action_utils.py:16-29)
ACTIONS_ORDINAL = np.array([[-1, 0], [-1, 1], [0, 1], [1, 1], [1, 0], [1, -1], [0, -1], [-1, -1]])
ACTIONS_CARDINAL = ACTIONS_ORDINAL[::2]


def resolve_obs_kind(obs_type: str, grid: np.ndarray, obs_n: int, continuous: bool = False):
    """Substring dispatch in the reference's order room -> mdp -> hansen -> grid (rooms.py:19-67).
    Returns (GPT_OBS_* kind, n, single-observation space)."""
    vec, has_goal = "vector" in obs_type, "goal" in obs_type
    n_cells = int((grid >= 0).sum())
    n_rooms = len(np.unique(grid)) - 1
    if "room" in obs_type:
        if has_goal:
            return N.OBS_ROOM_GOAL, 0, Discrete(n_rooms ** 2)
        return N.OBS_ROOM, 0, Discrete(n_rooms)
    if "mdp" in obs_type:
        if vec:
            if continuous:
                hi = np.array(grid.shape) - 1 - 1e-6
                return (N.OBS_VEC_MDP_GOAL, 0, Box(1.0, np.tile(hi, 2), (4,))) if has_goal else (N.OBS_VEC_MDP, 0, Box(1.0, hi, (2,)))
            hi = np.array(grid.shape) - 2
            if has_goal:
                return N.OBS_VEC_MDP_GOAL, 0, Box(1, np.tile(hi, 2), (4,), dtype=int)
            return N.OBS_VEC_MDP, 0, Box(1, hi, (2,), dtype=int)
        if has_goal:
            return N.OBS_MDP_GOAL, 0, Discrete(n_cells ** 2)
        return N.OBS_MDP, 0, Discrete(n_cells)
    if "hansen" in obs_type:
        k = 8 if "8" in obs_type else 4
        if vec:
            if has_goal:
                return N.OBS_VEC_HANSEN_GOAL, k, Box(0, 2, (k,), dtype=int)
            return N.OBS_VEC_HANSEN, k, Box(0, 1, (k,), dtype=int)
        return N.OBS_HANSEN, k, Discrete(2 ** k * (k + 1))
    if "grid" in obs_type:
        return N.OBS_GRID, int(obs_n), Box(0, 2, (obs_n, obs_n), dtype=int)
    raise NotImplementedError("Observation type not recognized")


def slip_cumsum(n: int, p_fail: float) -> np.ndarray:
    """Row-wise float64 cumsum of the slip matrix, computed with numpy exactly like the reference
    (action_utils.py:38-48, :85-87) so the thresholds are bit-identical."""
    m = np.full((n, n), p_fail / (n - 1), dtype=np.float64)
    np.fill_diagonal(m, 1 - p_fail)
    return np.ascontiguousarray(m.cumsum(axis=1))


def fixed_goal_yx(grid, layout, goal_xy):
    gy, gx = int(goal_xy[1]), int(goal_xy[0])
    if grid[gy, gx] < 0:
        ex, ey = ENDS[layout[:-1] if "b" in layout else layout]
        gy, gx = ey, ex
    return gy, gx


class RoomsEnv(DeviceVecEnv):
    """Basic ROOMS domain, vectorized, fused CUDA step."""

    metadata = {"name": "Rooms", "render_modes": ["human", "rgb_array"], "render_fps": 10}

    def __init__(self, num_envs: int, layout: str = "4", time_limit: int = 500, obs_type: str = "mdp", obs_n: int = 3,
                 action_failure_probability: float = 0.2, action_type: str = "ordinal",
                 agent_xy: Optional[Sequence[int]] = None, goal_xy: Optional[Sequence[int]] = (0, 0),
                 step_reward: float = 0.0, wall_reward: float = 0.0, goal_reward: float = 1.0,
                 render_mode: Optional[str] = None, *, device=None, rng_mode: str = "philox",
                 seed: Optional[int] = None, env_offset: int = 0, track_stats: bool = False, **kwargs):
        assert layout in LAYOUTS
        if agent_xy is not None:  # the reference raises ValueError for this kwarg (rooms.py:164-166)
            raise ValueError("agent_xy is not supported (it raises in the reference as well)")
        self.metadata = dict(self.metadata, name=f"Rooms__{layout}__{action_type}__{obs_type}")
        self.num_envs = int(num_envs)
        self.grid = np_to_grid(layout_to_np(LAYOUTS[layout]))
        self.gridshape = np.array(self.grid.shape)
        self.valid_states = np.flatnonzero(self.grid >= 0)
        kind, n, self.single_observation_space = resolve_obs_kind(obs_type, self.grid, obs_n)
        self._obs_kind, self._obs_n = kind, n
        self.actions = ACTIONS_CARDINAL if action_type == "cardinal" else ACTIONS_ORDINAL
        self.single_action_space = Discrete(self.actions.shape[0])
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        self.time_limit = time_limit
        self.step_reward, self.goal_reward, self.wall_reward = step_reward, goal_reward, wall_reward
        self.render_mode = render_mode
        n_act = self.actions.shape[0]
        self.action_matrix = np.full((n_act, n_act), action_failure_probability / (n_act - 1), dtype=np.float64)
        np.fill_diagonal(self.action_matrix, 1 - action_failure_probability)
        self.fixed_goal = None if goal_xy is None else fixed_goal_yx(self.grid, layout, goal_xy)
        if self.fixed_goal is not None and kind in (N.OBS_ROOM_GOAL, N.OBS_MDP_GOAL) and not (
                self.fixed_goal[0] < self.grid.shape[0] and self.fixed_goal[1] < self.grid.shape[1]):
            # the reference's default goal for layouts '32'/'32b' lies outside the grid; it indexes the
            # table with it at reset() and raises IndexError (rooms.py:27,42-45)
            raise IndexError(f"fixed goal {self.fixed_goal} is outside the {self.grid.shape} grid")

        cfg = N.GptConfig()
        cfg.family = N.FAMILY_ROOMS
        cfg.time_limit = int(time_limit)
        cfg.rooms_h, cfg.rooms_w = self.grid.shape
        grid8 = np.ascontiguousarray(self.grid, dtype=np.int8)
        cfg.rooms_grid = grid8.ctypes.data_as(C.POINTER(C.c_int8))
        cfg.rooms_n_actions = n_act
        thr = slip_cumsum(n_act, action_failure_probability)
        cfg.rooms_slip_cumsum = thr.ctypes.data_as(C.POINTER(C.c_double))
        cfg.rooms_obs_kind, cfg.rooms_obs_n = kind, n
        cfg.rooms_goal_y, cfg.rooms_goal_x = self.fixed_goal if self.fixed_goal is not None else (-1, -1)
        cfg.rooms_step_reward, cfg.rooms_wall_reward, cfg.rooms_goal_reward = step_reward, wall_reward, goal_reward
        self._create(cfg, device=device, rng_mode=rng_mode, seed=seed, env_offset=env_offset, track_stats=track_stats,
                     keepalive=(grid8, thr))

    def _shape_obs(self, obs):
        if self._obs_kind == N.OBS_GRID:
            return obs.reshape(obs.shape[0], self._obs_n, self._obs_n)
        return obs

    # ---- state access (reference attributes agent_yx / goal_yx / elapsed, rooms.py:185-187) ----
    def _cells_to_yx(self, cells):
        w = int(self.grid.shape[1])
        c = cells.to(torch.int64)
        return torch.stack((c // w, c % w), -1)

    @property
    def agent_yx(self) -> torch.Tensor:
        return self._cells_to_yx(self._arrays["pos"][: self.num_envs])

    @property
    def goal_yx(self) -> torch.Tensor:
        if self.fixed_goal is not None:
            return torch.tensor(self.fixed_goal, device=self.device).expand(self.num_envs, 2).clone()
        return self._cells_to_yx(self._arrays["goal"][: self.num_envs])

    @property
    def elapsed(self) -> torch.Tensor:
        return self._arrays["elapsed"][: self.num_envs]

    def get_state(self):
        return {"agent": self.agent_yx, "goal": self.goal_yx, "elapsed": self.elapsed.clone()}

    def set_state(self, agent, goal, elapsed):
        b, w = self.num_envs, int(self.grid.shape[1])
        agent = np.asarray(agent)
        for name, yx in (("agent", agent),) + ((("goal", np.asarray(goal)),) if self.fixed_goal is None else ()):
            if yx.size and (yx.min() < 0 or (yx[:, 0] >= self.grid.shape[0]).any() or (yx[:, 1] >= w).any()):
                raise ValueError(f"set_state: {name} positions must lie inside the {self.grid.shape[0]}x{w} grid")
        self._arrays["pos"][:b].copy_(torch.as_tensor(agent[:, 0] * w + agent[:, 1]).to(torch.int16))
        if self.fixed_goal is None:
            goal = np.asarray(goal)
            self._arrays["goal"][:b].copy_(torch.as_tensor(goal[:, 0] * w + goal[:, 1]).to(torch.int16))
        self._arrays["elapsed"][:b].copy_(torch.as_tensor(np.asarray(elapsed)).to(torch.int32))

    # ---- gym API ----------------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """Reset all environments; returns the observation only, like the reference (:189)."""
        return self._reset(seed)

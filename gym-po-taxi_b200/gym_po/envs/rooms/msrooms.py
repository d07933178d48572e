"""Vectorized multistory FourRooms on B200 — host side.

Drop-in for the reference's ``MultistoryFourRoomsEnv`` (gym_po/envs/rooms/msrooms.py:257-432): same constructor
kwargs, ``reset()`` returns ``(obs, {})`` (reference :371-383), ``step(action)`` returns the 5-tuple with same-step
autoreset.  The step runs in one fused CUDA kernel (csrc/gpt_msrooms.cu); the stair teleport (:419-428) is folded
into the kernel's move table.

Reference behaviour kept on purpose: every walkable cell reads as 2 ("stairs") in the Hansen
observations (:154-155, :184-185); a goal given by the caller is always replaced by ``END_XYZ`` on the top floor
(:340-346 — ``grid[goal] <= 3`` holds for every cell); the 'room' observations use the raw grid value and
``n = grid.max() - 4`` (:203-213), so their declared space size is not positive; ``agent_xyz`` raises (:354).

Observation dtypes are compact (the reference returns int64 / float64): scalar observations int32, vector
observations uint8.  Values are identical.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from ... import _native as N
from ..._device_env import DeviceVecEnv
from ...spaces import Box, Discrete, batch_space
from .rooms import slip_cumsum

__all__ = ["MultistoryFourRoomsEnv", "FR_MAP", "END_XYZ", "START_XYZ"]

END_XYZ = (9, 7, -1)    # east hallway, top floor (msrooms.py:17)
START_XYZ = (1, 1, 0)   # (msrooms.py:18)
UPSTAIRS_YX = (1, 11)   # NE corner (msrooms.py:21-24)
DOWNSTAIRS_YX = (11, 1)  # SW corner
WALL, WALK, STAIR_DOWN, STAIR_UP = 0, 1, 2, 3   # GR_CNST (msrooms.py:27-31)

# 13x13 FourRooms: 0 = wall, 1..4 = rooms clockwise from the NE (msrooms.py:50-66).  Built from room rectangles
# and the four doorways rather than spelled out.
FR_MAP = np.zeros((13, 13), dtype=np.int64)
FR_MAP[1:6, 1:6] = 4
FR_MAP[1:7, 7:12] = 1
FR_MAP[8:12, 7:12] = 2
FR_MAP[7:12, 1:6] = 3
FR_MAP[3, 6] = 4    # west-east doorway, north
FR_MAP[6, 2] = 3    # north-south doorway, west
FR_MAP[7, 9] = 1    # north-south doorway, east
FR_MAP[10, 6] = 2   # west-east doorway, south

ACTIONS_ORDINAL_Z = np.array([[0, -1, 0], [0, -1, 1], [0, 0, 1], [0, 1, 1], [0, 1, 0], [0, 1, -1], [0, 0, -1], [0, -1, -1]])
ACTIONS_CARDINAL_Z = ACTIONS_ORDINAL_Z[::2]


def multistory_grid(floor_map: np.ndarray, floors: int) -> np.ndarray:
    """[S,H,W] walk map: 0 wall, 1 walkable, 2 stair down, 3 stair up (msrooms.py:69-90)."""
    walk = (np.asarray(floor_map) > 0).astype(np.int64)
    ms = np.repeat(walk[None], floors, 0)
    if floors > 1:
        ms[1:, DOWNSTAIRS_YX[0], DOWNSTAIRS_YX[1]] = STAIR_DOWN
        ms[:-1, UPSTAIRS_YX[0], UPSTAIRS_YX[1]] = STAIR_UP
    return ms


def resolve_ms_obs_kind(obs_type: str, grid: np.ndarray):
    """Substring dispatch room -> mdp -> hansen (msrooms.py:192-254) -> (GPT_OBS_* kind, n, single space)."""
    vec, has_goal = "vector" in obs_type, "goal" in obs_type
    a_max = np.array(grid.shape) - 2
    a_max[0] += 1
    a_min = np.array([0, 1, 1])
    if "room" in obs_type:
        assert not vec
        n = int(grid.max()) - 4
        return (N.OBS_ROOM_GOAL, 0, Discrete(int(n ** 2))) if has_goal else (N.OBS_ROOM, 0, Discrete(n))
    if "mdp" in obs_type:
        if vec:
            if has_goal:
                return N.OBS_VEC_MDP_GOAL, 0, Box(np.tile(a_min, 2), np.tile(a_max, 2), (6,), dtype=int)
            return N.OBS_VEC_MDP, 0, Box(a_min, a_max, (3,), dtype=int)
        n = int((grid > 0).sum())
        return (N.OBS_MDP_GOAL, 0, Discrete(n ** 2)) if has_goal else (N.OBS_MDP, 0, Discrete(n))
    if "hansen" in obs_type:
        k = 8 if "8" in obs_type else 4
        if vec:
            return (N.OBS_VEC_HANSEN_GOAL, k, Box(0, 3, (k,), dtype=int)) if has_goal else (N.OBS_VEC_HANSEN, k, Box(0, 2, (k,), dtype=int))
        return N.OBS_HANSEN, k, Discrete(int(3 ** k * (k + 1)))
    raise NotImplementedError("Observation type not recognized")


class MultistoryFourRoomsEnv(DeviceVecEnv):
    """Vectorized multistory FourRooms, fused CUDA step."""

    metadata = {"name": "MultistoryFourRoomsV2", "render_modes": ["human", "rgb_array"], "render_fps": 10}

    def __init__(self, num_envs: int, grid_z: int = 1, floor_map: np.ndarray = FR_MAP, time_limit: int = 500,
                 obs_type: str = "mdp", obs_n: int = 3, action_failure_probability: float = 1.0 / 3,
                 action_type: str = "cardinal", agent_xyz: Optional[Sequence[int]] = None,
                 goal_xyz: Optional[Sequence[int]] = END_XYZ, step_reward: float = 0.0, wall_reward: float = 0.0,
                 goal_reward: float = 1.0, render_mode: Optional[str] = None, *, device=None, rng_mode: str = "philox",
                 seed: Optional[int] = None, env_offset: int = 0, **kwargs):
        if agent_xyz is not None:  # the reference indexes the grid with an array here and raises (msrooms.py:352-354)
            raise ValueError("agent_xyz is not supported (it raises in the reference as well)")
        floor_map = np.asarray(floor_map)
        self.grid = multistory_grid(floor_map, int(grid_z))
        self.metadata = dict(self.metadata, name=f"MultistoryFourRoomsV2{grid_z}__{action_type}__{obs_type}")
        self.gridshape = np.array(self.grid.shape)
        kind, n, self.single_observation_space = resolve_ms_obs_kind(obs_type, self.grid)
        self._obs_kind, self._obs_n = kind, n
        self.valid_states = np.flatnonzero(self.grid > WALL)
        zs = np.unravel_index(self.valid_states, self.grid.shape)[0]
        self.valid_agent_states = self.valid_states[zs == 0]
        self.valid_goal_states = self.valid_states[zs == self.gridshape[0] - 1]
        self.render_mode = render_mode
        self.actions = ACTIONS_CARDINAL_Z if action_type == "cardinal" else ACTIONS_ORDINAL_Z
        self.num_envs = int(num_envs)
        self.single_action_space = Discrete(self.actions.shape[0])
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        try:
            self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        except Exception:  # 'room' observations declare a non-positive size in the reference too
            self.observation_space = None
        self.time_limit = time_limit
        self.step_reward, self.goal_reward, self.wall_reward = step_reward, goal_reward, wall_reward
        n_act = self.actions.shape[0]
        self.action_matrix = np.full((n_act, n_act), action_failure_probability / (n_act - 1), dtype=np.float64)
        np.fill_diagonal(self.action_matrix, 1 - action_failure_probability)
        if goal_xyz is not None:  # any given goal ends up at END_XYZ on the top floor (msrooms.py:340-346)
            self.fixed_goal = (int(self.gridshape[0]) - 1, END_XYZ[1], END_XYZ[0])
        else:
            self.fixed_goal = None

        cfg = N.GptConfig()
        cfg.family = N.FAMILY_MSROOMS
        cfg.time_limit = int(time_limit)
        cfg.rooms_h, cfg.rooms_w = floor_map.shape
        floor8 = np.ascontiguousarray((floor_map > 0).astype(np.int8))
        cfg.rooms_grid = floor8.ctypes.data_as(C.POINTER(C.c_int8))
        cfg.rooms_n_actions = n_act
        thr = slip_cumsum(n_act, action_failure_probability)
        cfg.rooms_slip_cumsum = thr.ctypes.data_as(C.POINTER(C.c_double))
        cfg.rooms_obs_kind, cfg.rooms_obs_n = kind, n
        cfg.rooms_step_reward, cfg.rooms_wall_reward, cfg.rooms_goal_reward = step_reward, wall_reward, goal_reward
        cfg.ms_floors = int(grid_z)
        cfg.ms_goal_cell = -1 if self.fixed_goal is None else int(np.ravel_multi_index(self.fixed_goal, self.grid.shape))
        cfg.ms_up_y, cfg.ms_up_x = UPSTAIRS_YX
        cfg.ms_down_y, cfg.ms_down_x = DOWNSTAIRS_YX
        self._create(cfg, device=device, rng_mode=rng_mode, seed=seed, env_offset=env_offset, keepalive=(floor8, thr))

    # ---- state access (reference attributes agent_zyx / goal_zyx / elapsed, msrooms.py:379-381) ----
    def _cells_to_zyx(self, cells):
        _, h, w = (int(v) for v in self.grid.shape)
        c = cells.to(torch.int64)
        return torch.stack((c // (h * w), (c // w) % h, c % w), -1)

    def _zyx_to_cells(self, zyx):
        zyx = np.asarray(zyx)
        return torch.as_tensor(np.ravel_multi_index((zyx[:, 0], zyx[:, 1], zyx[:, 2]), self.grid.shape)).to(torch.int16)

    @property
    def agent_zyx(self) -> torch.Tensor:
        return self._cells_to_zyx(self._arrays["pos"][: self.num_envs])

    @property
    def goal_zyx(self) -> torch.Tensor:
        if self.fixed_goal is not None:
            return torch.tensor(self.fixed_goal, device=self.device).expand(self.num_envs, 3).clone()
        return self._cells_to_zyx(self._arrays["goal"][: self.num_envs])

    @property
    def elapsed(self) -> torch.Tensor:
        return self._arrays["elapsed"][: self.num_envs]

    def get_state(self):
        return {"agent": self.agent_zyx, "goal": self.goal_zyx, "elapsed": self.elapsed.clone()}

    def set_state(self, agent, goal, elapsed):
        b = self.num_envs
        self._arrays["pos"][:b].copy_(self._zyx_to_cells(agent))
        if self.fixed_goal is None:
            self._arrays["goal"][:b].copy_(self._zyx_to_cells(goal))
        self._arrays["elapsed"][:b].copy_(torch.as_tensor(np.asarray(elapsed)).to(torch.int32))

    # ---- gym API ----------------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """Reset all environments; returns ``(obs, {})`` like the reference (:383)."""
        return self._reset(seed), {}

    def render(self):
        raise NotImplementedError  # as in the reference (msrooms.py:430-432)

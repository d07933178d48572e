"""TEST INFRASTRUCTURE — numpy oracle for the multistory FourRooms env.

Restates ``MultistoryFourRoomsEnv`` of the reference (gym_po/envs/rooms/msrooms.py: map + stairs :50-90,
observation functions :131-254, ctor :266-369, ``reset`` :371-383, ``_reset_some`` :385-390, ``step`` :392-413,
``_out_of_bounds`` :415-417, ``_transit_stairs`` :419-428).  The reference file only imports after the
signature repair of SURVEY Appendix C; its quirks are kept: every walkable cell reads as "stairs" (2) in the
Hansen observations (:154-155, :184-185), a user goal is always replaced by END_XYZ on the top floor
(:340-346: ``grid[goal] <= 3`` holds for every cell), the 'room' observation is the raw grid value.
"""
from __future__ import annotations

import numpy as np

from .draws import GeneratorDraws
from .rooms import DIRS4, DIRS8, slip_matrix, slip_sample

# 13x13 FourRooms, 0 = wall, 1..4 = rooms (msrooms.py:50-66)
FR_MAP = np.array([
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 4, 4, 4, 4, 4, 0, 1, 1, 1, 1, 1, 0],
    [0, 4, 4, 4, 4, 4, 0, 1, 1, 1, 1, 1, 0],
    [0, 4, 4, 4, 4, 4, 4, 1, 1, 1, 1, 1, 0],
    [0, 4, 4, 4, 4, 4, 0, 1, 1, 1, 1, 1, 0],
    [0, 4, 4, 4, 4, 4, 0, 1, 1, 1, 1, 1, 0],
    [0, 0, 3, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0],
    [0, 3, 3, 3, 3, 3, 0, 0, 0, 1, 0, 0, 0],
    [0, 3, 3, 3, 3, 3, 0, 2, 2, 2, 2, 2, 0],
    [0, 3, 3, 3, 3, 3, 0, 2, 2, 2, 2, 2, 0],
    [0, 3, 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 0],
    [0, 3, 3, 3, 3, 3, 0, 2, 2, 2, 2, 2, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
])
WALL, WALK, STAIR_DOWN, STAIR_UP = 0, 1, 2, 3   # GR_CNST (msrooms.py:27-31); "goal" shares the value 1
MAX_CONST = 3
END_XYZ = (9, 7, -1)
UP_YX = np.array([1, 11])     # stair-up cell (NE); arriving from below lands on DOWN_YX of the floor above
DOWN_YX = np.array([11, 1])   # stair-down cell (SW)


def multistory_grid(floor_map, floors):
    """msrooms.py:69-90 — [S,H,W] walk map with stairs"""
    walk = floor_map.copy()
    walk[floor_map > 0] = 1
    ms = np.stack([walk for _ in range(floors)], 0)
    if floors > 1:
        ms[1:, DOWN_YX[0], DOWN_YX[1]] = STAIR_DOWN
        ms[:-1, UP_YX[0], UP_YX[1]] = STAIR_UP
    return ms


def _dirs_z(n):
    d = DIRS4 if n == 4 else DIRS8
    return np.concatenate((np.zeros((len(d), 1), dtype=int), d), -1)


def _squares(agent, grid, n):
    nb = agent[:, None, :] + _dirs_z(n)[None]
    sq = grid[nb[..., 0], nb[..., 1], nb[..., 2]].copy()
    sq[(sq > 0) & (sq <= MAX_CONST)] = 2   # every walkable cell (1, 2, 3) aliases to "stairs"
    sq[sq > MAX_CONST] = 1
    return nb, sq


def hansen_scalar(agent, grid, goal, n):
    """msrooms.py:163-189 — base-3 digits times (index of the neighbour holding the goal)+1, float64"""
    nb, sq = _squares(agent, grid, n)
    env_i, dir_i = np.nonzero((goal[:, None, :] == nb).all(-1))
    mult = np.ones(goal.shape[0])
    mult[env_i] = dir_i + 1
    return sq.dot(3 ** np.arange(n)) * mult


def hansen_vector(agent, grid, goal, n):
    """msrooms.py:131-160"""
    nb, sq = _squares(agent, grid, n)
    if goal is not None:
        sq[(goal[:, None, :] == nb).all(-1)] = 3
    return sq


def make_obs_fn(obs_type, grid):
    """msrooms.py:192-254"""
    vec, has_goal = "vector" in obs_type, "goal" in obs_type
    at = lambda p: grid[p[:, 0], p[:, 1], p[:, 2]]
    if "room" in obs_type:
        assert not vec
        n = grid.max() - 4
        if has_goal:
            return lambda a, g: (at(a) - 4) + n * (at(g) - 4)
        return lambda a, g: at(a)
    if "mdp" in obs_type:
        if vec:
            if has_goal:
                return lambda a, g: np.concatenate((a, g), -1)
            return lambda a, g: a
        free = (grid - 1) >= 0
        n = int(free.sum())
        table = (free.cumsum() - 1).reshape(grid.shape)
        tat = lambda p: table[p[:, 0], p[:, 1], p[:, 2]]
        if has_goal:
            return lambda a, g: tat(a) + n * tat(g)
        return lambda a, g: tat(a)
    if "hansen" in obs_type:
        k = 8 if "8" in obs_type else 4
        if vec:
            if has_goal:
                return lambda a, g: hansen_vector(a, grid, g, k)
            return lambda a, g: hansen_vector(a, grid, None, k)
        return lambda a, g: hansen_scalar(a, grid, g, k)
    raise NotImplementedError("Observation type not recognized")


class MSRoomsOracle:
    def __init__(self, num_envs, grid_z=1, floor_map=FR_MAP, time_limit=500, obs_type="mdp", obs_n=3,
                 action_failure_probability=1.0 / 3, action_type="cardinal", agent_xyz=None, goal_xyz=END_XYZ,
                 step_reward=0.0, wall_reward=0.0, goal_reward=1.0, draws=None, **_):
        if agent_xyz is not None:
            raise ValueError("agent_xyz is not usable in the reference (array used as an index, msrooms.py:354)")
        self.num_envs = int(num_envs)
        self.grid = multistory_grid(np.asarray(floor_map), grid_z)
        self._obs_fn = make_obs_fn(obs_type, self.grid)
        shape = self.grid.shape
        walk = np.array(np.nonzero(self.grid > WALL))
        self.agent_cells = np.ravel_multi_index(walk[:, walk[0] == 0], shape)             # spawn: bottom floor
        self.goal_cells = np.ravel_multi_index(walk[:, walk[0] == shape[0] - 1], shape)   # random goal: top floor
        self.dirs = _dirs_z(4 if action_type == "cardinal" else 8)
        self.n_actions = len(self.dirs)
        self.P = slip_matrix(self.n_actions, action_failure_probability)
        self.time_limit = time_limit
        self.step_reward, self.wall_reward, self.goal_reward = step_reward, wall_reward, goal_reward
        if goal_xyz is not None:   # always ends up at END_XYZ on the top floor (msrooms.py:340-346)
            self.fixed_goal = np.array([shape[0] - 1, END_XYZ[1], END_XYZ[0]])
        else:
            self.fixed_goal = None
        self.rng = draws if draws is not None else GeneratorDraws()
        self.draws = {}

    def _blank_draws(self):
        b = self.num_envs
        return {"u": np.zeros(b), "reset_agent": np.full(b, -1, np.int32), "reset_goal": np.full(b, -1, np.int32)}

    def _spawn(self, mask):
        b = int(mask.sum())
        if self.fixed_goal is not None:
            self.goal[mask] = self.fixed_goal
        else:
            cells = self.rng.choice(self.goal_cells, b, where=mask, kind="reset_goal")
            self.goal[mask] = np.stack(np.unravel_index(cells, self.grid.shape), -1)
            self.draws["reset_goal"][mask] = cells
        cells = self.rng.choice(self.agent_cells, b, where=mask, kind="reset_agent")
        self.agent[mask] = np.stack(np.unravel_index(cells, self.grid.shape), -1)
        self.draws["reset_agent"][mask] = cells

    @property
    def state(self):
        return {"agent": self.agent.copy(), "goal": self.goal.copy(), "elapsed": self.elapsed.copy()}

    def set_state(self, agent, goal, elapsed):
        self.agent = np.array(agent, dtype=np.int64)
        self.goal = np.array(goal, dtype=np.int64)
        self.elapsed = np.array(elapsed, dtype=np.int64)

    def reset(self, *, seed=None, options=None):
        """msrooms.py:371-383 — returns (obs, {})"""
        if seed is not None:
            self.rng.reseed(seed)
        b = self.num_envs
        self.draws = self._blank_draws()
        self.elapsed = np.zeros(b, dtype=np.int64)
        self.goal = np.zeros((b, 3), dtype=np.int64)
        self.agent = np.zeros((b, 3), dtype=np.int64)
        self._spawn(np.ones(b, dtype=bool))
        return self._obs_fn(self.agent, self.goal), {}

    def step(self, action):
        """msrooms.py:392-413"""
        action = np.asarray(action)
        self.draws = self._blank_draws()
        self.elapsed += 1
        u = self.rng.random(self.num_envs, kind="slip", action=action, cumsum=self.P.cumsum(axis=1))
        self.draws["u"][:] = u
        target = self.agent + self.dirs[slip_sample(self.P[action], u)]
        blocked = self.grid[target[:, 0], target[:, 1], target[:, 2]] == WALL
        self.agent[~blocked] = target[~blocked]
        # stairs teleport, only for agents that moved (:419-428); both masks taken before either moves
        here = self.grid[self.agent[:, 0], self.agent[:, 1], self.agent[:, 2]]
        up = (here == STAIR_UP) & ~blocked
        down = (here == STAIR_DOWN) & ~blocked
        self.agent[up, 0] += 1
        self.agent[up, 1:] = DOWN_YX
        self.agent[down, 0] -= 1
        self.agent[down, 1:] = UP_YX
        at_goal = (self.agent == self.goal).all(-1)
        rew = np.zeros(self.num_envs, dtype=np.float32)
        rew += self.step_reward
        rew[blocked] = self.wall_reward
        rew[at_goal] = self.goal_reward
        truncated = self.elapsed > self.time_limit
        again = at_goal | truncated
        if again.any():
            self.elapsed[again] = 0
            self._spawn(again)
        return self._obs_fn(self.agent, self.goal), rew, at_goal, truncated, {}

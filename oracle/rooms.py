"""TEST INFRASTRUCTURE — numpy oracle for the vectorized discrete ROOMS / FourRooms step.

Restates the reference's ``RoomsEnv`` (gym_po/envs/rooms/rooms.py: ctor :84-175,
``reset`` :177-189, ``_reset_some`` :191-196, ``step`` :198-222, ``_out_of_bounds``
:224-226, obs dispatch :15-68), the observation functions
(rooms/observations.py :16-131), the slip sampler (rooms/action_utils.py :38-48,
:73-90) and the layout -> integer-grid conversion (rooms/layouts.py :217-232).
The map *data* is read from the text asset shared with the product package.
"""
from __future__ import annotations

import os

import numpy as np

from .draws import GeneratorDraws

_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gym-po-taxi_b200", "gym_po",
                      "envs", "rooms", "layouts.txt")

# compass tables (rooms/action_utils.py:16-29): N NE E SE S SW W NW; cardinal = every second one
DIRS8 = np.array([[-1, 0], [-1, 1], [0, 1], [1, 1], [1, 0], [1, -1], [0, -1], [-1, -1]])
DIRS4 = DIRS8[::2]


def _parse_asset():
    maps, ends, starts = {}, {}, {}
    name = None
    with open(_ASSET) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            if line.startswith("@"):
                parts = line[1:].split()
                name = parts[0]
                maps[name] = []
                for kv in parts[1:]:
                    k, v = kv.split("=")
                    xy = tuple(int(t) for t in v.split(","))
                    (ends if k == "end" else starts)[name] = xy
            else:
                maps[name].append(line)
    return maps, ends, starts


_MAPS, ENDS, STARTS = _parse_asset()
LAYOUT_NAMES = tuple(_MAPS)


def load_layout(name):
    """layout name -> int grid, -1 wall, k>=0 room id by rank of the room character
    (rooms/layouts.py:217-232)."""
    chars = np.array([list(r) for r in _MAPS[name]])
    rooms = sorted(set(chars.ravel().tolist()) - {"x"})
    grid = np.full(chars.shape, -1, dtype=np.int64)
    for k, ch in enumerate(rooms):
        grid[chars == ch] = k
    return grid


def slip_matrix(n, p_fail):
    """rooms/action_utils.py:38-48"""
    m = np.full((n, n), p_fail / (n - 1), dtype=np.float64)
    np.fill_diagonal(m, 1 - p_fail)
    return m


def slip_sample(rows, u):
    """a' = #{j : cumsum(P[a])_j < u}   (rooms/action_utils.py:84-90)"""
    return (rows.cumsum(axis=1) < u[:, None]).sum(axis=1)


# ---- observation functions (rooms/observations.py) ----------------------------
def dense_state_table(grid):
    """observations.py:16-29"""
    free = grid >= 0
    return int(free.sum()), (free.cumsum() - 1).reshape(grid.shape)


def count_rooms(grid):
    """observations.py:32-41"""
    return len(np.unique(grid)) - 1


def hansen_scalar(agent, grid, goal, n):
    """observations.py:44-71 — bit i = neighbour i is EMPTY, times (index of the neighbour
    holding the goal)+1; float64 like the reference."""
    dirs = DIRS4 if n == 4 else DIRS8
    nb = agent[:, None, :] + dirs[None]
    env_i, dir_i = np.nonzero((goal[:, None, :] == nb).all(-1))
    mult = np.ones(agent.shape[0])
    mult[env_i] = dir_i + 1
    empty = (grid[nb[..., 0], nb[..., 1]] >= 0).astype(np.int64)
    return empty.dot(2 ** np.arange(n)) * mult


def hansen_vector(agent, grid, goal, n):
    """observations.py:106-131 — 0 wall / 1 empty / 2 goal (goal only when given)"""
    dirs = DIRS4 if n == 4 else DIRS8
    nb = agent[:, None, :] + dirs[None]
    out = (grid[nb[..., 0], nb[..., 1]] >= 0).astype(np.int64)
    if goal is not None:
        out[(goal[:, None, :] == nb).all(-1)] = 2
    return out


def window(agent, grid, goal, n):
    """observations.py:74-103 — n x n egocentric crop, outside-the-map cells read grid[0,0]"""
    off = n // 2
    oy, ox = np.mgrid[:n, :n] - off
    y = agent[:, 0, None, None] + oy[None]
    x = agent[:, 1, None, None] + ox[None]
    outside = (y < 0) | (x < 0) | (y >= grid.shape[0]) | (x >= grid.shape[1])
    y = np.where(outside, 0, y)
    x = np.where(outside, 0, x)
    out = (grid[y, x] >= 0).astype(np.int64)
    out[(goal[:, 0, None, None] == y) & (goal[:, 1, None, None] == x)] = 2
    return out


def make_obs_fn(obs_type, grid, obs_n, to_cell=lambda a: a):
    """obs dispatch by substring, reference order room -> mdp -> hansen -> grid
    (rooms/rooms.py:15-68; crooms.py:16-88 passes floor(pos/cell) through ``to_cell``).
    Returns (single-obs shape, fn(agent, goal))."""
    vec, has_goal = "vector" in obs_type, "goal" in obs_type
    if "room" in obs_type:
        n = count_rooms(grid)
        if has_goal:
            return (), lambda a, g: grid[tuple(to_cell(a).T)] + n * grid[tuple(to_cell(g).T)]
        return (), lambda a, g: grid[tuple(to_cell(a).T)]
    if "mdp" in obs_type:
        if vec:  # raw positions (not converted to cells, also in the continuous env)
            if has_goal:
                return (4,), lambda a, g: np.concatenate((a, g), -1)
            return (2,), lambda a, g: a
        n, table = dense_state_table(grid)
        if has_goal:
            return (), lambda a, g: table[tuple(to_cell(a).T)] + n * table[tuple(to_cell(g).T)]
        return (), lambda a, g: table[tuple(to_cell(a).T)]
    if "hansen" in obs_type:
        k = 8 if "8" in obs_type else 4
        if vec:
            if has_goal:
                return (k,), lambda a, g: hansen_vector(to_cell(a), grid, to_cell(g), k)
            return (k,), lambda a, g: hansen_vector(to_cell(a), grid, None, k)
        return (), lambda a, g: hansen_scalar(to_cell(a), grid, to_cell(g), k)
    if "grid" in obs_type:
        return (obs_n, obs_n), lambda a, g: window(to_cell(a), grid, to_cell(g), obs_n)
    raise NotImplementedError("Observation type not recognized")


def resolve_goal(grid, layout, goal_xy):
    """fixed goal (x,y) -> (y,x); walls fall back to ENDS[layout] (rooms/rooms.py:153-158)"""
    gy, gx = goal_xy[1], goal_xy[0]
    if grid[gy, gx] < 0:
        ex, ey = ENDS[layout[:-1] if "b" in layout else layout]
        gy, gx = ey, ex
    return np.array([gy, gx])


class RoomsOracle:
    def __init__(self, num_envs, layout="4", time_limit=500, obs_type="mdp", obs_n=3,
                 action_failure_probability=0.2, action_type="ordinal", agent_xy=None, goal_xy=(0, 0),
                 step_reward=0.0, wall_reward=0.0, goal_reward=1.0, draws=None, **_):
        assert layout in _MAPS
        if agent_xy is not None:
            raise ValueError("agent_xy raises in the reference too (rooms/rooms.py:164-166)")
        self.num_envs = int(num_envs)
        self.grid = load_layout(layout)
        self.obs_shape, self._obs_fn = make_obs_fn(obs_type, self.grid, obs_n)
        self.valid_cells = np.flatnonzero(self.grid >= 0)
        self.dirs = DIRS4 if action_type == "cardinal" else DIRS8
        self.n_actions = len(self.dirs)
        self.P = slip_matrix(self.n_actions, action_failure_probability)
        self.time_limit = time_limit
        self.step_reward, self.wall_reward, self.goal_reward = step_reward, wall_reward, goal_reward
        self.fixed_goal = None if goal_xy is None else resolve_goal(self.grid, layout, goal_xy)
        self.rng = draws if draws is not None else GeneratorDraws()
        self.draws = {}

    def _blank_draws(self):
        b = self.num_envs
        return {"u": np.zeros(b, np.float64), "reset_agent": np.full(b, -1, np.int32),
                "reset_goal": np.full(b, -1, np.int32)}

    def _spawn(self, mask):
        """goal first, then agent (rooms/rooms.py:195-196)"""
        b = int(mask.sum())
        if self.fixed_goal is not None:
            self.goal[mask] = self.fixed_goal
        else:
            cells = self.rng.choice(self.valid_cells, b, where=mask, kind="reset_goal")
            self.goal[mask] = np.stack(np.unravel_index(cells, self.grid.shape), -1)
            self.draws["reset_goal"][mask] = cells
        cells = self.rng.choice(self.valid_cells, b, where=mask, kind="reset_agent")
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
        """rooms/rooms.py:177-189 — returns obs only"""
        if seed is not None:
            self.rng.reseed(seed)
        b = self.num_envs
        self.draws = self._blank_draws()
        self.elapsed = np.zeros(b, dtype=np.int64)
        self.goal = np.zeros((b, 2), dtype=np.int64)
        self.agent = np.zeros((b, 2), dtype=np.int64)
        self._spawn(np.ones(b, dtype=bool))
        return self._obs_fn(self.agent, self.goal)

    def step(self, action):
        """rooms/rooms.py:198-222"""
        action = np.asarray(action)
        self.draws = self._blank_draws()
        self.elapsed += 1
        u = self.rng.random(self.num_envs, kind="slip", action=action, cumsum=self.P.cumsum(axis=1))
        self.draws["u"][:] = u
        actual = slip_sample(self.P[action], u)
        target = self.agent + self.dirs[actual]
        blocked = self.grid[target[:, 0], target[:, 1]] == -1
        self.agent[~blocked] = target[~blocked]
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

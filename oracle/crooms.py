"""TEST INFRASTRUCTURE — numpy oracle for the continuous-position ROOMS step.

Restates the reference's ``CRoomsEnv`` (gym_po/envs/rooms/crooms.py: ctor :104-244 with
the ``sample_action`` closures :175-198, ``reset`` :251-266, ``_reset_some`` :268-274,
``step`` :276-298, ``_apply_action`` :300-331, ``_out_of_bounds`` :333-338) and
``grid_to_coord`` / ``coord_to_grid`` (rooms/utils.py:7-20).  All arithmetic is float64
like the reference.
"""
from __future__ import annotations

import numpy as np

from .draws import GeneratorDraws
from .rooms import DIRS4, DIRS8, load_layout, make_obs_fn, resolve_goal, slip_matrix, slip_sample, _MAPS

MAX_VELOCITY = 5.0  # crooms.py:169


def cell_centre(cell_yx, cell_size=1.0):
    """rooms/utils.py:7-12"""
    return cell_yx * cell_size + cell_size / 2


def cell_of(pos_yx, cell_size=1.0):
    """rooms/utils.py:15-20"""
    return np.floor(pos_yx / cell_size).astype(int)


class CRoomsOracle:
    def __init__(self, num_envs, layout="4", time_limit=500, use_velocity=False, cell_size=1.0,
                 obs_type="mdp", obs_m=3, action_failure_probability=0.2, action_type="yx",
                 action_std=0.2, action_power=1.0, agent_xy=None, goal_xy=(0, 0), step_reward=0.0,
                 wall_reward=0.0, goal_reward=1.0, goal_threshold=0.5, draws=None, **_):
        assert layout in _MAPS
        if agent_xy is not None:
            raise ValueError("agent_xy raises in the reference too (rooms/crooms.py:232-235)")
        self.num_envs = int(num_envs)
        self.grid = load_layout(layout)
        self.shape_yx = np.array(self.grid.shape)
        self.cell_size = cell_size
        self.obs_shape, self._obs_fn = make_obs_fn(obs_type, self.grid, obs_m,
                                                   to_cell=lambda p: cell_of(p, cell_size))
        self.valid_cells = np.flatnonzero(self.grid >= 0)
        self.continuous_actions = action_type == "yx"
        if not self.continuous_actions:
            self.dirs = DIRS4 if action_type == "cardinal" else DIRS8
            self.P = slip_matrix(len(self.dirs), action_failure_probability)
        self.action_std, self.action_power = action_std, action_power
        self.use_velocity = use_velocity
        self.time_limit = time_limit
        self.step_reward, self.wall_reward, self.goal_reward = step_reward, wall_reward, goal_reward
        self.goal_threshold = goal_threshold
        # fixed goal at a cell centre; note the goal sampler ignores cell_size (crooms.py:222-230)
        self.fixed_goal = None if goal_xy is None else cell_centre(resolve_goal(self.grid, layout, goal_xy))
        self.rng = draws if draws is not None else GeneratorDraws()
        self.draws = {}

    def _blank_draws(self):
        b = self.num_envs
        return {"u": np.zeros(b, np.float64), "noise": np.zeros((b, 2), np.float64),
                "resample": np.zeros((b, 2), np.float64), "reset_agent": np.full(b, -1, np.int32),
                "reset_goal": np.full(b, -1, np.int32)}

    def _spawn(self, mask):
        """goal, then agent, both at unit-cell centres; velocity zeroed (crooms.py:268-274)"""
        b = int(mask.sum())
        if self.fixed_goal is not None:
            self.goal[mask] = self.fixed_goal
        else:
            cells = self.rng.choice(self.valid_cells, b)
            self.goal[mask] = cell_centre(np.stack(np.unravel_index(cells, self.grid.shape), -1))
            self.draws["reset_goal"][mask] = cells
        cells = self.rng.choice(self.valid_cells, b)
        self.agent[mask] = cell_centre(np.stack(np.unravel_index(cells, self.grid.shape), -1))
        self.draws["reset_agent"][mask] = cells
        self.velocity[mask] = 0.0

    @property
    def state(self):
        return {"agent": self.agent.copy(), "goal": self.goal.copy(), "velocity": self.velocity.copy(),
                "elapsed": self.elapsed.copy()}

    def set_state(self, agent, goal, velocity, elapsed):
        self.agent = np.array(agent, dtype=np.float64)
        self.goal = np.array(goal, dtype=np.float64)
        self.velocity = np.array(velocity, dtype=np.float64)
        self.elapsed = np.array(elapsed, dtype=np.int64)

    def reset(self, *, seed=None, options=None):
        """crooms.py:251-266 — returns obs only"""
        if seed is not None:
            self.rng.reseed(seed)
        b = self.num_envs
        self.draws = self._blank_draws()
        self.elapsed = np.zeros(b, dtype=np.int64)
        self.goal = np.zeros((b, 2))
        self.agent = np.zeros((b, 2))
        self.velocity = np.zeros((b, 2))
        self._spawn(np.ones(b, dtype=bool))
        return self._obs_fn(self.agent, self.goal)

    def step(self, action):
        """crooms.py:276-298"""
        action = np.asarray(action)
        b = self.num_envs
        self.draws = self._blank_draws()
        self.elapsed += 1

        # noisy action (crooms.py:175-178 / :188-196)
        if self.continuous_actions:
            noise = self.rng.normal(self.action_std, action.shape)
            self.draws["noise"][:] = noise
            push = action + noise
        else:
            u = self.rng.random(b)
            self.draws["u"][:] = u
            push = self.dirs[slip_sample(self.P[action], u)]
            if self.action_std:
                noise = self.rng.normal(self.action_std, push.shape)
                self.draws["noise"][:] = noise
                push = push + noise
        push = push * self.action_power

        # _apply_action (crooms.py:300-331)
        if self.use_velocity:
            self.velocity += push
            np.clip(self.velocity, -MAX_VELOCITY, MAX_VELOCITY, out=self.velocity)
            target = self.agent + self.velocity
        else:
            target = self.agent + push
        target = target.clip(0, self.shape_yx - 1 - 1e-6)
        tc = cell_of(target, self.cell_size)
        blocked = self.grid[tc[:, 0], tc[:, 1]] == -1
        self.agent[~blocked] = target[~blocked]
        if blocked.any():
            centre = cell_centre(cell_of(self.agent[blocked], self.cell_size), self.cell_size)
            jitter = self.rng.normal(0.5, centre.shape)
            self.draws["resample"][blocked] = jitter
            self.agent[blocked] = np.clip(centre + jitter, centre - self.cell_size / 2,
                                          centre + self.cell_size / 2 - 1e-8)
            self.velocity[blocked] = 0.0

        # reward / done (crooms.py:290-297)
        rew = np.zeros(b, dtype=np.float32)
        at_goal = np.linalg.norm(self.agent - self.goal, 2, -1) <= self.goal_threshold
        rew += self.step_reward
        rew[blocked] = self.wall_reward
        rew[at_goal] = self.goal_reward
        truncated = self.elapsed > self.time_limit
        again = at_goal | truncated
        if again.any():
            self.elapsed[again] = 0
            self._spawn(again)
        return self._obs_fn(self.agent, self.goal), rew, at_goal, truncated, {}

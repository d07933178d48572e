"""TEST INFRASTRUCTURE — numpy oracle for the vectorized Taxi POMDP step.

Restates ``TaxiVecEnv`` of the reference (gym_po/envs/extended_taxi.py):
constructor tables :158-230, ``reset`` :232-242, ``step`` :244-287,
``_reset_mask`` :344-352, ``_reset_passenger_and_destination`` :354-364,
``_obs``/``_hansen_obs`` :366-372, helpers :57-118.  Movement is evaluated on the
bordered character map exactly as the reference does (target cell is ``'|'`` or a
``'|'`` separator is crossed) — deliberately *not* through the per-cell wall-bit
table the CUDA kernel uses, so the two derivations check each other.
"""
from __future__ import annotations

import numpy as np

from .draws import GeneratorDraws

# map data (extended_taxi.py:26-32, :45-54)
TAXI_MAP = ("R: | : :G", " : | : : ", " : : : : ", " | : | : ", "Y| : |B: ")
EXTENDED_TAXI_MAP = ("R  |   G", "   |    ", "   |    ", "        ", "        ",
                     "  |  |  ", "  |  |  ", "Y |  |B ")

_MOVE_DY = np.array([-1, 1, 0, 0, 0])   # N S W E pickup/dropoff (extended_taxi.py:154)
_MOVE_DX = np.array([0, 0, -1, 1, 0])
PICK_DROP = 4


class TaxiOracle:
    def __init__(self, num_envs=1, time_limit=200, num_passengers=1, map=TAXI_MAP, hansen_obs=False,
                 reward_goal=1.0, reward_bad=-0.5, reward_any=-0.05, draws=None):
        self.num_envs = int(num_envs)
        self.time_limit = time_limit
        self.n_dropoffs = num_passengers
        self.hansen = bool(hansen_obs)
        self.rew_goal, self.rew_bad, self.rew_any = reward_goal, reward_bad, reward_any
        self.rng = draws if draws is not None else GeneratorDraws()

        # bordered character map and the navigable sub-grid (extended_taxi.py:57-70)
        chars = np.array([list(row) for row in map])
        self.desc = np.pad(chars, 1, constant_values="|")
        self.pseudo = bool((self.desc == ":").any())
        self.tgrid = self.desc[1:-1, 1:-1:2] if self.pseudo else self.desc[1:-1, 1:-1]
        self.rows, self.cols = self.tgrid.shape
        self.is_wall = self.desc == "|"

        # named locations, row-major (extended_taxi.py:117-118, :182-185); sentinel row = "aboard"
        ly, lx = np.nonzero((self.tgrid != "|") & (self.tgrid != " ") & (self.tgrid != ":"))
        self.nlocs = len(ly)
        self.loc_r = np.concatenate((ly, [-1]))
        self.loc_c = np.concatenate((lx, [-1]))

        # wall bits N=1 S=2 W=4 E=8 around every cell (extended_taxi.py:102-114)
        self.hansen_bits = np.zeros((self.rows, self.cols), dtype=np.int64)
        for r in range(self.rows):
            for c in range(self.cols):
                br, bc = self._bordered(r, c)
                w = self.is_wall
                self.hansen_bits[r, c] = (int(w[br - 1, bc]) + 2 * int(w[br + 1, bc])
                                          + 4 * int(w[br, bc - 1]) + 8 * int(w[br, bc + 1]))

        # sizes (extended_taxi.py:73-81, :198-201)
        self.ns = self.rows * self.cols * (self.nlocs + 1) * self.nlocs
        self.no = (16 if self.hansen else self.rows * self.cols) * (self.nlocs + 1) * self.nlocs
        self.na = 5

        # reset distribution: uniform over valid states (extended_taxi.py:205-218)
        valid = [self.encode(r, c, p, d)
                 for r in range(self.rows) for c in range(self.cols) if self.tgrid[r, c] != "|"
                 for p in range(self.nlocs) for d in range(self.nlocs) if d != p]
        self.valid_states = np.array(valid)
        self.state_distribution = np.zeros(self.ns)
        self.state_distribution[self.valid_states] += 1
        self.state_distribution /= self.state_distribution.sum()

        # state (extended_taxi.py:189, :226-229); step() is legal before reset()
        self.s = np.zeros(self.num_envs, dtype=np.int64)
        self.elapsed = np.zeros(self.num_envs, dtype=np.int64)
        self.ndrop = np.zeros(self.num_envs, dtype=np.int64)
        self.draws = self._blank_draws()

    # ---- helpers -------------------------------------------------------
    def _bordered(self, r, c):
        """navigable (r,c) -> coordinates in the bordered map (extended_taxi.py:68/:70)"""
        return (r + 1, 2 * c + 1) if self.pseudo else (r + 1, c + 1)

    def encode(self, r, c, p, d):
        """extended_taxi.py:97-99"""
        return ((r * self.cols + c) * (self.nlocs + 1) + p) * self.nlocs + d

    def decode(self, s):
        """extended_taxi.py:84-94"""
        d = s % self.nlocs
        t = s // self.nlocs
        p = t % (self.nlocs + 1)
        t = t // (self.nlocs + 1)
        return t // self.cols, t % self.cols, p, d

    def _blank_draws(self):
        b = self.num_envs
        return {"reset_state": np.full(b, -1, np.int32), "new_p": np.full(b, -1, np.int8),
                "new_d": np.full(b, -1, np.int8)}

    @property
    def state(self):
        return {"s": self.s.copy(), "elapsed": self.elapsed.copy(), "ndrop": self.ndrop.copy()}

    def set_state(self, s, elapsed, ndrop):
        self.s = np.array(s, dtype=np.int64)
        self.elapsed = np.array(elapsed, dtype=np.int64)
        self.ndrop = np.array(ndrop, dtype=np.int64)

    # ---- API -----------------------------------------------------------
    def reset(self, *, seed=None, options=None):
        """extended_taxi.py:232-242"""
        if seed is not None:
            self.rng.reseed(seed)
        self.draws = self._blank_draws()
        self._full_reset(np.ones(self.num_envs, dtype=bool))
        return self._obs(), {}

    def step(self, actions):
        """extended_taxi.py:244-287"""
        actions = np.asarray(actions)
        self.draws = self._blank_draws()
        self.elapsed += 1
        r, c, p, d = self.decode(self.s)

        # move unless the target is a wall or a '|' separator is crossed (:248-260)
        dy, dx = _MOVE_DY[actions], _MOVE_DX[actions]
        r2 = np.clip(r + dy, 0, self.rows - 1)
        c2 = np.clip(c + dx, 0, self.cols - 1)
        br, bc = self._bordered(r2, c2)
        free = ~self.is_wall[br, bc]
        crossed = (dx != 0) & self.is_wall[br, bc - dx]
        go = free & ~crossed
        r = np.where(go, r2, r)
        c = np.where(go, c2, c)

        # pickup / dropoff (:262-275)
        act = actions == PICK_DROP
        at_dest = (self.loc_r[d] == r) & (self.loc_c[d] == c)
        goal_move = act & (p == self.nlocs) & at_dest
        self.ndrop[goal_move] += 1
        at_pass = (self.loc_r[p] == r) & (self.loc_c[p] == c)
        pickup = act & (p < self.nlocs) & at_pass
        p = np.where(pickup, self.nlocs, p)
        self.s = self.encode(r, c, p, d)
        bad = act & ~goal_move & ~pickup
        rew = np.full(self.num_envs, self.rew_any, dtype=np.float32)
        rew[goal_move] = self.rew_goal
        rew[bad] = self.rew_bad

        # termination / truncation (:276-279)
        terminated = self.ndrop == self.n_dropoffs
        truncated = self.elapsed > self.time_limit

        # respawn passenger+destination after an intermediate delivery (:283-285, :354-364)
        respawn = goal_move & ~(terminated | truncated)
        b = int(respawn.sum())
        if b:
            new_p = self.rng.integers(self.nlocs, size=b, where=respawn, kind="new_p")
            new_d = self.rng.integers(self.nlocs, size=b, where=respawn, kind="new_d")
            while True:
                clash = new_p == new_d
                if not clash.any():
                    break
                new_d[clash] = self.rng.integers(self.nlocs, size=int(clash.sum()), kind="redraw_d")
            self.s[respawn] = self.encode(r[respawn], c[respawn], new_p, new_d)
            self.draws["new_p"][respawn] = new_p
            self.draws["new_d"][respawn] = new_d

        self._full_reset(terminated | truncated)
        return self._obs(), rew, terminated, truncated, {}

    def _full_reset(self, mask):
        """extended_taxi.py:344-352"""
        b = int(mask.sum())
        if b:
            fresh = self.rng.multinomial_argmax(self.ns, self.state_distribution, b, where=mask, kind="reset_state")
            self.s[mask] = fresh
            self.elapsed[mask] = 0
            self.ndrop[mask] = 0
            self.draws["reset_state"][mask] = fresh

    def _obs(self):
        """extended_taxi.py:366-372"""
        if not self.hansen:
            return self.s
        r, c, p, d = self.decode(self.s)
        return (self.hansen_bits[r, c] * (self.nlocs + 1) + p) * self.nlocs + d

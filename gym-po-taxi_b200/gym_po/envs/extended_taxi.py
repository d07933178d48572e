"""Vectorized Taxi POMDP on B200 — host side.

Drop-in for the reference's ``gym_po.envs.extended_taxi`` (same class / alias names, constructor
kwargs, ``reset()`` -> ``(obs, {})`` and ``step(actions)`` -> 5-tuple with same-step autoreset;
reference gym_po/envs/extended_taxi.py:149-377).  This module only parses the map into small tables;
the whole step runs in one fused CUDA kernel (csrc/gpt_taxi.cu) behind the C ABI.

Differences a caller can observe (also listed in INTEGRATION.md):
* inputs/outputs are CUDA ``torch`` tensors: obs int32, reward float32, terminated/truncated bool;
  they are views of buffers the next ``step`` overwrites;
* random numbers come from Philox4x32-7 (``rng_mode='philox'``) instead of numpy's PCG64.  The
  *laws* are the reference's (including its argmax-of-multinomial reset distribution); streams
  differ.  ``rng_mode='replay'`` consumes pre-drawn values for bit-exact comparison.
"""
from __future__ import annotations

import ctypes as C
from functools import partial
from math import lgamma
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _native as N
from .._device_env import DeviceVecEnv
from ..spaces import Discrete, batch_space

__all__ = ["TaxiVecEnv", "ExtendedTaxiVecEnv", "HansenTaxiVecEnv", "ExtendedHansenTaxiVecEnv", "EXTENDED_TAXI_MAP",
           "TAXI_MAP"]

# Map data (reference extended_taxi.py:26-32 and :45-54).  ':' is a passable separator column,
# '|' a wall, letters are the pickup/dropoff locations.
TAXI_MAP = ("R: | : :G", " : | : : ", " : : : : ", " | : | : ", "Y| : |B: ")
EXTENDED_TAXI_MAP = ("R  |   G", "   |    ", "   |    ", "        ", "        ", "  |  |  ", "  |  |  ", "Y |  |B ")

WALL, PSEUDO, FLOOR = "|", ":", " "
EXACT_LAW_MAX_WORK = 4_000_000   # ns * n_valid above which Philox mode falls back to a uniform reset law


def parse_taxi_map(rows: Sequence[str]):
    """Map strings -> (n_rows, n_cols, wall_bits[cells], loc_cells[nlocs], is_wall_cell[cells]).

    ``wall_bits`` bit0 N, bit1 S, bit2 W, bit3 E is 1 when a '|' sits on that side of the cell in the
    '|'-bordered map — the reference's ``hansen_encodings`` (extended_taxi.py:102-114).  The same bit
    says whether a move in that direction is blocked (proved equivalent to :248-260 for every cell and
    direction by tests/test_host_tables.py against the oracle's character-map rule).
    """
    width = {len(r) for r in rows}
    if len(width) != 1:
        raise ValueError("all map rows must have the same length")
    sep = any(PSEUDO in r for r in rows)          # cells on even columns, separators on odd ones
    body = [WALL + r + WALL for r in rows]
    edge = WALL * len(body[0])
    full = [edge] + body + [edge]
    n_rows = len(rows)
    n_cols = (len(rows[0]) + 1) // 2 if sep else len(rows[0])
    col = (lambda c: 2 * c + 1) if sep else (lambda c: c + 1)
    wall_bits = np.zeros(n_rows * n_cols, dtype=np.uint8)
    is_wall = np.zeros(n_rows * n_cols, dtype=bool)
    locs = []
    for r in range(n_rows):
        for c in range(n_cols):
            y, x = r + 1, col(c)
            bits = ((full[y - 1][x] == WALL) | (full[y + 1][x] == WALL) << 1 | (full[y][x - 1] == WALL) << 2
                    | (full[y][x + 1] == WALL) << 3)
            wall_bits[r * n_cols + c] = bits
            ch = full[y][x]
            is_wall[r * n_cols + c] = ch == WALL
            if ch not in (WALL, PSEUDO, FLOOR):
                locs.append(r * n_cols + c)
    return n_rows, n_cols, wall_bits, np.array(locs, dtype=np.int32), is_wall


def argmax_multinomial_law(n_trials: int, n_cat: int) -> np.ndarray:
    """Exact law of ``multinomial(n_trials, uniform(n_cat)).argmax()`` (ties -> lowest index).

    The reference resets a finished env to ``np_random.multinomial(ns, state_distribution).argmax()``
    (extended_taxi.py:348-350), which is NOT uniform over the valid states: low ids win ties.  With
    iid Poisson(lam = n/N) counts conditioned on their sum (Poissonisation),

        P(max = m, tie size k) = C(N,k) q_m^k * P(N-k counts all < m and summing to n - k m) / P(Pois(n) = n)

    and, by exchangeability, the winner is the smallest index of a uniformly random k-subset:
    P(rank j | k) = C(N-1-j, k-1) / C(N, k).  Hence

        P(rank j) = sum_{m,k} C(N-1-j, k-1) q_m^k A_m^{(N-k)}[n - k m] / Z

    with A_m^{(r)} the r-fold convolution of the Poisson pmf truncated below m (degree m-1, so each
    convolution is O(n m)).  Everything is a sum of positive terms; float64 is ample.
    """
    n, N_ = int(n_trials), int(n_cat)
    lam = n / N_
    ks = np.arange(n + 1)
    logpmf = -lam + ks * np.log(lam) - np.array([lgamma(k + 1) for k in ks])
    pmf = np.exp(logpmf)
    log_z = -n + n * np.log(n) - lgamma(n + 1)            # log P(Pois(N*lam = n) = n)
    j = np.arange(N_)
    lg = np.array([lgamma(v + 1) for v in range(N_ + 1)])
    law = np.zeros(N_)
    total = 0.0
    m = max(1, -(-n // N_))                               # the maximum is at least ceil(n/N)
    while m <= n:
        trunc = pmf[:m]
        kmax = min(N_, n // m)
        # powers[r] = r-fold convolution restricted to sums <= n, for r = N-kmax .. N-1
        need_lo = N_ - kmax
        cur = np.zeros(n + 1)
        cur[0] = 1.0
        powers = {}
        for r in range(0, N_):
            if r >= need_lo:
                powers[r] = cur
            cur = np.convolve(cur, trunc)[: n + 1]
        mass_m = 0.0
        for k in range(1, kmax + 1):
            a = powers[N_ - k][n - k * m] if N_ - k >= 0 else 0.0
            if N_ - k == 0:
                a = 1.0 if n - k * m == 0 else 0.0
            if a <= 0.0:
                continue
            # log C(N-1-j, k-1) for the ranks where it is defined
            top = N_ - 1 - j
            ok = top >= k - 1
            logc = np.full(N_, -np.inf)
            logc[ok] = lg[top[ok]] - lg[k - 1] - lg[top[ok] - (k - 1)]
            term = np.exp(logc + k * logpmf[m] + np.log(a) - log_z)
            law += term
            mass_m += term.sum()
        total += mass_m
        if total > 1 - 1e-13 or (m > 4 * lam + 20 and mass_m < 1e-18):
            break
        m += 1
    if not abs(total - 1.0) < 1e-9:
        raise ArithmeticError(f"argmax-multinomial law does not sum to 1 (got {total})")
    return law / law.sum()


def law_to_cdf32(law: np.ndarray) -> np.ndarray:
    """Inclusive upper bounds u32: a uniform 32-bit draw r selects the first j with r <= cdf[j]."""
    cdf = np.floor(np.cumsum(law) * 2.0**32).astype(np.int64) - 1
    cdf = np.clip(cdf, 0, 2**32 - 1)
    cdf[-1] = 2**32 - 1
    return cdf.astype(np.uint32)


class TaxiVecEnv(DeviceVecEnv):
    """Vectorized Taxi environment (fused CUDA step)."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 5, "name": "Taxi"}
    ACTION_NAMES = ["North", "South", "West", "East", "Pickup/Dropoff"]
    ACTION_DICT = dict(enumerate(ACTION_NAMES))
    ACTIONS_YX = np.array([[-1, 0], [1, 0], [0, -1], [0, 1], [0, 0]], dtype=int)

    def __init__(self, num_envs: int = 1, time_limit: int = 200, num_passengers: int = 1,
                 map: Sequence[str] = TAXI_MAP, hansen_obs: bool = False, reward_goal: float = 1.0,
                 reward_bad: float = -0.5, reward_any: float = -0.05, render_mode: Optional[str] = None, *,
                 device=None, rng_mode: str = "philox", seed: Optional[int] = None, env_offset: int = 0,
                 track_stats: bool = False):
        self.render_mode = render_mode
        self.num_envs = int(num_envs)
        self.GOAL_MOVE, self.BAD_MOVE, self.ANY_MOVE = reward_goal, reward_bad, reward_any
        self.rows, self.cols, wall_bits, loc_cells, is_wall = parse_taxi_map(map)
        self._map_rows = tuple(map)
        self.hansen_encodings = wall_bits.reshape(self.rows, self.cols).astype(int)
        self.nlocs = len(loc_cells)
        self.np_locs = np.concatenate((np.stack(np.divmod(loc_cells, self.cols), -1), [[-1, -1]]))
        self.time_limit = time_limit
        self.n_dropoffs = num_passengers
        self.hansen = bool(hansen_obs)
        if self.hansen:
            self.name = "HansenTaxi-v4"

        self.single_action_space = Discrete(5)
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.na = 5
        cells = self.rows * self.cols
        self.ns = cells * (self.nlocs + 1) * self.nlocs
        # limits of the packed device tables (checked before any table is built)
        if self.nlocs < 2:
            raise ValueError("the map needs at least two pickup/dropoff locations")
        if self.ns >= 65536 or self.nlocs > 254:
            raise ValueError(f"state space too large for the device tables: rows*cols*(nlocs+1)*nlocs = {self.ns} >= 65536")
        if not 0 <= int(num_passengers) <= 255:
            raise ValueError("num_passengers must be in [0, 255]")
        self.no = (16 if self.hansen else cells) * (self.nlocs + 1) * self.nlocs
        self.single_observation_space = Discrete(self.no)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)

        # valid reset states: taxi on a non-wall cell, passenger waiting at a location != destination
        valid = [((cell * (self.nlocs + 1)) + p) * self.nlocs + d
                 for cell in range(cells) if not is_wall[cell]
                 for p in range(self.nlocs) for d in range(self.nlocs) if d != p]
        self.valid_states = np.array(valid, dtype=np.int32)
        self.state_distribution = np.zeros(self.ns)
        self.state_distribution[self.valid_states] = 1.0 / len(valid)
        cdf = None
        if rng_mode == "philox":
            if self.ns * len(valid) <= EXACT_LAW_MAX_WORK:
                self.reset_law = argmax_multinomial_law(self.ns, len(valid))
            else:  # documented deviation: the exact law costs O(ns * n_valid * max_count) on the host
                import warnings
                warnings.warn("custom map too large for the exact argmax-of-multinomial reset law; Philox mode resets "
                              "uniformly over the valid states (rng_mode='replay' stays exact)", RuntimeWarning)
                self.reset_law = np.full(len(valid), 1.0 / len(valid))
            cdf = law_to_cdf32(self.reset_law)

        cfg = N.GptConfig()
        cfg.family = N.FAMILY_TAXI
        cfg.time_limit = int(time_limit)
        cfg.taxi_rows, cfg.taxi_cols, cfg.taxi_nlocs = self.rows, self.cols, self.nlocs
        cfg.taxi_n_dropoffs = int(num_passengers)
        cfg.taxi_hansen_obs = int(self.hansen)
        cfg.taxi_reward_goal, cfg.taxi_reward_bad, cfg.taxi_reward_any = reward_goal, reward_bad, reward_any
        wall_c = np.ascontiguousarray(wall_bits, dtype=np.uint8)
        loc_c = np.ascontiguousarray(loc_cells, dtype=np.int32)
        cfg.taxi_wall_bits = wall_c.ctypes.data_as(C.POINTER(C.c_uint8))
        cfg.taxi_loc_cell = loc_c.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.taxi_n_valid = len(valid)
        cfg.taxi_valid_states = self.valid_states.ctypes.data_as(C.POINTER(C.c_int32))
        if cdf is not None:
            cfg.taxi_reset_cdf = cdf.ctypes.data_as(C.POINTER(C.c_uint32))
        self._create(cfg, device=device, rng_mode=rng_mode, seed=seed, env_offset=env_offset,
                     track_stats=track_stats, keepalive=(wall_c, loc_c, cdf))
        self.lastaction = None

    # ---- reference-compatible helpers / attributes ------------------------------------------
    def encode(self, r, c, p, d):
        return ((r * self.cols + c) * (self.nlocs + 1) + p) * self.nlocs + d

    def decode(self, s):
        d = s % self.nlocs
        t = s // self.nlocs
        p = t % (self.nlocs + 1)
        t = t // (self.nlocs + 1)
        return t // self.cols, t % self.cols, p, d

    @property
    def s(self) -> torch.Tensor:
        """Encoded state, int32 device tensor (live view, like the reference's ``self.s``)."""
        return self._arrays["s"][: self.num_envs]

    @property
    def elapsed(self) -> torch.Tensor:
        return self._arrays["elapsed"][: self.num_envs]

    @property
    def n_dropoffs_completed(self) -> torch.Tensor:
        return self._arrays["ndrop"][: self.num_envs]

    def get_state(self):
        return {"s": self.s.clone(), "elapsed": self.elapsed.clone(), "ndrop": self.n_dropoffs_completed.clone()}

    def set_state(self, s, elapsed, ndrop):
        b = self.num_envs
        s = np.asarray(s)
        if s.size and (s.min() < 0 or s.max() >= self.ns):   # the kernels index their tables with it unchecked
            raise ValueError(f"set_state: state ids must be in [0, {self.ns})")
        self._arrays["s"][:b].copy_(torch.as_tensor(np.asarray(s)).to(torch.int32))
        self._arrays["elapsed"][:b].copy_(torch.as_tensor(np.asarray(elapsed)).to(torch.int32))
        self._arrays["ndrop"][:b].copy_(torch.as_tensor(np.asarray(ndrop)).to(torch.uint8))

    # ---- gym API ----------------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """Fully reset all environments -> ``(obs, {})`` (reference :232-242)."""
        self.lastaction = None
        self._last_actions = None
        return self._reset(seed), {}

    def step(self, actions):
        out = super().step(actions)
        self._last_actions = actions   # env 0's action is shown by render(); looked up lazily (no sync here)
        return out

    def render(self, idx=None):
        """RGB frame of the selected envs (default: env 0), like the reference (:289-342): copies only their encoded
        states to the host and draws there (gym_po/envs/taxi_render.py).  ``render_mode='human'`` also blits to a
        pygame window when pygame is importable."""
        from .taxi_render import render_taxi
        idx = np.arange(1) if idx is None else np.atleast_1d(np.asarray(idx, dtype=np.int64))
        s = self.s[torch.as_tensor(idx, device=self.device)].cpu().numpy()
        name = None
        last = getattr(self, "_last_actions", None)
        if last is not None and not bool(self._terminated[0]):   # cleared on done[0] = terminated only (reference :280)
            name = self.ACTION_NAMES[int(torch.as_tensor(last).reshape(-1)[0])]
        img = render_taxi(self._map_rows, s, self.nlocs, self.np_locs, self.cols, self.hansen, name)
        if self.render_mode == "human":  # pragma: no cover - needs a display
            import pygame
            if getattr(self, "_viewer", None) is None:
                pygame.init()
                self._viewer = pygame.display.set_mode(img.shape[:-1])
            self._viewer.blit(pygame.surfarray.make_surface(img.swapaxes(0, 1)), (0, 0))
            pygame.display.update()
        return img


HansenTaxiVecEnv = partial(TaxiVecEnv, hansen_obs=True)
ExtendedTaxiVecEnv = partial(TaxiVecEnv, map=EXTENDED_TAXI_MAP)
ExtendedHansenTaxiVecEnv = partial(HansenTaxiVecEnv, map=EXTENDED_TAXI_MAP)

"""Vectorized car-flag ("heaven / hell with a priest") on B200 — host side.

Drop-in for the reference's ``CarVecEnv`` / ``DiscreteActionCarVecEnv`` (gym_po/envs/car_flag.py:23-144,
:286-303): same constructor arguments, ``reset()`` -> ``(obs, {})``, ``step(actions)`` -> 5-tuple with
same-step autoreset.  As in the reference the observation IS the live float32 state tensor ``s`` [B,3]
(position, velocity, priest indicator) — ``clone()`` it to keep a copy.  ``truncated = elapsed >=
time_limit`` (this env uses ``>=``, the others ``>``).  The step is one fused CUDA kernel (csrc/gpt_car.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from .. import _native as N
from .._device_env import DeviceVecEnv
from ..spaces import Box, Discrete, batch_space

__all__ = ["CarVecEnv", "DiscreteActionCarVecEnv"]


class CarVecEnv(DeviceVecEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 10}
    MAX_POS = 1.1
    MIN_POS = -MAX_POS
    POS_RANGE = MAX_POS - MIN_POS
    MAX_SPEED = 0.07
    MIN_ACT = -1.0
    MAX_ACT = 1.0
    PRIEST = 0.5
    PRIEST_THRESHOLD = 0.2
    POWER = 0.0015

    def __init__(self, num_envs: int, time_limit: int = 160, render_mode: Optional[str] = None, *, device=None,
                 rng_mode: str = "philox", seed: Optional[int] = None, env_offset: int = 0,
                 action_dtype=torch.float32, _num_actions: int = 0):
        self.num_envs = int(num_envs)
        self.single_observation_space = Box(np.array([self.MIN_POS, -self.MAX_SPEED, -1.0]),
                                            np.array([self.MAX_POS, self.MAX_SPEED, 1.0]), dtype=np.float32)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        self.single_action_space = Box(self.MIN_ACT, self.MAX_ACT, (1,), dtype=np.float32)
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.render_mode = render_mode
        self.time_limit = time_limit
        cfg = N.GptConfig()
        cfg.family = N.FAMILY_CAR
        cfg.time_limit = int(time_limit)
        cfg.c_action_f64 = int(action_dtype == torch.float64)
        keep = ()
        if _num_actions:
            self._actions = np.ascontiguousarray(np.linspace(self.MIN_ACT, self.MAX_ACT, _num_actions))
            cfg.car_num_actions = int(_num_actions)
            cfg.car_action_table = self._actions.ctypes.data_as(C.POINTER(C.c_double))
            keep = (self._actions,)
        self._create(cfg, device=device, rng_mode=rng_mode, seed=seed, env_offset=env_offset, keepalive=keep)
        # reference defaults before the first reset(): heavens = +1, priests = +0.5 (car_flag.py:78-80)
        self._arrays["flags"].fill_(3)

    def _device_actions(self, actions):
        t = actions
        if isinstance(t, torch.Tensor) and t.dim() == 2 and t.shape[1] == 1:
            t = t.reshape(-1)                       # the reference flattens [B,1] forces (car_flag.py:116)
        elif not isinstance(t, torch.Tensor):
            t = np.asarray(actions).reshape(-1)
        return super()._device_actions(t)

    # ---- state (reference attributes s / elapsed / heavens / hells / priests) ----
    @property
    def s(self) -> torch.Tensor:
        return self._obs

    @property
    def elapsed(self) -> torch.Tensor:
        return self._arrays["elapsed"][: self.num_envs]

    @property
    def heavens(self) -> torch.Tensor:
        return (self._arrays["flags"][: self.num_envs] & 1).float() * 2 - 1

    @property
    def hells(self) -> torch.Tensor:
        return -self.heavens

    @property
    def priests(self) -> torch.Tensor:
        return ((self._arrays["flags"][: self.num_envs] >> 1) & 1).double() - 0.5

    def get_state(self):
        return {"s": self.s.clone(), "elapsed": self.elapsed.clone(), "heavens": self.heavens, "priests": self.priests}

    def set_state(self, s, elapsed, heavens, priests):
        b = self.num_envs
        self._arrays["obs"][:b].copy_(torch.as_tensor(np.asarray(s, dtype=np.float32)))
        self._arrays["elapsed"][:b].copy_(torch.as_tensor(np.asarray(elapsed)).to(torch.int32))
        fl = (np.asarray(heavens) > 0).astype(np.uint8) | ((np.asarray(priests) > 0).astype(np.uint8) << 1)
        self._arrays["flags"][:b].copy_(torch.as_tensor(fl))

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        return self._reset(seed), {}

    def render(self):
        """Number-line frame of env 0 (reference car_flag.py:146-185): copies its 3 state floats + flag bits to
        the host and draws there (gym_po/envs/car_render.py)."""
        from .car_render import render_car
        s0 = self.s[0].cpu().numpy()
        img = render_car(float(s0[0]), float(s0[2]), float(self.heavens[0]), float(self.priests[0]))
        if self.render_mode != "rgb_array" and self.render_mode is not None:  # pragma: no cover - needs a display
            import pygame
            if getattr(self, "viewer", None) is None:
                pygame.init()
                self.viewer = pygame.display.set_mode(img.shape[:-1])
            self.viewer.blit(pygame.surfarray.make_surface(img), (0, 0))
            pygame.display.update()
        return img


class DiscreteActionCarVecEnv(CarVecEnv):
    """Discrete action car environment: evenly spaced forces along the control dimension."""

    def __init__(self, num_actions: int, *args, **kwargs):
        super().__init__(*args, _num_actions=int(num_actions), **kwargs)
        nact = num_actions // 2
        self.action_names = ["<" * i + ":" for i in reversed(range(1, nact + 1))] + [":" + ">" * i for i in range(1, nact + 1)]
        if num_actions % 2 == 1:
            self.action_names.insert(nact, ":")
        self.single_action_space = Discrete(num_actions)
        self.action_space = batch_space(self.single_action_space, self.num_envs)

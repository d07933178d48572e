"""Observation / action space descriptors.

The reference uses ``gymnasium.spaces`` and ``gymnasium.vector.utils.batch_space``
(extended_taxi.py:193-201, rooms/rooms.py:25-64, :141-143).  gymnasium is used when it is
importable; otherwise these minimal stand-ins expose the same attributes (``n``, ``shape``,
``low``/``high``, ``nvec``, ``dtype``, ``sample()``, ``contains()``).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not installed in the build image
    from gymnasium.spaces import Box, Discrete, MultiDiscrete  # type: ignore
    from gymnasium.vector.utils import batch_space  # type: ignore
except Exception:

    class _Space:
        def __init__(self, shape, dtype):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self._rng = None

        @property
        def np_random(self):
            if self._rng is None:
                self._rng = np.random.default_rng()
            return self._rng

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class Discrete(_Space):
        def __init__(self, n, start=0):
            super().__init__((), np.int64)
            self.n, self.start = int(n), int(start)

        def sample(self):
            return self.start + int(self.np_random.integers(self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(_Space):
        def __init__(self, nvec, dtype=np.int64):
            self.nvec = np.asarray(nvec, dtype=dtype)
            super().__init__(self.nvec.shape, dtype)

        def sample(self):
            return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x >= 0) & (x < self.nvec)).all())

        def __repr__(self):
            return f"MultiDiscrete(shape={self.shape})"

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()

        def sample(self):
            if self.dtype.kind == "f":
                return self.np_random.uniform(self.low, self.high).astype(self.dtype)
            return self.np_random.integers(self.low, self.high + 1).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())

        def __repr__(self):
            return f"Box(shape={self.shape}, dtype={self.dtype})"

    def batch_space(space, n=1):
        if isinstance(space, Discrete):
            return MultiDiscrete(np.full((n,), space.n, dtype=np.int64))
        if isinstance(space, Box):
            rep = (n,) + (1,) * space.low.ndim
            return Box(np.tile(space.low, rep), np.tile(space.high, rep), dtype=space.dtype)
        raise NotImplementedError(type(space))

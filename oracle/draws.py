"""TEST INFRASTRUCTURE — random-draw sources for the oracle envs.

The reference draws from ``self.np_random`` / ``self.rng`` (a numpy ``Generator``
over PCG64).  The call sites and their order are:

* Taxi full reset      ``multinomial(ns, state_distribution, b).argmax(-1)``   extended_taxi.py:348-350
* Taxi respawn         ``integers(nlocs, size=b)`` x2 + redraw loop           extended_taxi.py:360-363
* slip sampler         ``random(B)``                                          rooms/action_utils.py:84
* cell sampling        ``choice(valid_states, b)``                            rooms/rooms.py:160-162,170-172
* Gaussian noise       ``normal(scale=s, size=shape)``                        rooms/crooms.py:178,194,324
* tag target           ``integers(4)``                                        ant_tag.py:109
* car respawn          ``uniform(-0.2, 0.2, (b,1))``, ``choice([-1,1], b)`` x2  car_flag.py:100-110

Every call also carries keyword *context* (``where`` = bool mask of the envs the draw is for, ``kind`` = which
reference call site, ``action`` / ``cumsum`` for the slip).  ``GeneratorDraws`` and ``RecordedDraws`` ignore it — the
reference's generator knows nothing about env indices; a per-env counter-based source (tests/philox_host.py, which
rebuilds on the host the draws the Philox-mode CUDA kernels consume) needs it.
"""
from __future__ import annotations

import numpy as np


def make_generator(seed=None) -> np.random.Generator:
    """gymnasium>=0.26 ``seeding.np_random``: Generator(PCG64(SeedSequence(seed)))."""
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class GeneratorDraws:
    """Issues the reference's Generator calls, one for one, on ``gen``."""

    def __init__(self, gen: np.random.Generator | None = None, seed=None):
        self.gen = gen if gen is not None else make_generator(seed)

    def reseed(self, seed):
        self.gen = make_generator(seed)

    def multinomial_argmax(self, n, pvals, b, **ctx):
        # row chunks: Generator.multinomial fills rows sequentially from the bit stream, so chunked calls return the
        # same values as one call, without the reference's [b, ns] int64 temporary (16.8 GB at b = 2^22)
        step = max(1, (1 << 24) // max(len(pvals), 1))
        if b <= step:
            return self.gen.multinomial(n, pvals, b).argmax(-1)
        return np.concatenate([self.gen.multinomial(n, pvals, min(step, b - i)).argmax(-1) for i in range(0, b, step)])

    def integers(self, high, size=None, **ctx):
        return self.gen.integers(high, size=size)

    def random(self, b, **ctx):
        return self.gen.random(b)

    def choice(self, values, b, **ctx):
        return self.gen.choice(values, b)

    def normal(self, scale, size, **ctx):
        return self.gen.normal(scale=scale, size=size)

    def uniform(self, low, high, size, **ctx):
        return self.gen.uniform(low, high, size)


class RecordedDraws:
    """Hands back values recorded from the reference, checking the call kind.

    ``log`` is a list of ``(kind, array)`` with kind in {multinomial_argmax,
    integers, random, choice, normal}; produced by ``tests/golden/make_golden.py``.
    """

    def __init__(self, log):
        self.log = list(log)
        self.pos = 0

    def reseed(self, seed):
        pass

    def _next(self, kind, size=None):
        if self.pos >= len(self.log):
            raise RuntimeError(f"recorded draws exhausted at call {self.pos} ({kind})")
        k, v = self.log[self.pos]
        self.pos += 1
        if k != kind:
            raise RuntimeError(f"draw {self.pos - 1}: oracle asked for {kind}, reference drew {k}")
        v = np.asarray(v)
        if size is not None and int(np.prod(np.shape(v))) != int(np.prod(size)):
            raise RuntimeError(f"draw {self.pos - 1} ({kind}): size {np.shape(v)} != requested {size}")
        return v.copy()

    def multinomial_argmax(self, n, pvals, b, **ctx):
        return self._next("multinomial_argmax", (b,))

    def integers(self, high, size=None, **ctx):
        return self._next("integers", () if size is None else (size,))

    def random(self, b, **ctx):
        return self._next("random", (b,))

    def choice(self, values, b, **ctx):
        return self._next("choice", (b,))

    def normal(self, scale, size, **ctx):
        return self._next("normal", size).reshape(size)

    def uniform(self, low, high, size, **ctx):
        return self._next("uniform", size).reshape(size)

    @property
    def exhausted(self):
        return self.pos == len(self.log)

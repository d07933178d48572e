"""TEST INFRASTRUCTURE — loader for the *real* reference package.

Only used (a) by ``tests/golden/make_golden.py`` to generate the committed
golden fixtures and (b) by the ``not gpu`` tests that cross-check the oracle
port against the reference when ``/root/reference`` is present (this
container only; the directory does not exist on the GPU box).  Nothing in the
product package, ``bench.py`` or the ``gpu`` tests imports this module.

What it does:

* Shim 1 — registers in-memory stand-ins for the third-party modules the
  reference imports but this image lacks (``gymnasium``, ``pyglet``,
  ``dotsi``).  Only the symbols the reference touches exist.  The seeding
  helper is ``Generator(PCG64(SeedSequence(seed)))`` exactly like gymnasium
  >= 0.26 ``seeding.np_random``.
* Shim 2 — loads ``/root/reference/gym_po`` under the alias module name
  ``_gym_po_reference`` with a ``SourceFileLoader`` whose ``get_data`` applies
  the one documented in-memory repair: parameter names that a botched
  search-and-replace fused with their annotation (``actionsNDArray`` ->
  ``actions``; reference rooms/action_utils.py:52,61,74, rooms/msrooms.py:132
  ..., rooms/render_utils.py:29,39).  No reference file is modified or copied.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import re
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("GYM_PO_REFERENCE_ROOT", "/root/reference")
ALIAS = "_gym_po_reference"
_FUSED = re.compile(r"\b([A-Za-z_]\w*?)NDArray\b(?=\s*[,)])")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gym_po", "__init__.py"))


# --------------------------------------------------------------------------
# Shim 1: third-party stand-ins
# --------------------------------------------------------------------------
def _np_random(seed=None):
    ss = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(ss)), ss.entropy


class _Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = None

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = _np_random()
        return self._np_random


class _Discrete(_Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def sample(self):
        return self.start + int(self.np_random.integers(self.n))


class _MultiDiscrete(_Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
        super().__init__(shape, dtype, seed)

    def sample(self):
        u = self.np_random.uniform(self.low, self.high)
        return u.astype(self.dtype)


def _batch_space(space, n=1):
    if isinstance(space, _Discrete):
        return _MultiDiscrete(np.full((n,), space.n, dtype=np.int64))
    if isinstance(space, _Box):
        rep = (n,) + (1,) * space.low.ndim
        return _Box(np.tile(space.low, rep), np.tile(space.high, rep), dtype=space.dtype)
    raise NotImplementedError(type(space))


class _Env:
    metadata: dict = {"render_modes": []}
    render_mode = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = _np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, _ = _np_random(seed)

    def close(self):
        pass


class _EzPickle:
    def __init__(self, *a, **k):
        pass


def _install_stubs():
    if "gymnasium" in sys.modules and not getattr(sys.modules["gymnasium"], "_is_oracle_stub", False):
        return  # a real gymnasium is importable; use it
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m._is_oracle_stub = True
        sys.modules[name] = m
        return m

    spaces = mod("gymnasium.spaces", Space=_Space, Discrete=_Discrete, Box=_Box, MultiDiscrete=_MultiDiscrete)
    core = mod("gymnasium.core", ObsType=object, ActType=object, RenderFrame=object, Env=_Env)
    seeding = mod("gymnasium.utils.seeding", np_random=_np_random)
    utils = mod("gymnasium.utils", seeding=seeding, EzPickle=_EzPickle)
    vutils = mod("gymnasium.vector.utils", batch_space=_batch_space)
    vector = mod("gymnasium.vector", utils=vutils)
    registration = mod("gymnasium.envs.registration", register=lambda *a, **k: None)
    mujoco = mod("gymnasium.envs.mujoco", MujocoEnv=type("MujocoEnv", (_Env,), {}))
    envs = mod("gymnasium.envs", registration=registration, mujoco=mujoco)
    mod("gymnasium", Env=_Env, Space=_Space, spaces=spaces, core=core, utils=utils,
        vector=vector, envs=envs)
    if "pyglet" not in sys.modules:
        mod("pyglet", options={})
    if "dotsi" not in sys.modules:
        class DotsiDict(dict):
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__
        mod("dotsi", DotsiDict=DotsiDict, Dict=DotsiDict)


# --------------------------------------------------------------------------
# Shim 2: alias import with the in-memory signature repair
# --------------------------------------------------------------------------
class _RepairLoader(importlib.machinery.SourceFileLoader):
    def get_data(self, path):
        data = super().get_data(path)
        if path.endswith(".py"):
            data = _FUSED.sub(r"\1", data.decode("utf-8")).encode("utf-8")
        return data

    # never read or write .pyc files: the repair must always be applied
    def get_code(self, fullname):
        src = self.get_data(self.get_filename(fullname))
        return compile(src, self.get_filename(fullname), "exec", dont_inherit=True)


class _AliasFinder(importlib.abc.MetaPathFinder):
    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        if fullname != ALIAS and not fullname.startswith(ALIAS + "."):
            return None
        rel = fullname.split(".")[1:]
        base = os.path.join(self.root, "gym_po", *rel)
        if os.path.isdir(base):
            fn = os.path.join(base, "__init__.py")
            return importlib.util.spec_from_file_location(
                fullname, fn, loader=_RepairLoader(fullname, fn), submodule_search_locations=[base])
        fn = base + ".py"
        if os.path.isfile(fn):
            return importlib.util.spec_from_file_location(fullname, fn, loader=_RepairLoader(fullname, fn))
        return None


_loaded = None


def load_reference():
    """Import the reference as ``_gym_po_reference`` and return its ``envs`` sub-package."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder(REFERENCE_ROOT))
    _loaded = importlib.import_module(ALIAS + ".envs")
    return _loaded


# --------------------------------------------------------------------------
# Recording proxy for np.random.Generator
# --------------------------------------------------------------------------
class RecordingGenerator:
    """Wraps a ``Generator``; every call's *result* is appended to ``log`` as
    ``(method, result_copy)`` so a replay source can hand back the same values."""

    _METHODS = ("random", "integers", "choice", "multinomial", "normal", "uniform")

    def __init__(self, gen):
        self._gen = gen
        self.log = []

    def __getattr__(self, name):
        attr = getattr(self._gen, name)
        if name not in self._METHODS:
            return attr

        def wrapped(*a, **k):
            out = attr(*a, **k)
            self.log.append((name, np.array(out, copy=True)))
            return out

        return wrapped

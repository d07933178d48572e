"""Device-side wrapper layer.

The reference's author stacks gymnasium's ``RecordEpisodeStatistics`` and ``NormalizeReward`` on the vector envs
(gym_po/tester.py:36-41).  These classes keep that usage —

    env = NormalizeReward(RecordEpisodeStatistics(TaxiVecEnv(1 << 22)), gamma=0.95)
    obs, reward, terminated, truncated, info = env.step(actions)
    info["episode"]["r"], info["episode"]["l"], info["_episode"]      # device tensors

— but run as fused CUDA kernels (csrc/gpt_wrappers.cu) on the env's own output tensors, on the same stream as the
step: nothing is copied to the host.  Semantics follow gymnasium 0.29 (``returns = returns*gamma*(1-terminated) +
reward``; 0.27 / 0.28 instead zero the returns on terminated|truncated AFTER normalising) as restated in
oracle/wrappers.py — gymnasium itself is not available here, so this parity is unpinned; the normalised reward is
float32 (gymnasium returns float64).  ``stats()`` gives whole-job episode statistics, summed
over ranks with one NCCL all-reduce of an 8-double vector (the only collective on this path).

CUDA graphs: ``RecordEpisodeStatistics.step`` can be captured together with a graph-mode env (its launch has no
per-step host state); ``NormalizeReward`` cannot — its two launches alternate between the halves of a double buffer
chosen on the host.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N
from ._device_env import _CudaBuffer
from .sharding import allreduce_stats

__all__ = ["RecordEpisodeStatistics", "NormalizeReward"]


class _DeviceWrapper:
    """Forwards everything to the wrapped env; subclasses post-process ``step``."""

    _FLAGS = 0

    def __init__(self, env, gamma=0.99, epsilon=1e-8):
        self.env = env
        base = env
        while isinstance(base, _DeviceWrapper):
            base = base.env
        self.unwrapped = base
        self.device, self.num_envs, self.capacity = base.device, base.num_envs, base.capacity
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib.gpt_wrap_create(self.device.index, self.num_envs, self._FLAGS, float(gamma), float(epsilon),
                                          C.byref(handle)))
        self._w = handle
        z = lambda dt: torch.zeros(self.capacity, dtype=dt, device=self.device)
        self._t = {}
        if self._FLAGS & N.WRAP_RECORD:
            self._t.update(ep_return=z(torch.float32), ep_length=z(torch.int32), last_return=z(torch.float32),
                           last_length=z(torch.int32))
        if self._FLAGS & N.WRAP_NORMALIZE:
            self._t.update(disc_return=z(torch.float32), norm_reward=z(torch.float32))
        self._io = N.GptWrapIO()
        for k, v in self._t.items():
            setattr(self._io, k, v.data_ptr())

    def __getattr__(self, name):   # spaces, num_envs, state access ... come from the wrapped env
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def close(self):
        w, self._w = getattr(self, "_w", None), None
        if w:
            N.lib.gpt_wrap_destroy(w)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    # every stepping entry point goes through the wrapper's kernel; the multi-step / host-buffer paths would skip it
    # (statistics and the return RMS would silently stop updating), so they are refused instead of forwarded
    def _post(self, result):
        raise NotImplementedError

    def step(self, actions):
        return self._post(self.env.step(actions))

    def step_dlpack(self, actions):
        return self._post(self.env.step_dlpack(actions))

    def step_many(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__}.step_many would bypass the wrapper kernel: call step() per step, "
                                  "or use env.unwrapped.step_many() and forgo the wrapper's statistics")

    def step_host(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__}.step_host would bypass the wrapper kernel: call step() with "
                                  "device actions, or use env.unwrapped.step_host()")

    def _launch(self, reward, terminated, truncated):
        """reward / terminated / truncated: views of tensors with ``capacity`` rows (the env's output arrays)."""
        for name, t in (("reward", reward), ("terminated", terminated), ("truncated", truncated)):
            if (t.storage_offset() != 0 or not t.is_contiguous()
                    or t.untyped_storage().nbytes() < self.capacity * t.element_size()):
                raise ValueError(f"{name} must be the leading view of a tensor with `capacity` rows")
            setattr(self._io, name, t.data_ptr())
        with torch.cuda.device(self.device):
            N.check(N.lib.gpt_wrap_step(self._w, C.byref(self._io),
                                        C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def _state(self):
        p, off = C.c_void_p(), C.c_int32()
        N.check(N.lib.gpt_wrap_state_ptr(self._w, C.byref(p), C.byref(off)))
        with torch.cuda.device(self.device):
            return torch.as_tensor(_CudaBuffer(p.value, (18,), "<f8"), device=self.device), off.value

    @property
    def launch_count(self) -> int:
        return int(N.lib.gpt_wrap_launch_count(self._w))


class RecordEpisodeStatistics(_DeviceWrapper):
    """``info["episode"] = {"r": return, "l": length}`` of the episodes that ended on this step (0 elsewhere) and
    ``info["_episode"]`` = the done mask, like gymnasium's vector-aware wrapper; plus running whole-batch totals."""

    _FLAGS = N.WRAP_RECORD

    def __init__(self, env):
        super().__init__(env)

    def _post(self, result):
        obs, reward, terminated, truncated, info = result
        self._launch(reward, terminated, truncated)
        b = self.num_envs
        info = dict(info)
        info["episode"] = {"r": self._t["last_return"][:b], "l": self._t["last_length"][:b]}
        info["_episode"] = terminated | truncated
        return obs, reward, terminated, truncated, info

    @property
    def episode_returns(self) -> torch.Tensor:
        return self._t["ep_return"][: self.num_envs]

    @property
    def episode_lengths(self) -> torch.Tensor:
        return self._t["ep_length"][: self.num_envs]

    def stats_tensor(self) -> torch.Tensor:
        """float64[8] device view {episodes, sum_return, sum_length, sum_return^2, env_steps, 0, 0, 0}."""
        return self._state()[0][:8]

    def stats(self) -> dict:
        """Whole-job statistics: summed over ranks (NCCL all-reduce when torch.distributed is initialised)."""
        return allreduce_stats(self.stats_tensor().clone())


class NormalizeReward(_DeviceWrapper):
    """Scales rewards so that the exponential moving discounted return has unit variance (gymnasium
    ``NormalizeReward``): two fused launches per step, no host round trip."""

    _FLAGS = N.WRAP_NORMALIZE

    def __init__(self, env, gamma: float = 0.99, epsilon: float = 1e-8):
        super().__init__(env, gamma, epsilon)
        self.gamma, self.epsilon = gamma, epsilon

    def _post(self, result):
        obs, reward, terminated, truncated, info = result
        self._launch(reward, terminated, truncated)
        return obs, self._t["norm_reward"][: self.num_envs], terminated, truncated, info

    @property
    def returns(self) -> torch.Tensor:
        return self._t["disc_return"][: self.num_envs]

    def return_rms(self) -> dict:
        """Running {count, mean, var} of the discounted returns (one small device-to-host read)."""
        st, off = self._state()
        c, m, v = st[off:off + 3].cpu().tolist()
        return {"count": c, "mean": m, "var": v}

"""Host-side base class shared by the B200 vector envs.

Owns the torch tensors (state, outputs, replay draws) the C library works on, binds them
through DLPack, and implements the gym vector-env plumbing (``reset`` / ``step`` / state
access).  PyTorch is used only for device memory and streams; all env arithmetic happens in
libgpt_b200.so's fused kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
from torch.utils.dlpack import to_dlpack

from . import _native as N

_TORCH_DTYPES = {
    N.DT_U8: torch.uint8, N.DT_I8: torch.int8, N.DT_U16: torch.int16,  # uint16 is stored as int16 (values < 2^15)
    N.DT_I32: torch.int32, N.DT_F32: torch.float32, N.DT_F64: torch.float64,
}
_NUMPY_DTYPES = {
    N.DT_U8: np.uint8, N.DT_I8: np.int8, N.DT_U16: np.int16, N.DT_I32: np.int32, N.DT_F32: np.float32,
    N.DT_F64: np.float64,
}


#: GPT_DEBUG_ACTIONS=1: range-check discrete actions before every step and raise IndexError like the reference
#: (numpy fancy indexing, extended_taxi.py:248 / rooms.py:210); the hot kernels only mask the action byte.  Costs a
#: device synchronisation per step — a debugging aid, off by default.
DEBUG_ACTIONS = os.environ.get("GPT_DEBUG_ACTIONS", "0") not in ("", "0")

# raw cudaStream_t of torch's current stream without building a Stream object (about 1 us cheaper per step)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


class _CudaBuffer:
    """Zero-copy view of a device pointer the library owns (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


class DeviceVecEnv:
    """Common machinery; subclasses fill a :class:`GptConfig` and call :meth:`_create`."""

    is_vector_env = True

    # ------------------------------------------------------------------ construction
    def _create(self, cfg: N.GptConfig, *, device=None, rng_mode="philox", seed=None, env_offset=0,
                track_stats=False, keepalive=()):
        if not torch.cuda.is_available():
            raise RuntimeError("gym_po (B200 build) needs a CUDA device; there is no CPU fallback")
        if rng_mode not in ("philox", "replay"):
            raise ValueError("rng_mode must be 'philox' or 'replay'")
        dev = torch.device("cuda" if device is None else device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self._dev_index = dev.index
        self.rng_mode = rng_mode
        if seed is None:  # the reference seeds its generator from OS entropy
            seed = int(np.random.SeedSequence().generate_state(1, np.uint64)[0])
        cfg.abi_version = N.ABI_VERSION
        cfg.rng_mode = N.RNG_REPLAY if rng_mode == "replay" else N.RNG_PHILOX
        cfg.device = dev.index
        cfg.num_envs = int(self.num_envs)
        cfg.env_offset = int(env_offset)
        cfg.seed = int(seed) & (2**64 - 1)
        cfg.track_stats = int(bool(track_stats))
        self._keepalive = keepalive
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            N.check(N.lib.gpt_create(C.byref(cfg), C.byref(handle)))
        self._h = handle
        self.capacity = int(N.lib.gpt_capacity(self._h))
        self._arrays, self._descs = {}, {}
        desc = N.GptArrayDesc()
        for i in range(N.lib.gpt_array_count(self._h)):
            N.check(N.lib.gpt_array_info(self._h, i, C.byref(desc)))
            name = desc.name.decode()
            self._descs[name] = (i, desc.role, desc.dtype, desc.cols)
            if desc.role == N.ROLE_ACTION:
                self._action_dtype, self._action_cols = _TORCH_DTYPES[desc.dtype], desc.cols
                continue
            if desc.role == N.ROLE_REPLAY and rng_mode != "replay":
                continue
            shape = (self.capacity,) if desc.cols == 1 else (self.capacity, desc.cols)
            t = torch.zeros(shape, dtype=_TORCH_DTYPES[desc.dtype], device=dev)
            capsule = to_dlpack(t)  # must stay alive across the call: it owns the DLManagedTensor
            N.check(N.lib.gpt_bind_dlpack(self._h, i, N.dlpack_pointer(capsule)))
            del capsule
            self._arrays[name] = t
        ashape = (self.capacity,) if self._action_cols == 1 else (self.capacity, self._action_cols)
        self._action_pad = torch.zeros(ashape, dtype=self._action_dtype, device=dev)
        self._action_shape_cap = torch.Size(ashape)
        b = self.num_envs
        self._obs = self._shape_obs(self._arrays["obs"][:b])
        self._reward = self._arrays["reward"][:b]
        self._terminated = self._arrays["terminated"][:b].view(torch.bool)
        self._truncated = self._arrays["truncated"][:b].view(torch.bool)
        self._host = None
        self._pinned_user = []

    def _shape_obs(self, obs):
        return obs

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            N.lib.gpt_destroy(h)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        if _raw_stream is not None:
            return _raw_stream(self._dev_index)
        return torch.cuda.current_stream(self.device).cuda_stream

    def _on_device(self):
        return torch.cuda.device(self.device)

    def _n_discrete_actions(self):
        """Number of discrete actions (None for continuous-action envs) — used by the debug range check."""
        if self._action_dtype is not torch.int8:
            return None
        space = getattr(self, "single_action_space", None)
        return getattr(space, "n", None)

    def _debug_check_actions(self, actions):
        n = self._n_discrete_actions()
        if n is None:
            return
        t = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
        t = t[: self.num_envs]
        if t.numel() and bool(((t < 0) | (t >= n)).any()):   # before any narrowing cast to int8
            bad = t[(t < 0) | (t >= n)][0].item()
            raise IndexError(f"index {bad} is out of bounds for axis 0 with size {n}")

    def check_actions(self, actions: torch.Tensor) -> int:
        """Number of entries of an int8 device action tensor ``[capacity]`` outside ``[0, n)`` (``gpt_check_actions``;
        synchronises the stream)."""
        bad = C.c_int64()
        with self._on_device():
            N.check(N.lib.gpt_check_actions(self._h, C.c_void_p(actions.data_ptr()), self._stream(), C.byref(bad)))
        return bad.value

    def read_table(self, name: str, dtype) -> np.ndarray:
        """Host copy of one of the handle's static device tables (``gpt_table_read``; diagnostics / parity tests)."""
        n = C.c_int64()
        N.check(N.lib.gpt_table_read(self._h, name.encode(), None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        if n.value == 0:
            return buf.view(dtype)
        with self._on_device():
            N.check(N.lib.gpt_table_read(self._h, name.encode(), buf.ctypes.data_as(C.c_void_p), buf.nbytes, C.byref(n)))
        return buf.view(dtype)

    def _device_actions(self, actions):
        t = actions
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(actions))
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        if t.dtype != self._action_dtype:
            t = t.to(self._action_dtype)
        want_tail = () if self._action_cols == 1 else (self._action_cols,)
        if tuple(t.shape[1:]) != want_tail or t.shape[0] not in (self.num_envs, self.capacity):
            raise ValueError(f"actions must have shape ({self.num_envs},{','.join(map(str, want_tail))})")
        if t.shape[0] == self.capacity:
            return t.contiguous()
        self._action_pad[: self.num_envs].copy_(t)
        return self._action_pad

    def _results(self):
        return self._obs, self._reward, self._terminated, self._truncated, {}

    # ------------------------------------------------------------------ gym API
    def _reset(self, seed=None):
        with self._on_device():
            N.check(N.lib.gpt_reset(self._h, int(seed is not None), int(seed or 0) & (2**64 - 1), self._stream()))
        return self._obs

    def step(self, actions):
        """One fused kernel launch: transition + reward + done + autoreset + obs.

        ``actions``: CUDA tensor (int8 for discrete envs is passed zero-copy; other integer dtypes
        are converted on the device) or anything ``numpy.asarray`` accepts (uploaded).  Returns
        ``(obs, reward, terminated, truncated, {})`` as views of the env's output tensors, which are
        overwritten by the next ``step`` — ``clone()`` to keep them.
        """
        a = actions
        if DEBUG_ACTIONS:
            self._debug_check_actions(actions)
        # fast path: a contiguous device tensor of the action dtype with `capacity` rows goes straight to the library
        if not (type(a) is torch.Tensor and a.dtype is self._action_dtype and a.shape == self._action_shape_cap
                and a.device == self.device and a.is_contiguous()):
            a = self._device_actions(actions)
        if torch.cuda.current_device() == self._dev_index:
            rc = N.lib.gpt_step(self._h, a.data_ptr(), self._stream())
        else:
            with self._on_device():
                rc = N.lib.gpt_step(self._h, a.data_ptr(), self._stream())
        if rc:
            N.check(rc)
        return self._obs, self._reward, self._terminated, self._truncated, {}

    def step_dlpack(self, actions):
        """Like :meth:`step` for any ``__dlpack__`` producer; the C library validates device, dtype,
        shape ``[capacity(,cols)]`` and contiguity itself."""
        capsule = actions.__dlpack__() if hasattr(actions, "__dlpack__") else actions
        with self._on_device():
            N.check(N.lib.gpt_step_dlpack(self._h, N.dlpack_pointer(capsule), self._stream()))
        return self._results()

    def step_many(self, actions, out=None):
        """``T`` consecutive steps from an action stream ``[T, capacity(,cols)]`` (one launch per step,
        no host round trip).  With ``out`` = dict of rollout tensors ``[T, rows >= capacity, ...]`` named like
        the output arrays the results of step ``t`` land in ``out[name][t, :capacity]``."""
        t = actions
        if t.device != self.device or t.dtype != self._action_dtype or not t.is_contiguous() or t.shape[1] != self.capacity:
            raise ValueError("step_many needs a contiguous device tensor [T, capacity(,cols)] of the action dtype")
        stride = 0
        names = [n for n in ("obs", "reward", "terminated", "truncated") if self._descs[n][1] == N.ROLE_OUTPUT]
        if out is not None:
            for name in names:
                i, _, dt, cols = self._descs[name]
                o = out[name]
                if o.shape[0] < t.shape[0] or o.shape[1] < self.capacity or not o.is_contiguous():
                    raise ValueError(f"out['{name}'] must be contiguous [T, rows >= capacity, ...]")
                if stride not in (0, o.shape[1]):
                    raise ValueError("every out[...] tensor must have the same number of rows per rollout slot")
                stride = o.shape[1]          # rows between consecutive rollout slots (padded storage is fine)
                N.check(N.lib.gpt_bind(self._h, i, C.c_void_p(o.data_ptr()), o.shape[0] * o.shape[1]))
        try:
            with self._on_device():
                N.check(N.lib.gpt_step_many(self._h, C.c_void_p(t.data_ptr()), t.shape[0], stride, self._stream()))
        finally:
            if out is not None:
                for name in names:
                    i = self._descs[name][0]
                    a = self._arrays[name]
                    N.check(N.lib.gpt_bind(self._h, i, C.c_void_p(a.data_ptr()), a.shape[0]))
        return out

    # ------------------------------------------------------------------ host (end-to-end) path
    def _ensure_host(self):
        if self._host is None:
            b = self.num_envs
            _, _, odt, ocols = self._descs["obs"]
            pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
            ashape = (b,) if self._action_cols == 1 else (b, self._action_cols)
            self._host = {
                "actions": pin(ashape, self._action_dtype),
                "obs": pin((b,) if ocols == 1 else (b, ocols), _TORCH_DTYPES[odt]),
                "reward": pin((b,), torch.float32), "terminated": pin((b,), torch.uint8),
                "truncated": pin((b,), torch.uint8),
            }
            self._host_np = {k: v.numpy() for k, v in self._host.items()}
            self._host_io = N.GptHostIO(*(C.c_void_p(self._host[k].data_ptr())
                                          for k in ("actions", "obs", "reward", "terminated", "truncated")), None)
        return self._host_np

    def host_action_buffer(self) -> np.ndarray:
        """The pinned action buffer ``step_host`` uploads from (fill it in place to skip a host copy)."""
        return self._ensure_host()["actions"]

    def pinned_actions(self, n_slots: int) -> np.ndarray:
        """``[n_slots, num_envs(,cols)]`` numpy array backed by page-locked memory.  ``step_host(arr[i])`` uploads
        straight from it (no intermediate host copy)."""
        self._ensure_host()
        shape = (n_slots,) + tuple(self._host["actions"].shape)
        t = torch.zeros(shape, dtype=self._action_dtype).pin_memory()
        self._pinned_user.append(t)
        return t.numpy()

    def host_bytes_per_step(self):
        """(h2d, d2h) bytes one ``step_host`` call moves over PCIe."""
        hn = self._ensure_host()
        return hn["actions"].nbytes, sum(hn[k].nbytes for k in ("obs", "reward", "terminated", "truncated"))

    def _pinned_pointer(self, a: np.ndarray):
        """Address of ``a`` if it is a C-contiguous slice of a buffer from :meth:`pinned_actions`, else None."""
        if not (isinstance(a, np.ndarray) and a.flags.c_contiguous and a.dtype == self._host_np["actions"].dtype
                and a.shape == self._host_np["actions"].shape):
            return None
        addr = a.ctypes.data
        for t in self._pinned_user:
            lo = t.data_ptr()
            if lo <= addr and addr + a.nbytes <= lo + t.numel() * t.element_size():
                return addr
        return None

    def step_host(self, actions: np.ndarray):
        """numpy in / numpy out through ``gpt_step_host``: pinned host buffers, chunked
        H2D -> fused step -> D2H pipeline.  Returned arrays are views of pinned buffers that the
        next ``step_host`` overwrites."""
        hn = self._ensure_host()
        io = self._host_io
        if actions is not hn["actions"]:
            addr = self._pinned_pointer(actions)
            if addr is not None:
                io = N.GptHostIO(C.c_void_p(addr), self._host_io.obs, self._host_io.reward, self._host_io.terminated,
                                 self._host_io.truncated, None)
            else:
                np.copyto(hn["actions"], actions, casting="unsafe")
        if DEBUG_ACTIONS:
            self._debug_check_actions(actions)
        # the library orders its internal streams behind (and the caller's stream after) the work on torch's current
        # stream: reset() / step() / set_state() followed by step_host() is race free
        io.stream = self._stream()
        with self._on_device():
            N.check(N.lib.gpt_step_host(self._h, C.byref(io)))
        return (self._shape_obs(hn["obs"]), hn["reward"], hn["terminated"].view(np.bool_),
                hn["truncated"].view(np.bool_), {})

    # ------------------------------------------------------------------ RNG / replay / stats
    def set_replay(self, **draws):
        """Replay mode: upload the dense per-env draws for the NEXT reset/step (oracle ``env.draws``)."""
        if self.rng_mode != "replay":
            raise RuntimeError("set_replay needs rng_mode='replay'")
        for k, v in draws.items():
            name = "replay_" + k
            if name not in self._arrays:
                continue  # draw kinds this configuration does not consume
            dst = self._arrays[name]
            src = torch.as_tensor(np.ascontiguousarray(v)).to(dst.dtype)
            dst[: self.num_envs].copy_(src.reshape(dst[: self.num_envs].shape))

    @property
    def rng_counter(self) -> int:
        c = C.c_uint64()
        N.check(N.lib.gpt_get_counter(self._h, C.byref(c)))
        return c.value

    @rng_counter.setter
    def rng_counter(self, value: int):
        N.check(N.lib.gpt_set_counter(self._h, int(value)))

    def set_fused_steps(self, enable=True):
        """``step_many`` as one fused multi-step launch where the family supports it (Taxi, ROOMS, MSRooms; Philox mode)
        — on by default.  ``enable``: False / True, or ``"tma"`` (I/O by TMA bulk copies where the family has it: Taxi,
        where it is the default) / ``"threads"`` (per-thread loads and stores) to pick the fused kernel's I/O path."""
        mode = {"tma": 2, "threads": 3}.get(enable, None) if isinstance(enable, str) else int(bool(enable))
        if mode is None:
            raise ValueError("set_fused_steps: True, False, 'tma' or 'threads'")
        N.check(N.lib.gpt_set_fused_steps(self._h, mode))

    def set_graph_mode(self, enable: bool = True):
        """Make ``step()`` / ``step_many()`` capturable into a CUDA graph (every family, Philox mode): the Philox step
        counter moves into device memory and the step kernels read and advance it themselves, so every replay of a
        captured graph draws fresh random numbers (fused launches and programmatic dependent launch keep working).
        Call it outside of stream capture; ``step_host`` is unavailable while it is on."""
        with self._on_device():
            N.check(N.lib.gpt_set_graph_mode(self._h, int(bool(enable)), self._stream()))

    @property
    def launch_count(self) -> int:
        return int(N.lib.gpt_launch_count(self._h))

    def stats_tensor(self) -> torch.Tensor:
        """float64[8] device tensor {episodes, sum_return, sum_length, sum_return^2, env_steps, 0, 0, 0}
        living in the handle (zero-copy).  All-reduce it over ranks with
        ``torch.distributed.all_reduce`` (NCCL) to get whole-job statistics."""
        p = C.c_void_p()
        N.check(N.lib.gpt_stats_ptr(self._h, C.byref(p)))
        with self._on_device():
            return torch.as_tensor(_CudaBuffer(p.value, (8,), "<f8"), device=self.device)

    def stats_reset(self):
        with self._on_device():
            N.check(N.lib.gpt_stats_reset(self._h, self._stream()))

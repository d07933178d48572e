"""ctypes binding of libgpt_b200.so (include/gpt_b200.h).

There is no CPU fallback: importing this module without the built library raises
ImportError, and creating an env without a CUDA device raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPT_B200_LIB", os.path.join(_HERE, "..", "lib", "libgpt_b200.so"))

ABI_VERSION = 2
ENV_ALIGN = 512

FAMILY_TAXI, FAMILY_ROOMS, FAMILY_CROOMS, FAMILY_TAG, FAMILY_CAR, FAMILY_MSROOMS = 0, 1, 2, 3, 4, 5
RNG_PHILOX, RNG_REPLAY = 0, 1
(OBS_ROOM, OBS_ROOM_GOAL, OBS_MDP, OBS_MDP_GOAL, OBS_VEC_MDP, OBS_VEC_MDP_GOAL, OBS_HANSEN, OBS_VEC_HANSEN,
 OBS_VEC_HANSEN_GOAL, OBS_GRID) = range(10)
ROLE_STATE, ROLE_OUTPUT, ROLE_REPLAY, ROLE_ACTION = 0, 1, 2, 3
DT_U8, DT_I8, DT_U16, DT_I32, DT_F32, DT_F64 = range(6)


class GptConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("family", C.c_int32), ("rng_mode", C.c_int32), ("device", C.c_int32),
        ("num_envs", C.c_int64), ("env_offset", C.c_int64), ("seed", C.c_uint64),
        ("time_limit", C.c_int32), ("track_stats", C.c_int32),
        # taxi
        ("taxi_rows", C.c_int32), ("taxi_cols", C.c_int32), ("taxi_nlocs", C.c_int32),
        ("taxi_n_dropoffs", C.c_int32), ("taxi_hansen_obs", C.c_int32),
        ("taxi_reward_goal", C.c_float), ("taxi_reward_bad", C.c_float), ("taxi_reward_any", C.c_float),
        ("taxi_wall_bits", C.POINTER(C.c_uint8)), ("taxi_loc_cell", C.POINTER(C.c_int32)),
        ("taxi_n_valid", C.c_int32), ("taxi_valid_states", C.POINTER(C.c_int32)),
        ("taxi_reset_cdf", C.POINTER(C.c_uint32)),
        # rooms / crooms
        ("rooms_h", C.c_int32), ("rooms_w", C.c_int32), ("rooms_grid", C.POINTER(C.c_int8)),
        ("rooms_n_actions", C.c_int32), ("rooms_slip_cumsum", C.POINTER(C.c_double)),
        ("rooms_obs_kind", C.c_int32), ("rooms_obs_n", C.c_int32),
        ("rooms_goal_y", C.c_int32), ("rooms_goal_x", C.c_int32),
        ("rooms_step_reward", C.c_float), ("rooms_wall_reward", C.c_float), ("rooms_goal_reward", C.c_float),
        # continuous
        ("c_cell_size", C.c_double), ("c_action_std", C.c_double), ("c_action_power", C.c_double),
        ("c_goal_threshold", C.c_double), ("c_use_velocity", C.c_int32), ("c_action_f64", C.c_int32),
        ("c_state_f32", C.c_int32), ("car_num_actions", C.c_int32),
        ("car_action_table", C.POINTER(C.c_double)),
        # multistory rooms
        ("ms_floors", C.c_int32), ("ms_goal_cell", C.c_int32), ("ms_up_y", C.c_int32), ("ms_up_x", C.c_int32),
        ("ms_down_y", C.c_int32), ("ms_down_x", C.c_int32),
    ]


class GptArrayDesc(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("role", C.c_int32), ("dtype", C.c_int32), ("cols", C.c_int32),
                ("elem_size", C.c_int32)]


class GptHostIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p),
                ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("stream", C.c_void_p)]


class GptWrapIO(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("ep_return", C.c_void_p),
                ("ep_length", C.c_void_p), ("last_return", C.c_void_p), ("last_length", C.c_void_p),
                ("disc_return", C.c_void_p), ("norm_reward", C.c_void_p)]


WRAP_RECORD, WRAP_NORMALIZE = 1, 2

#: every symbol include/gpt_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gpt_create": (C.c_int, [C.POINTER(GptConfig), C.POINTER(C.c_void_p)]),
    "gpt_destroy": (C.c_int, [C.c_void_p]),
    "gpt_capacity": (C.c_int64, [C.c_void_p]),
    "gpt_array_count": (C.c_int, [C.c_void_p]),
    "gpt_array_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GptArrayDesc)]),
    "gpt_find_array": (C.c_int, [C.c_void_p, C.c_char_p]),
    "gpt_bind": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "gpt_bind_dlpack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gpt_reset": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]),
    "gpt_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpt_step_dlpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpt_step_many": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "gpt_set_fused_steps": (C.c_int, [C.c_void_p, C.c_int]),
    "gpt_set_graph_mode": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gpt_step_host": (C.c_int, [C.c_void_p, C.POINTER(GptHostIO)]),
    "gpt_get_counter": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "gpt_set_counter": (C.c_int, [C.c_void_p, C.c_uint64]),
    "gpt_set_env_offset": (C.c_int, [C.c_void_p, C.c_int64]),
    "gpt_stats_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gpt_stats_reset": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gpt_wrap_create": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_void_p)]),
    "gpt_wrap_destroy": (C.c_int, [C.c_void_p]),
    "gpt_wrap_step": (C.c_int, [C.c_void_p, C.POINTER(GptWrapIO), C.c_void_p]),
    "gpt_wrap_state_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "gpt_wrap_launch_count": (C.c_int64, [C.c_void_p]),
    "gpt_last_error": (C.c_char_p, []),
    "gpt_abi_version": (C.c_int, []),
    "gpt_launch_count": (C.c_int64, [C.c_void_p]),
    "gpt_check_actions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "gpt_table_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
}


def _load():
    path = os.path.abspath(LIB_PATH)
    if not os.path.isfile(path):
        raise ImportError(
            f"libgpt_b200.so not found at {path}: build it with `python __graft_entry__.py` "
            "(or `make -C gym-po-taxi_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.gpt_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.gpt_abi_version()} != {ABI_VERSION}")
    return lib


lib = _load()


def last_error() -> str:
    msg = lib.gpt_last_error()
    return msg.decode() if msg else ""


_EXC = {-1: ValueError, -2: RuntimeError, -3: RuntimeError, -4: TypeError}


def check(rc: int):
    if rc != 0:
        raise _EXC.get(rc, RuntimeError)(f"libgpt_b200: {last_error()} (code {rc})")


# PyCapsule("dltensor") -> DLManagedTensor*
_capsule_ptr = C.pythonapi.PyCapsule_GetPointer
_capsule_ptr.restype = C.c_void_p
_capsule_ptr.argtypes = [C.py_object, C.c_char_p]


def dlpack_pointer(capsule) -> int:
    return _capsule_ptr(capsule, b"dltensor")

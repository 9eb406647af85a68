"""ctypes binding of libhirl_b200.so (C ABI: include/hirl_b200.h).  No fallback: if the
library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

OUT_SCALARS = 8
ROW_WORDS, BC_WORDS = 32, 20
ACTOR_PARAMS, CRITIC_PARAMS = 138756, 276482
NET_ACTOR, NET_TARGET_ACTOR, NET_CRITIC, NET_TARGET_CRITIC, NET_BC_ACTOR = range(5)
BUF_ONLINE, BUF_EXPERT = 0, 1
ACT_RELU, ACT_LEAKY_RELU = 0, 1
E_INVALID, E_CUDA, E_UNSUPPORTED, E_NOMEM = -1, -2, -3, -4


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("batch", C.c_int32), ("online_capacity", C.c_int32),
                ("expert_capacity", C.c_int32), ("bc_rows", C.c_int32), ("max_steps", C.c_int32),
                ("activation", C.c_int32), ("replicas", C.c_int32), ("gamma", C.c_float), ("tau", C.c_float),
                ("noise_clamp", C.c_float), ("loss_lambda", C.c_float)]


class Step(C.Structure):
    _fields_ = [("n_online", C.c_int32), ("n_expert", C.c_int32), ("do_actor", C.c_int32), ("do_polyak", C.c_int32),
                ("bc_weight_now", C.c_float), ("bc_warm_up", C.c_float), ("force_soft_weight", C.c_float),
                ("noise", C.c_float * 4), ("critic_neg_step", C.c_float), ("critic_bc2_sqrt", C.c_float),
                ("actor_neg_step", C.c_float), ("actor_bc2_sqrt", C.c_float), ("mode", C.c_int32)]


class DrawJob(C.Structure):     # hirl_draw_job_t
    _fields_ = [("py", C.c_void_p), ("np", C.c_void_p), ("online_len", C.c_int32), ("expert_len", C.c_int32)]


STEP_DTYPE = np.dtype([("n_online", "<i4"), ("n_expert", "<i4"), ("do_actor", "<i4"), ("do_polyak", "<i4"),
                       ("bc_weight_now", "<f4"), ("bc_warm_up", "<f4"), ("force_soft_weight", "<f4"),
                       ("noise", "<f4", (4,)), ("critic_neg_step", "<f4"), ("critic_bc2_sqrt", "<f4"),
                       ("actor_neg_step", "<f4"), ("actor_bc2_sqrt", "<f4"), ("mode", "<i4")])
assert STEP_DTYPE.itemsize == C.sizeof(Step) == 64

_vp, _i, _sz, _i64 = C.c_void_p, C.c_int, C.c_size_t, C.c_int64
# every symbol include/hirl_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "hirl_abi_version": (_i, []),
    "hirl_last_error": (C.c_char_p, []),
    "hirl_agent_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "hirl_agent_destroy": (_i, [_vp]),
    "hirl_agent_sm_count": (_i, [_vp]),
    "hirl_agent_backend": (_i, [_vp]),
    "hirl_nccl_unique_id": (_i, [_vp]),
    "hirl_dp_fused_export": (_i, [_vp, _i, _i, _i, _vp]),
    "hirl_dp_fused_attach": (_i, [_vp, _vp, _i]),
    "hirl_allreduce_init": (_i, [_vp, _vp, _i, _i, _i]),
    "hirl_allreduce_attach": (_i, [_vp, _vp, _i, _i, _i]),
    "hirl_params_import": (_i, [_vp, _i, _i, _vp, _sz]),
    "hirl_params_export": (_i, [_vp, _i, _i, _vp, _sz]),
    "hirl_adam_import": (_i, [_vp, _i, _i, _vp, _vp, _sz]),
    "hirl_adam_export": (_i, [_vp, _i, _i, _vp, _vp, _sz]),
    "hirl_grads_export": (_i, [_vp, _i, _i, _vp, _sz]),
    "hirl_store_rows": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "hirl_bc_table_set": (_i, [_vp, _vp, _vp, _i]),
    "hirl_gather": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp]),
    "hirl_gather_bc": (_i, [_vp, _vp, _i, _vp, _vp]),
    "hirl_learn": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hirl_learn_async": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hirl_learn_wait": (_i, [_vp]),
    "hirl_learn_dev": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hirl_last_launch_count": (_i, [_vp]),
    "hirl_act": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "hirl_act_dev": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "hirl_store_transitions_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hirl_agent_stream": (_i, [_vp, C.POINTER(_vp)]),
    "hirl_host_sample_indices": (_i, [_vp, _i, _i, _vp]),
    "hirl_host_choice_noreplace": (_i, [_vp, _i, _i, _vp]),
    "hirl_host_choice_noreplace_async": (_i, [_vp, _i, _i, _vp]),
    "hirl_host_choice_many_async": (_i, [_vp, _i, _i, _i, _vp]),
    "hirl_host_async_wait": (_i, []),
    "hirl_host_draw_many": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _i]),
    "hirl_debug_stop_after": (_i, [_vp, _i]),
    "hirl_debug_profile": (_i, [_vp, _i, _vp, _sz]),
    "hirl_debug_export": (_i, [_vp, _i, _i, _vp, _sz, C.POINTER(_sz)]),
    "hirl_env_step_dev": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "hirl_env_step": (_i, [_i, _vp, _i64, _vp, _vp, _vp, _vp]),
    "hirl_env_offline_dev": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "hirl_env_offline": (_i, [_i, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "hirl_expert_scan": (_i, [_i, _vp, _vp, _i, _vp, _vp, C.POINTER(_i)]),
    "hirl_env_offline_f64": (_i, [_i, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "hirl_expert_scan_f64": (_i, [_i, _vp, _vp, _i, _vp, _vp, C.POINTER(_i)]),
    "hirl_measure_fp32_peak": (_i, [_i, C.POINTER(C.c_double)]),
}

_lib = None


class HirlError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Loads the shared library (building it first only if HIRL_B200_AUTOBUILD=1)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if os.environ.get("HIRL_B200_AUTOBUILD") == "1":
            _build.build()
        else:
            raise HirlError(f"{path} is missing: build it with `python -m hirl4ucav_b200.build` "
                            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    if lib.hirl_abi_version() != 1:
        raise HirlError("libhirl_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc == 0:
        return
    msg = load().hirl_last_error().decode(errors="replace")
    if rc == E_INVALID:
        raise ValueError(msg)
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == E_NOMEM:
        raise MemoryError(msg)
    raise HirlError(msg)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def i32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)

"""Batched Harfang env arithmetic on the GPU (reference hirl/environments/HarfangEnv_GYM.py).

The simulator (Harfang Dogfight Sandbox behind a TCP socket) is out of scope; what lives here
is the arithmetic the env performs on the simulator's answers:

  * ``step_math``        observation / reward / step_success / done for N raw states
                         (GYM.py:101-137, 160-169, 193-268; Infinite-env counters :519-527)
  * ``offline_reward``   the "for expert data" functions on normalised observations
                         (GYM.py:299-336), used by train_all.py:287-306
  * ``HarfangEnvMath``   scalar-call surface with the reference's method names
                         (get_reward / get_termination; get_loc_diff, GYM.py:299-301, is an internal
                         helper of get_reward there and is folded into the kernel) + fire rules

numpy inputs take the host entry points (H2D + kernel + D2H inside the call); torch CUDA
tensors take the asynchronous device entry points on the current stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib

F_LOCKED_PREV, F_LOCKED_NOW, F_MISSILE_PREV, F_MISSILE_NOW, F_FIRE_SUCCESS = 1, 2, 4, 8, 16
O_DONE, O_EPISODE_SUCCESS, O_FIRE_SUCCESS, O_FIRED, O_EXPERT_FIRE = 1, 2, 4, 8, 16


def pack_raw(p_ally, e_ally, p_opp, e_opp, angle_deg, health, act_fire, flags) -> np.ndarray:
    """(N,16) float32 raw records; layout documented in include/hirl_b200.h."""
    n = len(angle_deg)
    raw = np.empty((n, 16), np.float32)
    raw[:, 0:3], raw[:, 3:6], raw[:, 6:9], raw[:, 9:12] = p_ally, e_ally, p_opp, e_opp
    raw[:, 12], raw[:, 13], raw[:, 14] = angle_deg, health, act_fire
    raw[:, 15] = np.asarray(flags, np.uint32).view(np.float32)
    return raw


def decode_flags(flags):
    f = np.asarray(flags, np.int32)
    succ = ((f >> 8) & 0xFF).astype(np.uint8).view(np.int8).astype(np.int32)
    return {"done": (f & O_DONE) != 0, "episode_success": (f & O_EPISODE_SUCCESS) != 0,
            "fire_success": (f & O_FIRE_SUCCESS) != 0, "fired": (f & O_FIRED) != 0, "expert_fire": (f & O_EXPERT_FIRE) != 0,
            "step_success": succ}


def _f64c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def step_math(raw, counters=False, device=0):
    """obs (N,13) f32, reward (N,) f32, flags (N,) i32 [, counters int64[3]] for N raw states."""
    lib = _lib.load()
    if isinstance(raw, torch.Tensor):
        if not raw.is_cuda or raw.dtype != torch.float32 or not raw.is_contiguous():
            raise ValueError("device path needs a contiguous float32 CUDA tensor of shape (N,16)")
        n = raw.numel() // 16
        obs = torch.empty((n, 13), dtype=torch.float32, device=raw.device)
        rew = torch.empty((n,), dtype=torch.float32, device=raw.device)
        fl = torch.empty((n,), dtype=torch.int32, device=raw.device)
        cnt = torch.zeros((3,), dtype=torch.int64, device=raw.device) if counters else None
        with torch.cuda.device(raw.device):
            _lib.check(lib.hirl_env_step_dev(raw.data_ptr(), n, obs.data_ptr(), rew.data_ptr(), fl.data_ptr(),
                                             cnt.data_ptr() if counters else None, _stream_ptr()))
        return (obs, rew, fl, cnt) if counters else (obs, rew, fl)
    raw = _lib.f32c(raw).reshape(-1, 16)
    n = raw.shape[0]
    obs, rew, fl = np.empty((n, 13), np.float32), np.empty(n, np.float32), np.empty(n, np.int32)
    cnt = np.zeros(3, np.uint64)
    _lib.check(lib.hirl_env_step(device, _lib.ptr(raw), n, _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(fl),
                                 _lib.ptr(cnt) if counters else None))
    return (obs, rew, fl, cnt.astype(np.int64)) if counters else (obs, rew, fl)


def offline_reward(s, a, s2, device=0):
    """reward (N,) f32, step_success (N,) i8, done-of-next-state (N,) bool."""
    lib = _lib.load()
    if isinstance(s, torch.Tensor):
        n = s.numel() // 13
        rew = torch.empty((n,), dtype=torch.float32, device=s.device)
        su = torch.empty((n,), dtype=torch.int8, device=s.device)
        dn = torch.empty((n,), dtype=torch.uint8, device=s.device)
        with torch.cuda.device(s.device):
            _lib.check(lib.hirl_env_offline_dev(s.data_ptr(), a.data_ptr(), s2.data_ptr(), n, rew.data_ptr(), su.data_ptr(),
                                                dn.data_ptr(), _stream_ptr()))
        return rew, su, dn.bool()
    # float64 inputs (the expert CSV, Python floats) keep their precision: the reference compares them in float64
    f64 = any(np.asarray(x).dtype == np.float64 for x in (s, a, s2))
    cast = _f64c if f64 else _lib.f32c
    s, a, s2 = cast(s).reshape(-1, 13), cast(a).reshape(-1, 4), cast(s2).reshape(-1, 13)
    n = s.shape[0]
    rew, su, dn = np.empty(n, np.float32), np.empty(n, np.int8), np.empty(n, np.uint8)
    fn = lib.hirl_env_offline_f64 if f64 else lib.hirl_env_offline
    _lib.check(fn(device, _lib.ptr(s), _lib.ptr(a), _lib.ptr(s2), n, _lib.ptr(rew), _lib.ptr(su), _lib.ptr(dn)))
    return rew, su, dn.astype(bool)


def expert_scan(expert_states, expert_actions, device=0):
    """train_all.py:287-306 as one batched pass: rows (M,32) f32 and step_success (M,) i8."""
    f64 = np.asarray(expert_states).dtype == np.float64 or np.asarray(expert_actions).dtype == np.float64
    cast = _f64c if f64 else _lib.f32c
    es, ea = cast(expert_states).reshape(-1, 13), cast(expert_actions).reshape(-1, 4)
    n = es.shape[0]
    rows = np.empty((max(n - 1, 0), 32), np.float32)
    succ = np.empty(max(n - 1, 0), np.int8)
    kept = C.c_int(0)
    lib = _lib.load()
    fn = lib.hirl_expert_scan_f64 if f64 else lib.hirl_expert_scan
    _lib.check(fn(device, _lib.ptr(es), _lib.ptr(ea), n, _lib.ptr(rows), _lib.ptr(succ), C.byref(kept)))
    return rows[:kept.value], succ[:kept.value]


def fill_expert_buffer(agent, expert_states, expert_actions):
    """Drop-in for the expert-buffer loop of train_all.py:287-306."""
    rows, succ = expert_scan(expert_states, expert_actions, agent.device.index)
    agent.expert_buffer.store_rows(rows, succ)
    return len(rows)


class HarfangEnvMath:
    """The reference env's pure methods (GYM.py:299-336) and fire rules, same names, GPU-backed."""

    def __init__(self, device=0):
        self.device = device

    def get_reward(self, state, action, n_state):
        r, su, _ = offline_reward(np.asarray(state)[None], np.asarray(action)[None], np.asarray(n_state)[None], self.device)
        return float(r[0]), int(su[0])

    def get_termination(self, state):
        st = np.asarray(state)[None]
        z = np.zeros((1, 13), st.dtype if st.dtype == np.float64 else np.float32)
        _, _, dn = offline_reward(z, np.zeros((1, 4), z.dtype), st, self.device)
        return bool(dn[0])

    @staticmethod
    def expert_fires(obs):
        obs = np.asarray(obs)
        return (obs[..., 7] > 0) & (obs[..., 8] > 0)            # ai_data_col.py:63-64

    @staticmethod
    def agent_fires(action):
        return np.asarray(action)[..., 3] > 0                    # GYM.py:150

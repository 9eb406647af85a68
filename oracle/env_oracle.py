"""TEST INFRASTRUCTURE -- CPU oracle for the Harfang env arithmetic (numpy float64).

Restates, vectorised over N states, the arithmetic the reference performs one state
at a time in Python floats (paths relative to the reference checkout):

  * observation assembly            hirl/environments/HarfangEnv_GYM.py:193-268
                                    hirl/environments/constants.py:13-16
  * online reward + step_success    hirl/environments/HarfangEnv_GYM.py:101-137
    (+ counters of the Infinite env hirl/environments/HarfangEnv_GYM.py:497-535)
  * online termination              hirl/environments/HarfangEnv_GYM.py:160-169
  * offline loc_diff/reward/done    hirl/environments/HarfangEnv_GYM.py:299-336
                                    (second witness: hirl/data/calc_return.py:1-39)
  * fire rules                      hirl/data/straight_line/ai_data_col.py:63-64,
                                    hirl/environments/HarfangEnv_GYM.py:150
  * expert-buffer construction scan hirl/train_all.py:287-306

Never imported by the product package.  Pinned against the *unmodified* reference
class (run behind a gym stub and a fake dogfight_client) by oracle/gen_golden.py;
the reference has no tests of its own for this path ("parity unpinned").

Input convention (shared with the CUDA kernel, include/hirl_b200.h): one raw state
is 16 little-endian 32-bit words
    [0:3]  ally position (m)        [3:6]  ally Euler angles (rad)
    [6:9]  opponent position (m)    [9:12] opponent Euler angles (rad)
    [12]   target_angle (deg)       [13]   opponent health_level
    [14]   action[3] (fire command) [15]   uint32 flag word:
           bit0 locked_prev  bit1 locked_now  bit2 missile_prev  bit3 missile_now
           bit4 fire_success (sticky, value before this step)
All float32 inputs are widened to float64 before any arithmetic, exactly what the
reference would see if the simulator delivered those fp32-representable numbers.
"""
from __future__ import annotations

import math

import numpy as np

POS_NORM = 10000          # constants.py:14  NormStates["Plane_position"]
EULER_NORM = math.pi      # constants.py:15  NormStates["Plane_Euler_angles"]

F_LOCKED_PREV, F_LOCKED_NOW, F_MISSILE_PREV, F_MISSILE_NOW, F_FIRE_SUCCESS = 1, 2, 4, 8, 16
O_DONE, O_EPISODE_SUCCESS, O_FIRE_SUCCESS, O_FIRED, O_EXPERT_FIRE = 1, 2, 4, 8, 16


def split_raw(raw: np.ndarray):
    raw = np.ascontiguousarray(raw, dtype=np.float32).reshape(-1, 16)
    flags = raw[:, 15].copy().view(np.uint32)
    f64 = raw.astype(np.float64)
    return f64, flags


def pack_raw(p_ally, e_ally, p_opp, e_opp, angle_deg, health, act_fire, flags) -> np.ndarray:
    n = len(angle_deg)
    raw = np.empty((n, 16), np.float32)
    raw[:, 0:3], raw[:, 3:6], raw[:, 6:9], raw[:, 9:12] = p_ally, e_ally, p_opp, e_opp
    raw[:, 12], raw[:, 13], raw[:, 14] = angle_deg, health, act_fire
    raw[:, 15] = np.asarray(flags, np.uint32).view(np.float32)
    return raw


_LIB = None


def c_lib():
    """ctypes handle of oracle/_build/libenv_oracle.so (plain-C twin, see env_oracle.c), or None."""
    global _LIB
    if _LIB is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libenv_oracle.so")
        _LIB = ctypes.CDLL(path) if os.path.exists(path) else False
    return _LIB or None


def _pow(x, e: float):
    """x ** e exactly as the reference evaluates it on Python floats: libm pow(x, e).
    glibc's pow(x, 0.5) differs from sqrt(x) by one f64 ulp on ~0.08 % of inputs (and
    pow(x, 2) from x*x on a few per 10^4), so numpy's sqrt/square/SIMD power loops are NOT
    used.  Falls back to a Python-float loop when the C twin is not built."""
    x = np.ascontiguousarray(x, np.float64)
    lib = c_lib()
    if lib is not None:
        import ctypes
        y = np.empty_like(x)
        lib.env_oracle_pow(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(e), ctypes.c_long(x.size),
                           y.ctypes.data_as(ctypes.c_void_p))
        return y
    return np.array([v ** e for v in x.reshape(-1).tolist()], np.float64).reshape(x.shape)


def _norm3(a, b, c):
    return _pow(_pow(a, 2.0) + _pow(b, 2.0) + _pow(c, 2.0), 1 / 2)


def online_step_c(raw: np.ndarray):
    """Same as online_step, computed by the plain-C twin (multi-threaded)."""
    import ctypes
    lib = c_lib()
    if lib is None:
        raise RuntimeError("oracle/_build/libenv_oracle.so not built (make -C oracle)")
    raw = np.ascontiguousarray(raw, np.float32).reshape(-1, 16)
    n = raw.shape[0]
    obs = np.empty((n, 13), np.float64); rew = np.empty(n, np.float64); fl = np.empty(n, np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.env_oracle_online(vp(raw), ctypes.c_long(n), vp(obs), vp(rew), vp(fl))
    return obs, rew, fl


def online_step(raw: np.ndarray):
    """obs (N,13) f64, reward (N,) f64, out_flags (N,) int32 for N raw states.

    out_flags: bit0 done, bit1 episode_success, bit2 fire_success (sticky, after this
    step), bit3 fired (policy rule a[3] > 0, GYM.py:150), bit4 the expert's fire rule on the new observation
    (obs[7] > 0 and obs[8] > 0, data/straight_line/ai_data_col.py:63-64), byte1 = step_success as signed int8 (-1, 0, +1).
    """
    x, fl = split_raw(raw)
    pa, ea, po, eo = x[:, 0:3], x[:, 3:6], x[:, 6:9], x[:, 9:12]
    locked_prev = (fl & F_LOCKED_PREV) != 0
    locked_now = (fl & F_LOCKED_NOW) != 0
    missile_prev = (fl & F_MISSILE_PREV) != 0
    missile_now = (fl & F_MISSILE_NOW) != 0
    fire_success = (fl & F_FIRE_SUCCESS) != 0

    # -- observation (GYM.py:196-201, 213-218, 232-262): both positions are normalised
    #    first, then subtracted.
    obs = np.empty((x.shape[0], 13), np.float64)
    obs[:, 0:3] = pa / POS_NORM - po / POS_NORM
    obs[:, 3:6] = ea / EULER_NORM
    t_angle = x[:, 12] / 180
    obs[:, 6] = t_angle
    obs[:, 7] = np.where(locked_now, 1.0, -1.0)
    obs[:, 8] = np.where(missile_now, 1.0, -1.0)
    obs[:, 9:12] = eo / EULER_NORM
    obs[:, 12] = x[:, 13]

    # -- reward (GYM.py:101-137): sequential -= on a Python float
    alt = x[:, 1]
    d = pa - po
    loc = _norm3(d[:, 0], d[:, 1], d[:, 2])
    fired = x[:, 14] > 0                                     # GYM.py:150
    r = 0.0 - 0.0001 * loc
    r = r - t_angle * 10
    r = np.where(alt < 2000, r - 4, r)
    r = np.where(alt > 7000, r - 4, r)
    r = np.where(fired, r - 8, r)
    fail = fired & missile_prev & ~locked_prev
    hit = fired & missile_prev & locked_prev
    step_success = np.where(hit, 1, np.where(fail, -1, 0)).astype(np.int32)
    fire_success = fire_success | hit
    r = np.where((x[:, 13] <= 0.1) & fire_success, r + 600, r)

    # -- termination (GYM.py:160-169)
    done = (alt < 500) | (alt > 10000)
    ep = x[:, 13] <= 0
    done = done | ep
    out = (done.astype(np.int32) * O_DONE | ep.astype(np.int32) * O_EPISODE_SUCCESS
           | fire_success.astype(np.int32) * O_FIRE_SUCCESS | fired.astype(np.int32) * O_FIRED
           | ((obs[:, 7] > 0) & (obs[:, 8] > 0)).astype(np.int32) * O_EXPERT_FIRE      # ai_data_col.py:63-64
           | ((step_success & 0xFF) << 8))
    return obs, r, out.astype(np.int32)


def decode_flags(out_flags: np.ndarray):
    f = np.asarray(out_flags, np.int32)
    succ = ((f >> 8) & 0xFF).astype(np.uint8).view(np.int8).astype(np.int32)
    return {"done": (f & O_DONE) != 0, "episode_success": (f & O_EPISODE_SUCCESS) != 0,
            "fire_success": (f & O_FIRE_SUCCESS) != 0, "fired": (f & O_FIRED) != 0, "expert_fire": (f & O_EXPERT_FIRE) != 0,
            "step_success": succ}


def infinite_counters(out_flags: np.ndarray):
    """(total_fire, total_success, total_done): the Infinite env's running counters
    (GYM.py:519-527) as plain reductions over a batch of independent steps."""
    d = decode_flags(out_flags)
    return int((d["step_success"] != 0).sum()), int((d["step_success"] == 1).sum()), int(d["done"].sum())


# ---- offline ("for expert data") functions on normalised 13-vectors ------------------
def offline_loc_diff(s):
    s = np.asarray(s, np.float64).reshape(-1, 13)
    return _norm3(s[:, 0] * 10000, s[:, 1] * 10000, s[:, 2] * 10000)


def offline_reward(s, a, s2):
    """reward (N,) f64, step_success (N,) int32   (GYM.py:303-330)."""
    s = np.asarray(s, np.float64).reshape(-1, 13)
    a = np.asarray(a, np.float64).reshape(-1, 4)
    s2 = np.asarray(s2, np.float64).reshape(-1, 13)
    r = 0.0 - 0.0001 * offline_loc_diff(s2)
    r = r - s2[:, 6] * 10
    fired = a[:, 3] > 0
    r = np.where(fired, r - 8, r)
    have = s[:, 8] > 0
    succ = np.where(fired & have & (s[:, 7] > 0), 1, np.where(fired & have & (s[:, 7] < 0), -1, 0)).astype(np.int32)
    r = np.where(s2[:, 12] < 0.1, r + 600, r)
    return r, succ


def offline_done(s):
    s = np.asarray(s, np.float64).reshape(-1, 13)
    return s[:, 12] <= 0.1


def expert_fire_rule(obs):
    obs = np.asarray(obs, np.float64).reshape(-1, 13)
    return (obs[:, 7] > 0) & (obs[:, 8] > 0)     # ai_data_col.py:63-64


def expert_scan(states, actions):
    """The expert-buffer construction loop (train_all.py:287-306), restated.

    Walks i = 0.. ; row i is paired with row i+1; after a terminal next-state the walk
    skips one extra row.  Returns (use_idx, reward, done, step_success) for the rows
    that get stored, in storage order.  Like the reference, the walk needs row i+1, so
    it stops (the reference raises IndexError) at the last row; we stop cleanly.
    """
    states = np.asarray(states, np.float64)
    actions = np.asarray(actions, np.float64)
    n = states.shape[0]
    r_all, succ_all = offline_reward(states[:-1], actions[:-1], states[1:])
    done_all = offline_done(states[1:])
    use = []
    i = 0
    while i < n - 1:
        use.append(i)
        i += 2 if done_all[i] else 1
    use = np.asarray(use, np.int64)
    return use, r_all[use], done_all[use], succ_all[use]

"""TEST INFRASTRUCTURE -- drives the UNMODIFIED reference checkout (build container only).

The reference is pure Python and imports here (``/root/reference``); it does not
travel to the GPU box, so everything derived from it is frozen into
``tests/golden/`` by ``oracle/gen_golden.py``.  Nothing here is imported by the
product package, by the ``-m gpu`` tests, by ``smoke()`` or by ``bench.py``.

  * agent: ``hirl.agents.HIRL.Agent`` / ``hirl.agents.TD3.Agent`` / ``hirl.agents.BC.Agent``
    import as-is (torch, torchvision, numpy present).
  * env:   ``hirl.environments.HarfangEnv_GYM`` needs ``gym`` (absent) and talks to the
    simulator through ``dogfight_client``; we inject a stub ``gym.spaces.Box`` and replace
    the client's RPC functions with an in-process fake that serves one synthetic raw
    state.  ``_get_observation/_get_reward/_get_termination`` then run verbatim.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("HIRL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "hirl"))


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def hirl_agent_cls():
    _ensure_path()
    from hirl.agents.HIRL import Agent
    return Agent


def ref_memory_cls():
    _ensure_path()
    from hirl.utils.buffer import UniformMemory
    return UniformMemory


def td3_agent_cls():
    _ensure_path()
    sub = os.path.join(REF_ROOT, "hirl")   # TD3.py:7 uses the script-relative "utils.buffer"
    if sub not in sys.path:
        sys.path.insert(1, sub)
    from hirl.agents.TD3 import Agent
    return Agent


def bc_agent_cls():
    _ensure_path()
    from hirl.agents.BC import Agent
    return Agent


class FakeSim:
    """Serves one raw 16-word state record through the dogfight_client call surface."""

    def __init__(self):
        self.rec = None
        self.fired = False

    def load(self, rec):
        self.rec = np.asarray(rec, np.float32)
        self.flags = int(self.rec[15:16].view(np.uint32)[0])

    # --- the RPCs the env math reads (GYM.py:195, 209, 236, 245)
    def get_plane_state(self, pid):
        r = [float(v) for v in self.rec]   # JSON delivers Python floats
        if pid == "ally_1":
            return {"position": [r[0], r[1], r[2]], "Euler_angles": [r[3], r[4], r[5]],
                    "heading": 0.0, "pitch_attitude": 0.0, "roll_attitude": 0.0,
                    "user_pitch_level": 0.0, "user_yaw_level": 0.0, "user_roll_level": 0.0,
                    "target_locked": bool(self.flags & 2), "target_angle": r[12]}
        return {"position": [r[6], r[7], r[8]], "Euler_angles": [r[9], r[10], r[11]],
                "heading": 0.0, "pitch_attitude": 0.0, "roll_attitude": 0.0,
                "user_pitch_level": 0.0, "user_yaw_level": 0.0, "user_roll_level": 0.0,
                "target_locked": False, "target_angle": 0.0}

    def get_health(self, pid):
        return {"health_level": float(self.rec[13])}

    def get_missiles_device_slots_state(self, pid):
        return {"missiles_slots": [bool(self.flags & 8)]}

    def get_machine_missiles_list(self, pid):
        return ["m0"]


def harfang_env(kind="HarfangEnv"):
    """Returns (env, sim): an instance of the unmodified reference env bound to a FakeSim."""
    _ensure_path()
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Box:
            def __init__(self, low, high, dtype=np.float64):
                self.low, self.high, self.dtype = low, high, dtype

            def sample(self):
                return np.random.uniform(self.low, self.high)

        spaces.Box = Box
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    import hirl.environments.dogfight_client as df
    sim = FakeSim()
    noop = lambda *a, **k: None
    for name in dir(df):
        if name.startswith("_") or not callable(getattr(df, name)) or name in ("json",):
            continue
        setattr(df, name, getattr(sim, name, noop))
    import hirl.environments.HarfangEnv_GYM as G
    env = getattr(G, kind)()
    # opponent-script state normally set by reset() (simulator side, out of scope; GYM.py:376-392)
    env.serpentine_step, env.duration, env.oppo_yaw = 0, 250, 0.1
    return env, sim


def run_online_steps(raw: np.ndarray, kind="HarfangEnv"):
    """Feeds every raw record through env.step() of the unmodified reference.

    Returns obs (N,13) f64, reward (N,) f64, and the flag word in the oracle's encoding.
    """
    env, sim = harfang_env(kind)
    raw = np.ascontiguousarray(raw, np.float32).reshape(-1, 16)
    n = raw.shape[0]
    obs = np.empty((n, 13)); rew = np.empty(n); out = np.empty(n, np.int32)
    with quiet():
        for i in range(n):
            sim.load(raw[i])
            fl = sim.flags
            env.n_Ally_target_locked = bool(fl & 1)   # becomes Ally_target_locked in _get_observation
            env.n_missile1_state = bool(fl & 4)       # becomes missile1_state
            env.fire_success = bool(fl & 16)
            env.done = False
            env.episode_success = False
            env.state = None
            action = np.array([0.0, 0.0, 0.0, float(raw[i, 14])])
            n_state, reward, done, _, success = env.step(action)
            obs[i] = n_state
            rew[i] = reward
            out[i] = (int(bool(done)) | int(bool(env.episode_success)) << 1 | int(bool(env.fire_success)) << 2
                      | int(bool(env.now_missile_state)) << 3
                      | int(bool(n_state[7] > 0 and n_state[8] > 0)) << 4     # the expert's rule, verbatim from data/straight_line/ai_data_col.py:63-64
                      | ((int(success) & 0xFF) << 8))
    return obs, rew, out


def run_offline(states, actions, next_states):
    env, _ = harfang_env()
    n = len(states)
    r = np.empty(n); succ = np.empty(n, np.int32); done = np.empty(n, bool); loc = np.empty(n)
    for i in range(n):
        r[i], succ[i] = env.get_reward(states[i], actions[i], next_states[i])
        done[i] = env.get_termination(next_states[i])
        loc[i] = env.get_loc_diff(next_states[i])
    return r, succ, done, loc

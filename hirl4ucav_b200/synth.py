"""Synthetic inputs of the Harfang env's shape (the simulator itself is out of scope).

Distributions follow SURVEY.md section 8(d).  A private numpy Generator is used so the
three global RNG streams the agent consumes (``random``, ``numpy.random``, torch CPU;
reference hirl/utils/seed.py:5-10) are left untouched.
"""
from __future__ import annotations

import numpy as np

STATE_DIM, ACTION_DIM = 13, 4
F32_TENTH = float(np.float32(0.1))


def env_states(rng: np.random.Generator, n: int) -> np.ndarray:
    """(n,13) float64 observation-shaped rows whose values are fp32-representable."""
    s = np.empty((n, STATE_DIM), np.float64)
    s[:, 0:3] = rng.uniform(-0.6, 0.6, (n, 3))
    s[:, 3:6] = rng.uniform(-1, 1, (n, 3))
    s[:, 6] = rng.uniform(0, 1, n)
    s[:, 7] = np.where(rng.random(n) < 0.2, 1.0, -1.0)
    s[:, 8] = np.where(rng.random(n) < 0.8, 1.0, -1.0)
    s[:, 9:12] = rng.uniform(-1, 1, (n, 3))
    s[:, 12] = rng.choice(np.array([0.2, F32_TENTH, 0.05, 0.0]), n, p=[0.9, 0.04, 0.03, 0.03])
    return s.astype(np.float32).astype(np.float64)


def env_actions(rng: np.random.Generator, n: int, expert: bool = False) -> np.ndarray:
    a = rng.uniform(-1, 1, (n, ACTION_DIM))
    if expert:  # expert fire channel is +-1 (reference hirl/data/straight_line/ai_env.py:14-30)
        a[:, 3] = np.where(rng.random(n) < 0.01, 1.0, -1.0)
    return a.astype(np.float32).astype(np.float64)


def raw_states(rng: np.random.Generator, n: int) -> np.ndarray:
    """(n,16) float32 raw pursuit records (layout: include/hirl_b200.h, HirlRawState)."""
    raw = np.empty((n, 16), np.float32)
    for base in (0, 6):
        raw[:, base + 0] = rng.uniform(-5000, 5000, n)
        raw[:, base + 1] = rng.uniform(300, 10500, n)
        raw[:, base + 2] = rng.uniform(-5000, 5000, n)
    raw[:, 3:6] = rng.uniform(-np.pi, np.pi, (n, 3))
    raw[:, 9:12] = rng.uniform(-np.pi, np.pi, (n, 3))
    raw[:, 12] = rng.uniform(0, 180, n)
    edge = np.array([0.2, F32_TENTH, np.nextafter(np.float32(0.1), np.float32(1)),
                     np.nextafter(np.float32(0.1), np.float32(0)), 0.05, 0.0, 1.0, 0.2], np.float32)
    raw[:, 13] = edge[rng.integers(0, len(edge), n)]
    raw[:, 14] = rng.uniform(-1, 1, n)
    raw[:, 15] = rng.integers(0, 32, n, dtype=np.uint32).view(np.float32)
    # altitude edge cases
    k = max(1, n // 64)
    for v in (500.0, 2000.0, 7000.0, 10000.0, 499.99997, 10000.001):
        raw[rng.integers(0, n, k), 1] = v
    raw[rng.integers(0, n, k), 14] = 0.0
    return raw


def fill_agent(agent, n_online: int, n_expert_rows_hint: int = 0, data_seed: int = 1234, expert_states=None, expert_actions=None):
    """Fills an agent's online ring with n_online self-consistent synthetic transitions (reward /
    done / step_success from the offline reward kernel) and its expert ring from the expert table
    (the batched form of train_all.py:287-306).  Everything numerical runs on the GPU."""
    from .environments import harfang_math as hm
    rng = np.random.default_rng(data_seed + 1)
    s, a, s2 = env_states(rng, n_online), env_actions(rng, n_online), env_states(rng, n_online)
    r, succ, done = hm.offline_reward(s, a, s2, agent.device.index)
    rows = np.empty((n_online, 32), np.float32)
    rows[:, 0:13], rows[:, 13:17], rows[:, 17:30] = s, a, s2
    rows[:, 30], rows[:, 31] = r, done
    agent.buffer.store_rows(rows, succ)
    if expert_states is not None:
        hm.fill_expert_buffer(agent, expert_states, expert_actions)


def expert_table(n: int, data_seed: int = 1234):
    rng = np.random.default_rng(data_seed)
    return env_states(rng, n), env_actions(rng, n, expert=True)

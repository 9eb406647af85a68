"""TEST INFRASTRUCTURE -- seeded synthetic scenarios shared by golden generation, the parity
tests and bench.py's CPU legs.  Works with any class exposing the reference Agent surface
(hirl/agents/HIRL.py:149-215): the reference itself, oracle.hirl_oracle.OracleAgent and the
product hirl4ucav_b200.agents.HIRL.Agent.

Sizes/hyper-parameters: SURVEY.md section 8(d) / reference hirl/train_all.py:190-208.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from hirl4ucav_b200 import synth
from oracle import env_oracle

HP = dict(actorLR=1e-3, criticLR=1e-3, stateDim=13, actionDim=4, full1Dim=256, full2Dim=512,
          tau=0.005, gamma=0.99, bufferSize=10 ** 5, layerNorm=True, name="Harfang_GYM",
          bc_weight=0.5, expert_warm_up=True)


def seed_all(seed: int) -> None:
    """reference hirl/utils/seed.py:5-10 (CUDA seeding is irrelevant on the CPU oracle)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def make_scenario(n_online: int, n_expert: int, data_seed: int = 1234) -> dict:
    rng = np.random.default_rng(data_seed)
    es = synth.env_states(rng, n_expert)
    ea = synth.env_actions(rng, n_expert, expert=True)
    s = synth.env_states(rng, n_online)
    a = synth.env_actions(rng, n_online)
    s2 = synth.env_states(rng, n_online)
    r, succ = env_oracle.offline_reward(s, a, s2)
    d = env_oracle.offline_done(s2)
    use, er, ed, esucc = env_oracle.expert_scan(es, ea)
    return dict(expert_states=es, expert_actions=ea, s=s, a=a, s2=s2, r=r, d=d, succ=succ,
                ex_use=use, ex_r=er, ex_d=ed, ex_succ=esucc)


def build_agent(cls, sc: dict, batch: int, seed: int = 0, **extra):
    """Seeds the three global streams, constructs ``cls`` with the reference's positional
    signature and fills both replay rings through the public store() calls."""
    seed_all(seed)
    hp = dict(HP)
    hp.update(extra.pop("hp", {}))
    agent = cls(hp["actorLR"], hp["criticLR"], hp["stateDim"], hp["actionDim"], hp["full1Dim"], hp["full2Dim"],
                hp["tau"], hp["gamma"], hp["bufferSize"], batch, hp["layerNorm"], hp["name"],
                sc["expert_states"], sc["expert_actions"], hp["bc_weight"], hp["expert_warm_up"], **extra)
    fill_agent(agent, sc)
    return agent


def fill_agent(agent, sc: dict) -> None:
    s, a, s2, r, d, succ = sc["s"], sc["a"], sc["s2"], sc["r"], sc["d"], sc["succ"]
    for i in range(s.shape[0]):
        agent.store(s[i], a[i], s2[i], float(r[i]), bool(d[i]), int(succ[i]))
    es, ea = sc["expert_states"], sc["expert_actions"]
    for j, i in enumerate(sc["ex_use"]):
        agent.expert_buffer.store(es[i], ea[i], es[i + 1], float(sc["ex_r"][j]), bool(sc["ex_d"][j]), int(sc["ex_succ"][j]))


# the three BASELINE.json single-agent update configs
CONFIGS = {
    # name: (batch, bc_weight_now per call, expert_num_now)
    "soft256": dict(batch=256, bc_weight_now=100, expert_num_now=64),
    "fixed4096": dict(batch=4096, bc_weight_now=0.5, expert_num_now=1024),
    "linear4096": dict(batch=4096, bc_weight_now=0.5 - 3 / 5000, expert_num_now=0),
}

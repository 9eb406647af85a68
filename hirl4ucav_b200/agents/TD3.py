"""Drop-in for ``hirl.agents.TD3.Agent`` (reference hirl/agents/TD3.py:148-271): the same fused
kernel in mode 1 -- leaky_relu(0.01) towers (TD3.py:61-66, 131-132), no BC / expert terms,
actor loss = -mean Q1(s, pi(s)) (TD3.py:245-249)."""
from __future__ import annotations

from .. import _lib
from .HIRL import Agent as _HIRLAgent


class Agent(_HIRLAgent):
    _MODE = 1
    _GAIN = "leaky_relu"
    _ACT = _lib.ACT_LEAKY_RELU

    def __init__(self, actorLR, criticLR, stateDim, actionDim, full1Dim, full2Dim, tau, gamma, bufferSize, batchSize,
                 layerNorm, name, device=None, max_fused_steps=64):
        super().__init__(actorLR, criticLR, stateDim, actionDim, full1Dim, full2Dim, tau, gamma, bufferSize, batchSize,
                         layerNorm, name, None, None, 0.0, False, device=device, max_fused_steps=max_fused_steps)

    def learn(self, forced=None):
        out = super().learn(0.0, 0, 0, forced=forced)
        return out[0], out[1]

    def learn_many(self, n_calls, forced=None):
        return [(o[0], o[1]) for o in super().learn_many([(0.0, 0, 0)] * int(n_calls), forced=forced)]

"""Drop-in for ``hirl.agents.BC.Agent`` (reference hirl/agents/BC.py:136-197): behaviour-cloning
pre-training of the actor that HIRL-soft later loads as ``bc_actor``.  Same fused kernel, mode 2:
expert-state actor pass (leaky_relu towers, BC.py:129-130), plain MSE (no loss_lambda, BC.py:179),
actor Adam.  The reference quirk is preserved by construction: this leaky_relu checkpoint is later
evaluated by HIRL's relu Actor (HIRL.py:180 vs BC.py:129)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..utils import fast_rng
from .HIRL import Agent as _HIRLAgent, _adam_scalars


class Agent(_HIRLAgent):
    _MODE = 2
    _GAIN = "leaky_relu"
    _ACT = _lib.ACT_LEAKY_RELU

    def __init__(self, actorLR, stateDim, actionDim, full1Dim, full2Dim, layerNorm, name, batchsize, expert_states,
                 expert_actions, device=None, max_fused_steps=64):
        # the reference constructs ONE network (BC.py:138): only it may consume the init RNG stream
        self._bc_name = name
        super().__init__(actorLR, actorLR, stateDim, actionDim, full1Dim, full2Dim, 0.0, 0.0, 16, batchsize, layerNorm, name,
                         expert_states, expert_actions, 0.0, False, device=device, max_fused_steps=max_fused_steps)
        self.batchsize = self.batchSize
        self.bc_loss = np.float32(0)

    def _make_nets(self, stateDim, actionDim, full1Dim, full2Dim, name):
        from .. import nets
        self.actor = nets.make_actor(stateDim, actionDim, full1Dim, full2Dim, name, self._GAIN)   # BC.py:138: Actor(..., name)
        with torch.random.fork_rng(devices=[]):   # device-side placeholders must not touch the reference's RNG stream
            self.targetActor = nets.make_actor(stateDim, actionDim, full1Dim, full2Dim, "TargetActor_" + name, self._GAIN)
            self.critic = nets.make_critic(stateDim, actionDim, full1Dim, full2Dim, "Critic_" + name, self._GAIN)
            self.targetCritic = nets.make_critic(stateDim, actionDim, full1Dim, full2Dim, "TargetCritic_" + name, self._GAIN)
        self.bc_actor = None

    def train_actor(self, forced_bc=None):
        k = 0
        B = self.batchSize
        st = self._st_steps[k]
        if forced_bc is not None:
            idx = forced_bc
        else:   # BC.py:167: np.random.choice(samples_num, batchsize, replace=False), pre-drawn during the last launch when possible
            pre = self._pre_np.take([(len(self.expert_states), B)])
            idx = pre[0] if pre is not None else fast_rng.choice_noreplace(len(self.expert_states), B)
        self._st_bc[k, :] = idx
        if self.expert_upsample and forced_bc is None:   # BC.py:171-176
            print(f"upsample num is {len(self.target_indices)}")
            selected = np.random.choice(self.target_indices)
            self._st_bc[k, np.random.randint(B)] = selected
        self.last_draws = {"bc": self._st_bc[k, :].copy()}
        self._actor_t += 1
        a_neg, a_bs = _adam_scalars(self.actorLR, self._actor_t)
        st["n_online"], st["n_expert"], st["do_actor"], st["do_polyak"] = 0, 0, 1, 0
        st["bc_weight_now"], st["bc_warm_up"], st["force_soft_weight"] = 1.0, 0.0, -1.0
        st["noise"] = 0.0
        st["critic_neg_step"], st["critic_bc2_sqrt"], st["actor_neg_step"], st["actor_bc2_sqrt"] = 0.0, 1.0, a_neg, a_bs
        st["mode"] = 2
        self._launch(1, predraw_for=None if forced_bc is not None else 0)
        self.bc_loss = np.float32(self._st_out[k][2])
        return self.bc_loss

    def learn(self, *a, **k):
        raise AttributeError("BC agent trains with train_actor() (reference BC.py:160)")

    def saveCheckpoints(self, ajan, model_name):
        self.pull_parameters()
        torch.save(self.actor.state_dict(), self._ckpt_path(ajan, model_name, self.actor.name))

    def loadCheckpoints(self, ajan, model_name):
        self.actor.load_state_dict(torch.load(self._ckpt_path(ajan, model_name, self.actor.name), map_location="cpu"))
        self.push_parameters()

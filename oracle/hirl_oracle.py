"""TEST INFRASTRUCTURE -- CPU oracle for the HIRL agent update (torch, CPU only).

This is a *restatement* of the reference algorithm, written from its behaviour,
used ONLY as the checker (tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs).  The product package never imports it.

What it restates (paths relative to the reference checkout):
  * network shapes, layer order and initialisation   hirl/agents/HIRL.py:19-53, 105-126
  * twin-Q forward / Q1-only forward                 hirl/agents/HIRL.py:55-97
  * actor forward                                    hirl/agents/HIRL.py:128-140
  * replay ring + uniform sampling                   hirl/utils/buffer.py:11-54
  * Agent construction order (RNG stream!)           hirl/agents/HIRL.py:149-190
  * action selection                                 hirl/agents/HIRL.py:192-212
  * learn()                                          hirl/agents/HIRL.py:221-334
  * Polyak / hard update                             hirl/agents/HIRL.py:11-17
  * TD3 / BC variants (leaky_relu towers)            hirl/agents/TD3.py:201-260, hirl/agents/BC.py:160-185

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by
the reference itself).  This oracle is pinned by running the *unmodified*
reference in the build container on identical inputs and RNG streams
(oracle/gen_golden.py); the two agree BIT-FOR-BIT (losses, indices, all weights)
and the resulting vectors are committed under tests/golden/.

The arithmetic itself lives in PyTorch (third-party, torch>=1.13 per the
reference's setup.py:8; 2.11.0 in this image): Linear/LayerNorm/relu/tanh/
mse_loss/autograd/optim.Adam are called here exactly as the reference calls them.
"""
from __future__ import annotations

import random
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

Row = namedtuple("Row", ("state", "action", "next_state", "reward", "done", "step_success"))

SOFT_SENTINEL = 100  # HIRL.py:299  (bc_weight == 100 selects the Q-filter weight)


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
def _dense(n_in: int, n_out: int, reinit: bool, nonlin: str = "relu") -> nn.Linear:
    """nn.Linear default init, then (hidden layers only) the kaiming re-draw of the
    weight -- consumes the global CPU generator in the reference's order
    (HIRL.py:25-35, 111-121)."""
    lin = nn.Linear(n_in, n_out)
    if reinit:
        nn.init.kaiming_uniform_(lin.weight, a=0.01, mode="fan_in", nonlinearity=nonlin)
    return lin


class _Tower(nn.Module):
    """in -> h1 -> h2 -> out MLP with optional LayerNorm before the activation.

    Sub-module names are given by the caller so the state_dict keys equal the
    reference's (full1/layernorm1/full2/layernorm2/final...)."""

    def __init__(self):
        super().__init__()

    @staticmethod
    def add(owner: nn.Module, n_in, h1, h2, n_out, names, nonlin="relu"):
        f1, l1, f2, l2, fo = names
        owner.add_module(f1, _dense(n_in, h1, True, nonlin))
        owner.add_module(l1, nn.LayerNorm(h1))
        owner.add_module(f2, _dense(h1, h2, True, nonlin))
        owner.add_module(l2, nn.LayerNorm(h2))
        owner.add_module(fo, _dense(h2, n_out, False))

    @staticmethod
    def run(owner: nn.Module, x, names, layer_norm, act, rec=None, tag=""):
        f1, l1, f2, l2, fo = names
        g = owner.__getattr__
        z1 = g(f1)(x)
        h1 = act(g(l1)(z1) if layer_norm else z1)
        z2 = g(f2)(h1)
        h2 = act(g(l2)(z2) if layer_norm else z2)
        out = g(fo)(h2)
        if rec is not None:
            for k, v in (("z1", z1), ("h1", h1), ("z2", z2), ("h2", h2), ("out", out)):
                if v.requires_grad:
                    v.retain_grad()
                rec[tag + k] = v
        return out


_ACTOR_NAMES = ("full1", "layernorm1", "full2", "layernorm2", "final")
_Q1_NAMES = ("full1", "layernorm1", "full2", "layernorm2", "final1")
_Q2_NAMES = ("full3", "layernorm3", "full4", "layernorm4", "final2")


def _leaky(x):
    return F.leaky_relu(x)  # default slope 0.01 (TD3.py:61-66, BC.py:129-130)


def _init_gain_name(act):
    # the TD3/BC files re-draw hidden weights with the leaky_relu gain (TD3.py:26-46, BC.py:107-114)
    return "relu" if act is F.relu else "leaky_relu"


class OracleActor(nn.Module):
    def __init__(self, lr, state_dim, n_actions, h1, h2, layer_norm, name, act=F.relu):
        super().__init__()
        self.layerNorm = layer_norm
        self._act = act
        _Tower.add(self, state_dim, h1, h2, n_actions, _ACTOR_NAMES, _init_gain_name(act))
        self.optimizer = torch.optim.Adam(self.parameters(), lr=lr)
        self.name = name

    def forward(self, x, rec=None, tag=""):
        return torch.tanh(_Tower.run(self, x, _ACTOR_NAMES, self.layerNorm, self._act, rec, tag))


class OracleCritic(nn.Module):
    def __init__(self, lr, state_dim, n_actions, h1, h2, layer_norm, name, act=F.relu):
        super().__init__()
        self.layerNorm = layer_norm
        self._act = act
        _Tower.add(self, state_dim + n_actions, h1, h2, 1, _Q1_NAMES, _init_gain_name(act))
        _Tower.add(self, state_dim + n_actions, h1, h2, 1, _Q2_NAMES, _init_gain_name(act))
        self.optimizer = torch.optim.Adam(self.parameters(), lr=lr)
        self.name = name

    def forward(self, state, action, rec=None, tag=""):
        sa = torch.cat([state, action], 1)
        q1 = _Tower.run(self, sa, _Q1_NAMES, self.layerNorm, self._act, rec, tag + "q1.")
        q2 = _Tower.run(self, sa, _Q2_NAMES, self.layerNorm, self._act, rec, tag + "q2.")
        return q1, q2

    def onlyQ1(self, state, action, rec=None, tag=""):
        sa = torch.cat([state, action], 1)
        return _Tower.run(self, sa, _Q1_NAMES, self.layerNorm, self._act, rec, tag + "q1.")


def polyak_(target: nn.Module, source: nn.Module, tau: float) -> None:
    """t <- t*(1-tau) + p*tau, tensor by tensor (HIRL.py:11-13)."""
    with torch.no_grad():
        for t, p in zip(target.parameters(), source.parameters()):
            t.copy_(t * (1.0 - tau) + p * tau)


def clone_into_(target: nn.Module, source: nn.Module) -> None:
    with torch.no_grad():
        for t, p in zip(target.parameters(), source.parameters()):
            t.copy_(p)


# --------------------------------------------------------------------------------------
# replay ring  (hirl/utils/buffer.py:11-54; uniform path and the prioritised up-sampling path)
# --------------------------------------------------------------------------------------
class OracleMemory:
    def __init__(self, capacity, upsample=False):
        self.capacity = int(capacity)
        self.memory: list[Row] = []
        self.position = 0
        self.upsample = upsample
        self.priorities_memory = np.zeros(self.capacity) if upsample else None   # buffer.py:18
        self.last_indices: list[int] = []

    def store(self, *fields):
        row = Row(*fields)
        if self.upsample:   # buffer.py:21-29: successful shots weigh 1.0, everything else 0.1
            self.priorities_memory[self.position] = 1.0 if fields[-1] == 1 else 0.1
        if len(self.memory) < self.capacity:
            self.memory.append(row)
        else:
            self.memory[self.position] = row
        self.position = (self.position + 1) % self.capacity

    def sample(self, k, forced=None):
        # random.sample(list, k) and random.sample(range(len(list)), k) consume the
        # Mersenne stream identically and pick the same positions (buffer.py:45).
        if forced is not None:
            idx = list(forced)
        elif self.upsample:   # buffer.py:39-43: numpy stream, WITH replacement
            pr = self.priorities_memory[:len(self.memory)]
            idx = [int(i) for i in np.random.choice(len(self.memory), k, p=pr / np.sum(pr))]
        else:
            idx = random.sample(range(len(self.memory)), k)
        self.last_indices = idx
        picked = [self.memory[i] for i in idx]
        cols = Row(*zip(*picked))
        return cols.state, cols.action, cols.next_state, cols.reward, cols.done

    def fullEnough(self, k):
        return len(self.memory) >= k

    def __len__(self):
        return len(self.memory)


# --------------------------------------------------------------------------------------
# the agent
# --------------------------------------------------------------------------------------
class OracleAgent:
    """Restatement of hirl.agents.HIRL.Agent (HIRL.py:148-350) on CPU.

    ``dtype=torch.float64`` gives the "fp64 twin": weights are drawn in fp32 (same
    RNG stream) then widened; batch inputs are rounded through fp32 first, so the
    twin differs from the fp32 oracle only by arithmetic precision.
    """

    def __init__(self, actorLR, criticLR, stateDim, actionDim, full1Dim, full2Dim, tau, gamma,
                 bufferSize, batchSize, layerNorm, name, expert_states, expert_actions,
                 bc_weight, expert_warm_up, dtype=torch.float32, mode="HIRL"):
        self.dtype = dtype
        self.mode = mode  # "HIRL" | "TD3"
        self.tau, self.gamma = tau, gamma
        self.bufferSize, self.batchSize = bufferSize, batchSize
        self.actorTrainable = True
        self.actionDim = actionDim
        self.actionNoise = 0.1
        self.TD3LearningNoise = 0.2
        self.TD3LearningNoiseClamp = 0.5
        self.expert_states = expert_states
        self.expert_actions = expert_actions
        self.bc_weight = bc_weight
        self.expert_warm_up = expert_warm_up
        self.update_count = 0
        self.loss_lambda = 10000
        self.target_update_freq = 3
        self.buffer_upsample = False   # HIRL.py:184-185 (both off in the reference's constructor)
        self.expert_upsample = False
        if mode == "HIRL" and expert_actions is not None:
            self.target_indices = np.where(np.asarray(expert_actions)[:, 3] == 1)[0]   # HIRL.py:168, 217-219
        act = F.relu if mode == "HIRL" else _leaky

        # construction order fixes the init RNG stream (HIRL.py:172-180)
        mk_a = lambda tag: OracleActor(actorLR, stateDim, actionDim, full1Dim, full2Dim, layerNorm, tag + name, act)
        mk_c = lambda tag: OracleCritic(criticLR, stateDim, actionDim, full1Dim, full2Dim, layerNorm, tag + name, act)
        self.actor = mk_a("Actor_")
        self.targetActor = mk_a("TargetActor_")
        clone_into_(self.targetActor, self.actor)
        self.critic = mk_c("Critic_")
        self.targetCritic = mk_c("TargetCritic_")
        clone_into_(self.targetCritic, self.critic)
        self.bc_actor = mk_a("") if mode == "HIRL" else None
        if dtype != torch.float32:
            for net in self.nets().values():
                net.to(dtype)

        self.buffer = OracleMemory(bufferSize)
        self.expert_buffer = OracleMemory((len(expert_states) if expert_states is not None else 0) + 10)
        self.critic_loss = self.actor_loss = self.bc_loss = self.rl_loss = torch.zeros(())
        self.bc_fire_loss = 0.0
        self.last_draws: dict = {}

    def nets(self):
        d = {"actor": self.actor, "targetActor": self.targetActor, "critic": self.critic,
             "targetCritic": self.targetCritic}
        if self.bc_actor is not None:
            d["bc_actor"] = self.bc_actor
        return d

    # ---- acting (HIRL.py:192-212) ----------------------------------------------------
    def _act(self, state, sigma):
        s = torch.tensor(state, dtype=torch.float).to(self.dtype)
        a = self.actor(s)
        if sigma is not None:
            a = a + torch.normal(mean=torch.zeros(self.actionDim), std=torch.ones(self.actionDim) * sigma).to(self.dtype)
            a = a.clamp(-1, +1)
        return a.detach().to(torch.float32).numpy()

    def chooseAction(self, state):
        return self._act(state, self.actionNoise)

    def chooseActionSmallNoise(self, state):
        return self._act(state, self.actionNoise / 10)

    def chooseActionNoNoise(self, state):
        return self._act(state, None)

    def store(self, *fields):
        self.buffer.store(*fields)

    # ---- one update (HIRL.py:221-334) ----------------------------------------------
    def _f32(self, seq):
        # list-of-rows -> fp32 tensor, exactly the reference's conversion (HIRL.py:229-243),
        # then widened for the fp64 twin.
        return torch.tensor(np.array(seq), dtype=torch.float).to(self.dtype)

    def learn(self, bc_weight_now=0.0, expert_num_now=0, bc_warm_up_weight=0, forced=None, rec=None,
              force_soft_weight=None):
        forced = forced or {}
        B = self.batchSize
        # -- sampling ------------------------------------------------------------------
        if self.expert_warm_up and expert_num_now:
            s1, a1, n1, r1, d1 = self.buffer.sample(B - expert_num_now, forced.get("online"))
            idx_on = list(self.buffer.last_indices)
            s2, a2, n2, r2, d2 = self.expert_buffer.sample(expert_num_now, forced.get("expert"))
            idx_ex = list(self.expert_buffer.last_indices)
            S, A, S2 = self._f32(s1 + s2), self._f32(a1 + a2), self._f32(n1 + n2)
            R, D = self._f32(r1 + r2), self._f32(d1 + d2)
        else:
            s1, a1, n1, r1, d1 = self.buffer.sample(B, forced.get("online"))
            idx_on, idx_ex = list(self.buffer.last_indices), []
            S, A, S2, R, D = self._f32(s1), self._f32(a1), self._f32(n1), self._f32(r1), self._f32(d1)

        idx_bc = None
        if self.mode == "HIRL":
            # whole expert table is converted on every call in the reference (HIRL.py:245-246)
            bc_s_all = torch.tensor(np.array(self.expert_states), dtype=torch.float).to(self.dtype)
            bc_a_all = torch.tensor(np.array(self.expert_actions), dtype=torch.float).to(self.dtype)
            idx_bc = forced["bc"] if "bc" in forced else np.random.choice(bc_s_all.shape[0], B, replace=False)
            idx_bc = np.array(idx_bc)
            if self.expert_upsample and "bc" not in forced:   # HIRL.py:253-257: one row becomes a firing expert row
                selected = np.random.choice(self.target_indices)
                idx_bc[np.random.randint(self.batchSize)] = selected
            Sb, Ab = bc_s_all[idx_bc], bc_a_all[idx_bc]

        # -- TD target (HIRL.py:264-274) ----------------------------------------------------
        a_next = self.targetActor(S2, rec, "ta.")
        if "noise" in forced:
            noise = torch.as_tensor(forced["noise"], dtype=torch.float32)
        else:
            noise = torch.normal(mean=torch.zeros(self.actionDim), std=torch.ones(self.actionDim) * self.TD3LearningNoise)
        raw_noise = noise.clone()
        noise = noise.clamp(-self.TD3LearningNoiseClamp, self.TD3LearningNoiseClamp).to(self.dtype)
        a_next = (a_next + noise).clamp(-1, +1)
        tq1, tq2 = self.targetCritic(S2, a_next, rec, "tc.")
        tq_min = torch.min(tq1, tq2)
        y = R.reshape(-1, 1) + (self.gamma * tq_min * ((1 - D).reshape(-1, 1))).detach()
        self.last_draws = {"online": idx_on, "expert": idx_ex, "bc": idx_bc, "noise": raw_noise.numpy()}

        # -- critic step (HIRL.py:277-288) ---------------------------------------------------
        q1, q2 = self.critic(S, A, rec, "critic.")
        self.critic_loss = F.mse_loss(q1, y) + F.mse_loss(q2, y)
        self.critic.optimizer.zero_grad()
        self.critic_loss.backward()
        if rec is not None:
            rec["y"], rec["a_next"], rec["tq1"], rec["tq2"], rec["q1"], rec["q2"] = y, a_next, tq1, tq2, q1, q2
            rec["critic.grads"] = {k: p.grad.clone() for k, p in self.critic.named_parameters()}
            rec["S"], rec["A"], rec["S2"], rec["R"], rec["D"] = S, A, S2, R, D
        self.critic.optimizer.step()

        # -- actor step, every 2nd call (HIRL.py:291-332) ------------------------------------
        if self.actorTrainable:
            if self.mode == "HIRL":
                self.bc_weight = bc_weight_now
                pi = self.actor(S, rec, "actor.")
                rl_q = self.critic.onlyQ1(S, pi, rec, "qp.")
                self.rl_loss = -rl_q.mean()
                if self.bc_weight == SOFT_SENTINEL:
                    with torch.no_grad():
                        a_soft = self.bc_actor(S, rec, "bc.")
                    soft_q = self.critic.onlyQ1(S, a_soft, rec, "qs.")
                    w = (soft_q > rl_q).float().mean().detach().item()
                    if rec is not None:
                        rec["soft_q"], rec["rl_q"] = soft_q.detach(), rl_q.detach()
                        rec["soft_count_margin"] = (soft_q - rl_q).abs().min().item()
                    if force_soft_weight is not None:
                        w = float(force_soft_weight)
                    self.bc_weight = w + bc_warm_up_weight
                if self.bc_weight > 1:
                    self.bc_weight = 1
                pi_bc = self.actor(Sb, rec, "actor_bc.")
                self.bc_loss = F.mse_loss(pi_bc, Ab) * self.loss_lambda
                self.bc_fire_loss = F.mse_loss(pi_bc[:, 3].detach(), Ab[:, 3].detach()).item() * self.loss_lambda
                self.actor_loss = self.bc_loss * self.bc_weight + self.rl_loss * (1 - self.bc_weight)
            else:  # TD3.py:245-249
                pi = self.actor(S, rec, "actor.")
                self.actor_loss = -self.critic.onlyQ1(S, pi).mean()
            self.actor.optimizer.zero_grad()
            self.actor_loss.backward()
            if rec is not None:
                rec["actor.grads"] = {k: p.grad.clone() for k, p in self.actor.named_parameters()}
                rec["pi"] = pi
                if self.mode == "HIRL":
                    rec["pi_bc"], rec["rl_q_all"] = pi_bc, rl_q
                rec["Sb"], rec["Ab"] = (Sb, Ab) if self.mode == "HIRL" else (None, None)
            self.actor.optimizer.step()
            self.update_count += 1
            if self.update_count % self.target_update_freq == 0:
                polyak_(self.targetCritic, self.critic, self.tau)
                polyak_(self.targetActor, self.actor, self.tau)
        self.actorTrainable = not self.actorTrainable

        npf = lambda t: t.detach().to(torch.float32).numpy()
        if self.mode == "TD3":
            return npf(self.critic_loss.mean()), npf(self.actor_loss)
        return (npf(self.critic_loss.mean()), npf(self.actor_loss), npf(self.bc_loss), npf(self.rl_loss),
                self.bc_fire_loss, self.bc_weight)

    # ---- flat views used by the parity tests ------------------------------------------
    def flat(self, which):
        """All parameters of one net, state_dict order, as one fp64 numpy vector."""
        net = self.nets()[which]
        return np.concatenate([p.detach().double().reshape(-1).numpy() for p in net.parameters()])

    def state_dicts(self):
        return {k: {n: t.detach().clone() for n, t in net.state_dict().items()} for k, net in self.nets().items()}

    def adam_state(self, which):
        net = self.nets()[which]
        st = net.optimizer.state
        m = [st[p]["exp_avg"].double().reshape(-1).numpy() if p in st else np.zeros(p.numel()) for p in net.parameters()]
        v = [st[p]["exp_avg_sq"].double().reshape(-1).numpy() if p in st else np.zeros(p.numel()) for p in net.parameters()]
        return np.concatenate(m), np.concatenate(v)


class OracleBC:
    """Behaviour-cloning pre-trainer (BC.py:136-185): leaky_relu actor, plain MSE."""

    def __init__(self, lr, stateDim, actionDim, full1Dim, full2Dim, layerNorm, name, batchSize,
                 expert_states, expert_actions):
        self.actor = OracleActor(lr, stateDim, actionDim, full1Dim, full2Dim, layerNorm, name, _leaky)
        self.batchsize = batchSize
        self.expert_states, self.expert_actions = expert_states, expert_actions
        self.target_indices = np.where(np.asarray(expert_actions)[:, 3] == 1)[0]   # BC.py:152, 156-158
        self.expert_upsample = False                                               # BC.py:154
        self.bc_loss = torch.zeros(())

    def train_actor(self, forced_bc=None):
        s_all = torch.tensor(np.array(self.expert_states), dtype=torch.float)
        a_all = torch.tensor(np.array(self.expert_actions), dtype=torch.float)
        idx = np.array(forced_bc) if forced_bc is not None else np.random.choice(s_all.shape[0], self.batchsize, replace=False)
        if self.expert_upsample and forced_bc is None:   # BC.py:171-176
            selected = np.random.choice(self.target_indices)
            idx[np.random.randint(self.batchsize)] = selected
        self.last_bc = idx
        pred = self.actor(s_all[idx])
        self.bc_loss = F.mse_loss(pred, a_all[idx])
        self.actor.optimizer.zero_grad()
        self.bc_loss.backward()
        self.actor.optimizer.step()
        return self.bc_loss.detach().numpy()

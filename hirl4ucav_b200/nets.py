"""Host-side parameter containers of the five networks.

They exist for two reasons only: (1) initialisation must consume the torch CPU generator in
exactly the reference's order (HIRL.py:25-48, 111-121, 172-180) so that seeded runs start
from identical weights, and (2) checkpoints are ``state_dict`` files with the reference's key
names and shapes (HIRL.py:99-103, 142-146).  They never run a forward pass -- all arithmetic
happens in the CUDA library; the tensors here are a staging copy of the device arena.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def _dense(n_in, n_out, redraw, gain_name):
    lin = nn.Linear(n_in, n_out)
    if redraw:
        nn.init.kaiming_uniform_(lin.weight, a=0.01, mode="fan_in", nonlinearity=gain_name)
    return lin


class ParamNet(nn.Module):
    """Parameter holder with the reference's sub-module names."""

    def __init__(self, towers, name, gain_name="relu"):
        super().__init__()
        for n_in, h1, h2, n_out, names in towers:
            f1, l1, f2, l2, fo = names
            self.add_module(f1, _dense(n_in, h1, True, gain_name))
            self.add_module(l1, nn.LayerNorm(h1))
            self.add_module(f2, _dense(h1, h2, True, gain_name))
            self.add_module(l2, nn.LayerNorm(h2))
            self.add_module(fo, _dense(h2, n_out, False, gain_name))
        self.name = name
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, *a, **k):
        raise RuntimeError("host parameter containers do not compute; use the Agent methods (CUDA path)")

    def flat(self) -> np.ndarray:
        return np.concatenate([p.detach().reshape(-1).numpy() for p in self.parameters()]).astype(np.float32, copy=False)

    def load_flat(self, vec: np.ndarray) -> None:
        vec = np.asarray(vec, np.float32)
        off = 0
        with torch.no_grad():
            for p in self.parameters():
                n = p.numel()
                p.copy_(torch.from_numpy(vec[off:off + n].reshape(tuple(p.shape)).copy()))
                off += n
        assert off == vec.size


ACTOR_NAMES = ("full1", "layernorm1", "full2", "layernorm2", "final")
Q1_NAMES = ("full1", "layernorm1", "full2", "layernorm2", "final1")
Q2_NAMES = ("full3", "layernorm3", "full4", "layernorm4", "final2")


def make_actor(state_dim, n_actions, h1, h2, name, gain_name="relu"):
    return ParamNet([(state_dim, h1, h2, n_actions, ACTOR_NAMES)], name, gain_name)


def make_critic(state_dim, n_actions, h1, h2, name, gain_name="relu"):
    sa = state_dim + n_actions
    return ParamNet([(sa, h1, h2, 1, Q1_NAMES), (sa, h1, h2, 1, Q2_NAMES)], name, gain_name)

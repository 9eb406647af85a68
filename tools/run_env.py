"""Minimal driver for profilers: one env_step pass over 10 M synthetic pursuit states (3 launches)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hirl4ucav_b200 import synth  # noqa: E402
from hirl4ucav_b200.environments import harfang_math as hm  # noqa: E402

raw = torch.from_numpy(synth.raw_states(np.random.default_rng(0), 10_000_000)).cuda()
for _ in range(3):
    obs, rew, fl = hm.step_math(raw)
torch.cuda.synchronize()
print(float(rew[:8].sum()))

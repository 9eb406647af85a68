"""Minimal driver for profilers: builds the BASELINE soft256 agent and runs learn_many(12) a few times."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hirl4ucav_b200 import synth  # noqa: E402
from hirl4ucav_b200.agents.HIRL import Agent  # noqa: E402
from hirl4ucav_b200.utils.seed import set_seed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--bcw", type=float, default=100.0)
ap.add_argument("--expert", type=int, default=64)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
es, ea = synth.expert_table(20000)
set_seed(0)
ag = Agent(1e-3, 1e-3, 13, 4, 256, 512, 0.005, 0.99, 10 ** 5, a.batch, True, "Harfang_GYM", es, ea, 0.5, True, max_fused_steps=64)
ag.FUSE_SUB = 64   # one launch for all calls (profilers look at a single persistent launch)
synth.fill_agent(ag, 30000, expert_states=es, expert_actions=ea)
for _ in range(a.launches):
    out = ag.learn_many([(a.bcw, a.expert, 0)] * a.steps)
print([float(x) for x in out[-1]])

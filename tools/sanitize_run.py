"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): a few learn() calls of the
single-agent update kernel in the back-end chosen through the environment (HIRL_TC, HIRL_CLUSTER, HIRL_TCC), or of the
ensemble kernel (--ensemble).  Sizes are small because the sanitizer slows kernels down 10-100x; every code path of the
kernels still runs (critic-only calls, actor calls, the Polyak call)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from hirl4ucav_b200 import synth  # noqa: E402
from hirl4ucav_b200.utils.seed import set_seed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--ensemble", type=int, default=0, help="replicas (0: single agent)")
ap.add_argument("--dp", action="store_true", help="data-parallel mode with one rank (stage-range launches + NCCL all-reduces)")
a = ap.parse_args()
es, ea = synth.expert_table(1500)
set_seed(0)
if a.ensemble:
    from hirl4ucav_b200.ensemble import AgentEnsemble
    from hirl4ucav_b200.environments import harfang_math as hm
    ens = AgentEnsemble(list(range(a.ensemble)), 1e-3, 1e-3, 0.005, 0.99, 4000, a.batch, es, ea, max_fused_steps=8)
    erows, _ = hm.expert_scan(es, ea)
    rng = np.random.default_rng(0)
    for r in range(a.ensemble):
        s, ac, s2 = synth.env_states(rng, 2000), synth.env_actions(rng, 2000), synth.env_states(rng, 2000)
        rew, _, done = hm.offline_reward(s, ac, s2)
        rows = np.empty((2000, 32), np.float32)
        rows[:, 0:13], rows[:, 13:17], rows[:, 17:30], rows[:, 30], rows[:, 31] = s, ac, s2, rew, done
        ens.store_rows(r, rows)
        ens.store_rows(r, erows, expert=True)
    ens.reseed(1)
    out = ens.learn_many([(100, a.batch // 4, 0)] * a.steps)
    print("ensemble ok", out[:, -1, :4].tolist())
else:
    from hirl4ucav_b200.agents.HIRL import Agent
    ag = Agent(1e-3, 1e-3, 13, 4, 256, 512, 0.005, 0.99, 10 ** 4, a.batch, True, "Harfang_GYM", es, ea, 0.5, True, max_fused_steps=8,
               **({"data_parallel": (0, 1)} if a.dp else {}))
    synth.fill_agent(ag, 3000, expert_states=es, expert_actions=ea)
    out = ag.learn_many([(100, a.batch // 4, 0)] * a.steps)
    print("single ok", [float(x) for x in out[-1]])

"""First-light check of the tensor-core cluster kernel: one teacher-forced call at a time against the oracle, verbose."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HIRL_TCC"] = "1"
import numpy as np, torch
from hirl4ucav_b200 import _lib
from hirl4ucav_b200.agents.HIRL import Agent
from oracle import scenarios
from oracle.hirl_oracle import OracleAgent

def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
bcw = float(sys.argv[3]) if len(sys.argv) > 3 else 100
en = int(sys.argv[4]) if len(sys.argv) > 4 else B // 4
sc = scenarios.make_scenario(max(1000, 4 * B), max(600, 3 * B))
orc = scenarios.build_agent(OracleAgent, sc, B, 0)
ag = scenarios.build_agent(Agent, sc, B, 0)
scenarios.seed_all(1)
names = [n for n, _ in orc.critic.named_parameters()]
for t in range(steps):
    r32 = {}
    actor_step = orc.actorTrainable
    out32 = orc.learn(bcw, en, 0, rec=r32)
    d = orc.last_draws
    out = ag.learn(bcw, en, 0, forced={"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]})
    print(f"step {t} actor={actor_step}\n  got {[float(x) for x in out]}\n  ref {[float(x) for x in out32]}")
    g = ag.export_grads(_lib.NET_CRITIC)
    off = 0
    for n, p in orc.critic.named_parameters():
        ref = r32["critic.grads"][n].double().reshape(-1).numpy()
        print(f"    critic {n:22s} rel {rel(g[off:off + ref.size], ref):.2e}")
        off += ref.size
    if actor_step:
        g = ag.export_grads(_lib.NET_ACTOR)
        off = 0
        for n, p in orc.actor.named_parameters():
            ref = r32["actor.grads"][n].double().reshape(-1).numpy()
            print(f"    actor  {n:22s} rel {rel(g[off:off + ref.size], ref):.2e}")
            off += ref.size
    for w, nid in (("actor", 0), ("targetActor", 1), ("critic", 2), ("targetCritic", 3)):
        print(f"    weights {w:13s} rel {rel(ag.export_flat(nid), orc.flat(w)):.2e}")
    # teacher forcing
    for w in ("actor", "targetActor", "critic", "targetCritic"):
        getattr(ag, w).load_state_dict(orc.nets()[w].state_dict())
    ag.push_parameters()
    for w, nid in (("actor", _lib.NET_ACTOR), ("critic", _lib.NET_CRITIC)):
        if orc.nets()[w].optimizer.state:
            m, v = orc.adam_state(w)
            ag.import_adam(nid, m, v, ag._actor_t if w == "actor" else ag._critic_t)

import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from hirl4ucav_b200 import _lib
from hirl4ucav_b200.agents.HIRL import Agent
from oracle import scenarios
from oracle.hirl_oracle import OracleAgent
B=32
sc = scenarios.make_scenario(1000, 600)
orc = scenarios.build_agent(OracleAgent, sc, B, 0)
ag = scenarios.build_agent(Agent, sc, B, 0)
scenarios.seed_all(1)
rec={}
orc.learn(100, 8, 0, rec=rec)
d = orc.last_draws
ag.learn(100, 8, 0, forced={"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]})
gw = rec["actor.grads"]["full2.weight"].numpy().ravel()
parts = ag.debug_buffer(18).reshape(2, -1)
g = ag.export_grads(_lib.NET_ACTOR)
off = 13*256 + 256*3
gk = g[off:off+131072]
n = np.linalg.norm
print("oracle", n(gw), "part0", n(parts[0]), "part1", n(parts[1]), "sum-oracle", n(parts[0]+parts[1]-gw), "exported-oracle", n(gk-gw), "exported-(p0+p1)", n(gk-parts[0]-parts[1]))
bad = np.nonzero(np.abs(parts[0]+parts[1]-gw) > 1e-3*np.abs(gw).max())[0]
print("bad count", bad.size, "rows(n)", np.unique(bad//256)[:20], "cols(k)", np.unique(bad%256)[:20])
idx = [0, 1, 2, 255, 256, 257, 1000, 65536, 131071]
for i in idx:
    print(i, "oracle", gw[i], "p0", parts[0][i], "p1", parts[1][i], "exported", gk[i])
# is the exported array a permutation / transposition?
G = gk.reshape(512, 256); O = gw.reshape(512, 256)
print("exported == oracle transposed-view?", np.linalg.norm(gk.reshape(256, 512).T.ravel() - gw))
print("corr", np.dot(gk, gw) / (n(gk) * n(gw)))
cm, cv = ag.export_adam(_lib.NET_ACTOR)
mk = cm[off:off+131072]
print("adam m / (0.1*oracle g): ", (mk[:5] / (0.1 * gw[:5])))
badm = np.abs(gk - gw) > 1e-3 * np.abs(gw).max()
print("bad elements", badm.sum(), "of", badm.size)
B2 = badm.reshape(512, 256)
print("bad per k (nonzero):", {int(k): int(v) for k, v in enumerate(B2.sum(0)) if v})
print("bad per n (first 10 nonzero):", [(int(nn), int(v)) for nn, v in enumerate(B2.sum(1)) if v][:10])
print("adam m ok where g bad?", np.abs(mk[badm] - 0.1 * gw[badm]).max(), np.abs(mk[badm] - 0.1 * gk[badm]).max())

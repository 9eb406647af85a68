"""GPU box: where the host time of Agent.learn() goes (per-call wall clock, cProfile top entries)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hirl4ucav_b200 import synth  # noqa: E402
from hirl4ucav_b200.agents.HIRL import Agent  # noqa: E402
from hirl4ucav_b200.utils.seed import set_seed  # noqa: E402

es, ea = synth.expert_table(20000)
set_seed(0)
ag = Agent(1e-3, 1e-3, 13, 4, 256, 512, 0.005, 0.99, 10 ** 5, 256, True, "Harfang_GYM", es, ea, 0.5, True)
synth.fill_agent(ag, 30000, expert_states=es, expert_actions=ea)
for _ in range(50):
    ag.learn(100, 64, 0)
t0 = time.perf_counter()
N = 1000
for _ in range(N):
    ag.learn(100, 64, 0)
dt = (time.perf_counter() - t0) / N * 1e6
print(f"learn(): {dt:.1f} us/call; predraw hits py {ag._pre_py.hits} np {ag._pre_np.hits} misses {ag._pre_py.misses} {ag._pre_np.misses}")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    ag.learn(100, 64, 0)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)

# ---- the fused form: learn_many(60) ------------------------------------------------------------------------------------
chunk = [(100, 64, 0)] * 60
ag.learn_many(chunk)
t0 = time.perf_counter()
for _ in range(20):
    ag.learn_many(chunk)
dt = (time.perf_counter() - t0) / (20 * 60) * 1e6
print(f"learn_many(60): {dt:.1f} us/call")
for sub in (4, 16, 32):
    ag.FUSE_SUB = sub
    ag.learn_many(chunk)
    t0 = time.perf_counter()
    for _ in range(20):
        ag.learn_many(chunk)
    print(f"  FUSE_SUB={sub}: {(time.perf_counter() - t0) / (20 * 60) * 1e6:.1f} us/call")
ag.FUSE_SUB = 8
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    ag.learn_many(chunk)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

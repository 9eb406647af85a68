"""GPU: the ensemble kernel (BASELINE config 4): every replica equals a single agent run with the
same seeds; results do not depend on how many replicas share the launch; runs are deterministic."""
import numpy as np
import pytest

from hirl4ucav_b200 import _lib, synth
from hirl4ucav_b200.agents.HIRL import Agent
from hirl4ucav_b200.ensemble import AgentEnsemble
from hirl4ucav_b200.environments import harfang_math as hm
from hirl4ucav_b200.utils.seed import set_seed

pytestmark = pytest.mark.gpu
B, N_ON, N_EX = 64, 1500, 900


def rows_for(seed):
    rng = np.random.default_rng(1000 + seed)
    s, a, s2 = synth.env_states(rng, N_ON), synth.env_actions(rng, N_ON), synth.env_states(rng, N_ON)
    r, succ, done = hm.offline_reward(s, a, s2)
    rows = np.empty((N_ON, 32), np.float32)
    rows[:, 0:13], rows[:, 13:17], rows[:, 17:30], rows[:, 30], rows[:, 31] = s, a, s2, r, done
    return rows


def make_ensemble(seeds, es, ea, erows):
    ens = AgentEnsemble(seeds, 1e-3, 1e-3, 0.005, 0.99, 4000, B, es, ea, max_fused_steps=16)
    for r, seed in enumerate(seeds):
        ens.store_rows(r, rows_for(seed))
        ens.store_rows(r, erows, expert=True)
    ens.reseed(1)
    return ens


def rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_replicas_match_single_agents_and_are_independent_of_the_launch_shape():
    es, ea = synth.expert_table(N_EX)
    erows, _ = hm.expert_scan(es, ea)
    calls = [(100, 16, 0)] * 12
    seeds = [3, 4, 5, 6, 7]
    ens = make_ensemble(seeds, es, ea, erows)
    out = ens.learn_many(calls)
    # (a) each replica vs the single-agent class under the same seeds (different team size ->
    #     different partial-sum grouping: equal to fp32 round-off, not bit-identical)
    for r in (0, 3):
        set_seed(seeds[r])
        ag = Agent(1e-3, 1e-3, 13, 4, 256, 512, 0.005, 0.99, 4000, B, True, "Harfang_GYM", es, ea, 0.5, True, max_fused_steps=16)
        ag.buffer.store_rows(rows_for(seeds[r]))
        ag.expert_buffer.store_rows(erows)
        set_seed(seeds[r] + 1)
        single = ag.learn_many(calls)
        for k in range(12):
            for j in range(4):
                assert abs(out[r, k, j] - float(single[k][j])) <= 2e-4 * abs(float(single[k][j])) + 1e-5
            assert abs(out[r, k, 6] / B - single[k][5]) <= 1.0 / B + 1e-9 or calls[k][0] != 100 or single[k][5] == 1
        for nid in range(4):
            assert rel(ens.export_flat(r, nid), ag.export_flat(nid)) < 2e-4
    # (b) a different launch shape (2 replicas instead of 5, so other team assignment) gives the
    #     SAME bits per replica, and so does a repeated run: deterministic and replica-independent
    ens2 = make_ensemble(seeds[:2], es, ea, erows)
    out2 = ens2.learn_many(calls)
    assert np.array_equal(out2, out[:2])
    for r in range(2):
        for nid in range(4):
            assert np.array_equal(ens2.export_flat(r, nid), ens.export_flat(r, nid))
    # replicas really are different agents
    assert not np.array_equal(ens.export_flat(0, 0), ens.export_flat(1, 0))

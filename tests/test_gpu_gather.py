"""GPU: replay index streams + gather, bit-exact against the reference fixture and the oracle."""
import os

import numpy as np
import pytest
import torch

from hirl4ucav_b200.agents.HIRL import Agent
from oracle import scenarios
from oracle.hirl_oracle import OracleAgent

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["soft32", "soft256"])
def test_gather_matches_reference_batch(golden_dir, name):
    fx = np.load(os.path.join(golden_dir, f"agent_{name}.npz"))
    B = int(fx["batch"])
    sc = scenarios.make_scenario(int(fx["n_online"]), int(fx["n_expert"]))
    ag = scenarios.build_agent(Agent, sc, B, int(fx["seed"]))
    rows = ag.gather(fx["idx_online"][0], fx["idx_expert"][0])
    assert np.array_equal(rows[:, 0:13], fx["batch0.state"])
    assert np.array_equal(rows[:, 13:17], fx["batch0.action"])
    assert np.array_equal(rows[:, 17:30], fx["batch0.next_state"])
    assert np.array_equal(rows[:, 30], fx["batch0.reward"])
    assert np.array_equal(rows[:, 31], fx["batch0.done"])
    s, a = ag.gather_bc(fx["idx_bc"][0])
    assert np.array_equal(s, sc["expert_states"][fx["idx_bc"][0]].astype(np.float32))
    assert np.array_equal(a, sc["expert_actions"][fx["idx_bc"][0]].astype(np.float32))


def test_index_streams_equal_reference(golden_dir):
    """Same seeds -> the drop-in draws the very indices/noise the reference drew."""
    fx = np.load(os.path.join(golden_dir, "agent_soft32.npz"))
    B = int(fx["batch"])
    sc = scenarios.make_scenario(int(fx["n_online"]), int(fx["n_expert"]))
    ag = scenarios.build_agent(Agent, sc, B, int(fx["seed"]))
    scenarios.seed_all(int(fx["run_seed"]))
    for t in range(int(fx["steps"])):
        ag.learn(100, int(fx["expert_num_now"]), 0)
        d = ag.last_draws
        assert np.array_equal(d["online"], fx["idx_online"][t])
        assert np.array_equal(d["expert"], fx["idx_expert"][t])
        assert np.array_equal(d["bc"], fx["idx_bc"][t])
        assert np.array_equal(d["noise"], fx["noise"][t])


def test_ring_wrap_and_edges():
    sc = scenarios.make_scenario(300, 200)
    ag = scenarios.build_agent(Agent, sc, 16, 0, hp={"bufferSize": 256})   # 300 stores into 256 slots: wraps
    orc = scenarios.build_agent(OracleAgent, sc, 16, 0, hp={"bufferSize": 256})
    assert len(ag.buffer) == len(orc.buffer) == 256 and ag.buffer.position == orc.buffer.position
    idx = list(range(0, 256, 7))
    rows = ag.gather(idx, [0, 5])
    for j, i in enumerate(idx):
        m = orc.buffer.memory[i]
        assert np.array_equal(rows[j, :13], np.asarray(m.state, np.float32))
        assert rows[j, 30] == np.float32(m.reward) and rows[j, 31] == float(m.done)
    assert ag.gather([], []).shape == (0, 32)
    with pytest.raises(ValueError):
        ag.gather([256], [])
    with pytest.raises(ValueError):      # random.sample: sample larger than population
        ag.buffer.sample_indices(1000)

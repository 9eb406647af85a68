"""CPU: host-side mirrors of the reference interface that need no GPU."""
import random

import numpy as np
import torch

from hirl4ucav_b200 import nets, synth
from hirl4ucav_b200.utils import data_processor
from hirl4ucav_b200.utils.buffer import UniformMemory
from oracle import scenarios
from oracle.hirl_oracle import OracleActor, OracleCritic, OracleMemory


def test_seeded_initialisation_equals_reference_order():
    """Host parameter containers consume the torch CPU generator exactly like the reference's
    Actor/Critic constructors (HIRL.py:25-48, 111-121): same seed -> same weights, same keys."""
    torch.manual_seed(0)
    a, c = nets.make_actor(13, 4, 256, 512, "A"), nets.make_critic(13, 4, 256, 512, "C")
    torch.manual_seed(0)
    oa, oc = OracleActor(1e-3, 13, 4, 256, 512, True, "A"), OracleCritic(1e-3, 13, 4, 256, 512, True, "C")
    assert list(a.state_dict()) == list(oa.state_dict()) and list(c.state_dict()) == list(oc.state_dict())
    for x, y in zip(a.parameters(), oa.parameters()):
        assert torch.equal(x, y)
    for x, y in zip(c.parameters(), oc.parameters()):
        assert torch.equal(x, y)
    assert a.flat().size == 138756 and c.flat().size == 276482
    v = a.flat().copy()
    a.load_flat(v[::-1].copy()); a.load_flat(v)
    assert np.array_equal(a.flat(), v)


def test_ring_semantics_and_index_stream_match_reference_buffer():
    sc = scenarios.make_scenario(300, 50)
    mine, ref = UniformMemory(256), OracleMemory(256)
    for i in range(300):
        args = (sc["s"][i], sc["a"][i], sc["s2"][i], float(sc["r"][i]), bool(sc["d"][i]), int(sc["succ"][i]))
        mine.store(*args); ref.store(*args)
    assert len(mine) == len(ref) == 256 and mine.position == ref.position == 300 % 256
    assert mine.fullEnough(256) and not mine.fullEnough(257)
    random.seed(5)
    idx = mine.sample_indices(40)
    random.seed(5)
    ref.sample(40)
    assert idx == ref.last_indices                      # random.sample stream (buffer.py:45)
    random.seed(5)
    st, ac, ns, rw, dn = mine.sample(40)
    for j, i in enumerate(ref.last_indices):
        m = ref.memory[i]
        assert np.array_equal(st[j], np.asarray(m.state, np.float32)) and rw[j] == float(np.float32(m.reward)) and dn[j] == m.done
    assert len(mine.memory) == 256 and mine.memory[3].step_success == ref.memory[3].step_success


def test_expert_csv_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    es, ea = synth.env_states(rng, 40), synth.env_actions(rng, 40, expert=True)
    p = str(tmp_path / "expert_data_ai.csv")
    data_processor.write_data(es, ea, p)
    s2, a2 = data_processor.read_data(p)
    assert s2.shape == (40, 13) and a2.shape == (40, 4)
    assert np.allclose(s2, es, rtol=0, atol=5e-9 * np.abs(es).max() + 1e-8)   # str(np.ndarray) keeps 8 significant digits
    assert list(data_processor.up_sample(ea)) == list(np.where(ea[:, 3] == 1)[0])

"""CPU: the oracle restatements against the fixtures frozen from the UNMODIFIED reference
(oracle/gen_golden.py).  This is what pins the oracle on machines without the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import env_oracle, scenarios
from oracle.hirl_oracle import OracleAgent, OracleBC

NETS = ("actor", "targetActor", "critic", "targetCritic", "bc_actor")


@pytest.mark.parametrize("name", ["soft32", "fixed32", "soft256"])
def test_agent_oracle_matches_reference_fixture(golden_dir, name):
    torch.set_num_threads(1)
    fx = np.load(os.path.join(golden_dir, f"agent_{name}.npz"))
    B, steps = int(fx["batch"]), int(fx["steps"])
    sc = scenarios.make_scenario(int(fx["n_online"]), int(fx["n_expert"]))
    orc = scenarios.build_agent(OracleAgent, sc, B, int(fx["seed"]))
    for w in NETS:
        assert orc.flat(w).sum() == fx[f"{w}.init_sum"], f"initial weights of {w} differ from the reference's"
    scenarios.seed_all(int(fx["run_seed"]))
    bcw, en = float(fx["bc_weight_now"]), int(fx["expert_num_now"])
    for t in range(steps):
        out = orc.learn(bcw if bcw != 100 else 100, en, 0)
        assert [float(x) for x in out] == list(fx["outs"][t]), f"step {t}"
        d = orc.last_draws
        assert np.array_equal(d["online"], fx["idx_online"][t])          # random.sample stream
        if en:
            assert np.array_equal(d["expert"], fx["idx_expert"][t])
        assert np.array_equal(d["bc"], fx["idx_bc"][t])                  # np.random.choice stream
        assert np.array_equal(d["noise"], fx["noise"][t])                # torch.normal stream
    for w in NETS:
        a = orc.flat(w)
        assert np.array_equal(a[fx[f"{w}.probe_idx"]].astype(np.float32), fx[f"{w}.probe"])
        assert a.sum() == fx[f"{w}.sum"] and (a * a).sum() == fx[f"{w}.sumsq"]


def _upsample_agent(cls, mem_cls, sc, B, **kw):
    hp = scenarios.HP
    scenarios.seed_all(0)
    ag = cls(hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], hp["bufferSize"], B, True,
             hp["name"], sc["expert_states"], sc["expert_actions"], hp["bc_weight"], hp["expert_warm_up"], **kw)
    if mem_cls is not None:   # the oracle mirrors the reference: switched after construction (see gen_golden.upsample_case)
        ag.expert_upsample = True
        ag.buffer = mem_cls(hp["bufferSize"], True)
    scenarios.fill_agent(ag, sc)
    return ag


def test_upsampling_oracle_matches_reference_fixture(golden_dir):
    """buffer.py:21-29,39-43 + HIRL.py:253-257 switched on: weighted numpy draws with replacement."""
    from oracle.hirl_oracle import OracleMemory
    torch.set_num_threads(1)
    fx = np.load(os.path.join(golden_dir, "agent_upsample32.npz"))
    B, steps, en = int(fx["batch"]), int(fx["steps"]), int(fx["expert_num_now"])
    sc = scenarios.make_scenario(int(fx["n_online"]), int(fx["n_expert"]))
    orc = _upsample_agent(OracleAgent, OracleMemory, sc, B)
    scenarios.seed_all(1)
    for t in range(steps):
        assert [float(x) for x in orc.learn(100, en, 0)] == list(fx["outs"][t]), f"step {t}"
        d = orc.last_draws
        assert np.array_equal(d["online"], fx["idx_online"][t]) and np.array_equal(d["expert"], fx["idx_expert"][t])
        assert np.array_equal(d["bc"], fx["idx_bc"][t]) and np.array_equal(d["noise"], fx["noise"][t])
    assert any(len(set(r)) < len(r) for r in fx["idx_online"]), "fixture should contain repeated rows"
    for w in NETS:
        assert orc.flat(w).sum() == fx[f"{w}.sum"]


def test_td3_and_bc_oracles_match_reference_fixture(golden_dir):
    torch.set_num_threads(1)
    fx = np.load(os.path.join(golden_dir, "agent_td3_bc.npz"))
    sc = scenarios.make_scenario(2000, 800)
    hp = scenarios.HP
    scenarios.seed_all(0)
    orc = OracleAgent(hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], 10 ** 5, 64, True,
                      "Harfang_GYM", None, None, 0.0, False, mode="TD3")
    for i in range(2000):
        orc.store(sc["s"][i], sc["a"][i], sc["s2"][i], float(sc["r"][i]), bool(sc["d"][i]), int(sc["succ"][i]))
    scenarios.seed_all(1)
    for t in range(8):
        assert [float(x) for x in orc.learn()] == list(fx["td3.outs"][t])
    assert orc.flat("critic").sum() == fx["td3.critic.sum"]
    scenarios.seed_all(0)
    obc = OracleBC(1e-3, 13, 4, 256, 512, True, "Harfang_GYM", 64, sc["expert_states"], sc["expert_actions"])
    scenarios.seed_all(1)
    assert [float(obc.train_actor()) for _ in range(6)] == list(fx["bc.losses"])


def test_env_oracle_matches_reference_fixture(golden_dir):
    fx = np.load(os.path.join(golden_dir, "env_online.npz"))
    obs, rew, fl = env_oracle.online_step(fx["raw"])
    assert np.array_equal(fl, fx["flags"])
    assert np.array_equal(obs, fx["obs"])
    assert np.array_equal(rew, fx["reward"])
    if env_oracle.c_lib() is not None:
        o2, r2, f2 = env_oracle.online_step_c(fx["raw"])
        assert np.array_equal(f2, fx["flags"]) and np.array_equal(o2, fx["obs"]) and np.array_equal(r2, fx["reward"])
    fo = np.load(os.path.join(golden_dir, "env_offline.npz"))
    r, su = env_oracle.offline_reward(fo["s"], fo["a"], fo["s2"])
    assert np.array_equal(r, fo["reward"]) and np.array_equal(su, fo["success"])
    assert np.array_equal(env_oracle.offline_done(fo["s2"]), fo["done"])
    assert np.array_equal(env_oracle.offline_loc_diff(fo["s2"]), fo["loc"])


def test_env_threshold_edges_in_double():
    """x <= 0.1 on the widened fp32 value: 0.1f is ABOVE 0.1."""
    raw = np.zeros((4, 16), np.float32)
    raw[:, 1] = 3000.0
    raw[:, 13] = [np.float32(0.1), np.nextafter(np.float32(0.1), np.float32(0)), 0.0, -0.0]
    raw[:, 15] = np.array([16, 16, 16, 16], np.uint32).view(np.float32)   # fire_success sticky
    _, rew, fl = env_oracle.online_step(raw)
    d = env_oracle.decode_flags(fl)
    assert list(rew > 500) == [False, True, True, True]
    assert list(d["episode_success"]) == [False, False, True, True]


def test_expert_scan_matches_loop_semantics():
    sc = scenarios.make_scenario(10, 300)
    es, ea = sc["expert_states"], sc["expert_actions"]
    # straight transcription of the walk in train_all.py:287-306
    i, use = 0, []
    while i < es.shape[0] - 1:
        use.append(i)
        done = es[i + 1][-1] <= 0.1
        if done:
            i += 1
        i += 1
    assert use == list(sc["ex_use"])

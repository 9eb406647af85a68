"""GPU: the drop-in claim end to end.  The reference's own HIRL training loop (hirl/train_all.py:266-367: random
exploration -> expert-buffer construction -> chooseAction / env.step / store / learn, with the weight learn() returns fed
back in as the next call's bc_weight_now) is transcribed ONCE below with `agent` and `env` as parameters and run twice:

  * oracle side : oracle.hirl_oracle.OracleAgent (bit-identical to the unmodified reference agent, tests/test_oracle_golden.py)
                  + the numpy-f64 env restatement (bit-identical to the unmodified env classes on 120 000 states)
  * product side: the classes INTEGRATION.md's five import lines bring in (hirl4ucav_b200.agents.HIRL.Agent,
                  utils.seed.set_seed, utils.data_processor.read_data through a CSV round trip) + the CUDA env math

The external simulator is replaced on both sides by the same pre-generated stream of raw pursuit states (BASELINE.json:
"synthetic state vectors of the env's shape"); only the fire channel of a state depends on the action (GYM.py:150).
Bars: every host RNG draw identical (replay / expert / cloning indices, TD3 noise, exploration noise), actions within
2e-6, stored rewards / flags identical, the six values learn() returns within the free-run tolerance of
tests/test_gpu_update.py (1e-3 relative after ~70 coupled free-running updates at batch 32), bc_weight within one count of 1/B.
"""
import numpy as np
import pytest
import torch

from hirl4ucav_b200 import synth
from hirl4ucav_b200.agents.HIRL import Agent as HIRLAgent            # INTEGRATION.md section 2, line 1
from hirl4ucav_b200.environments import harfang_math as hm
from hirl4ucav_b200.utils.data_processor import read_data, write_data   # line 5
from hirl4ucav_b200.utils.seed import set_seed                       # line 3
from oracle import env_oracle, scenarios
from oracle.hirl_oracle import OracleAgent

pytestmark = pytest.mark.gpu


class _ActionSpace:
    def sample(self):
        return np.random.uniform(-1.0, 1.0, 4).astype(np.float32)     # gym.spaces.Box(-1, 1, (4,)).sample() stand-in (same on both sides)


class SyntheticHarfang:
    """reset() / step() / get_reward() / get_termination() of hirl/environments/HarfangEnv_GYM.py over a fixed stream of raw
    simulator states; `backend` supplies the arithmetic (oracle or CUDA)."""

    def __init__(self, raw_stream, backend):
        self.raw, self.t, self.backend = raw_stream, 0, backend
        self.action_space = _ActionSpace()

    def _observe(self, fire):
        raw = self.raw[self.t % len(self.raw)].copy()
        raw[14] = fire
        self.t += 1
        obs, rew, fl = self.backend["step"](raw[None])
        d = self.backend["decode"](fl)
        return np.asarray(obs[0], np.float64), float(rew[0]), bool(d["done"][0]), int(d["step_success"][0])

    def reset(self):
        return self._observe(0.0)[0]

    def step(self, action):
        obs, rew, done, succ = self._observe(float(action[3]))
        return obs, rew, done, {}, succ

    def get_reward(self, s, a, s2):
        return self.backend["get_reward"](s, a, s2)

    def get_termination(self, s2):
        return self.backend["get_termination"](s2)


def hirl_loop(agent, env, expert_states, expert_actions, batchSize, explorationEpisodes=2, trainingEpisodes=2, maxStep=40,
              warm_up_rate=10, bc_warm_up=False):
    """hirl/train_all.py:266-367 (--agent HIRL --type soft), logging / validation / plotting removed, nothing else changed."""
    log = {"learn": [], "actions": [], "stored": []}
    for episode in range(explorationEpisodes):                                    # :266-282
        state = env.reset()
        done = False
        for step in range(maxStep):
            if not done:
                action = env.action_space.sample()
                n_state, reward, done, info, step_success = env.step(action)
                if step == maxStep - 1:
                    break
                agent.store(state, action, n_state, reward, done, step_success)
                log["stored"].append((reward, done, step_success))
                state = n_state
    expert_num = batchSize                                                          # :286-308
    i = 0
    while i < expert_states.shape[0] - 1:
        state, action, next_state = expert_states[i], expert_actions[i], expert_states[i + 1]
        reward, step_success = env.get_reward(state, action, next_state)
        done = env.get_termination(next_state)
        if done:
            i += 1
        i += 1
        agent.expert_buffer.store(state, action, next_state, reward, done, step_success)
    log["expert_rows"] = len(agent.expert_buffer.memory)
    bc_warm_up_weight = 0
    for episode in range(trainingEpisodes):                                        # :320-367
        state = env.reset()
        done = False
        bc_weight_now = 100                                                        # soft: the sentinel, first call of the episode only
        for step in range(maxStep):
            if not done:
                action = agent.chooseAction(state)
                n_state, reward, done, info, step_success = env.step(action)
                if step == maxStep - 1:
                    break
                agent.store(state, action, n_state, reward, done, step_success)
                log["actions"].append(np.asarray(action, np.float64))
                log["stored"].append((reward, done, step_success))
                state = n_state
                if step % warm_up_rate == 0 and expert_num != 0:
                    expert_num -= 1
                if agent.buffer.fullEnough(agent.batchSize):
                    critic_loss, actor_loss, bc_loss, rl_loss, bc_fire_loss, bc_weight_now = agent.learn(bc_weight_now, expert_num, bc_warm_up_weight)
                    log["learn"].append((float(critic_loss), float(actor_loss), float(bc_loss), float(rl_loss), float(bc_fire_loss),
                                         float(bc_weight_now), dict(agent.last_draws)))
    return log


def test_reference_training_loop_runs_unchanged_on_the_drop_in(tmp_path):
    B = 32
    rng = np.random.default_rng(42)
    raw = synth.raw_states(rng, 400)
    raw[:, 1] = rng.uniform(2500.0, 6500.0, 400).astype(np.float32)       # keep the ally inside the altitude band: episodes run on
    raw[:, 13] = 0.2                                                       # opponent alive
    raw[2 * 41 + 1 + 30, 1] = 400.0                                        # ... except one crash: training episode 0 ends at its step 30
    es, ea = synth.expert_table(300)
    write_data(es, ea, str(tmp_path / "expert.csv"))                     # the on-disk format a user brings (data_processor.py:5-17)
    es_csv, ea_csv = read_data(str(tmp_path / "expert.csv"))
    assert es_csv.shape == es.shape and ea_csv.dtype == np.float64

    o_env = SyntheticHarfang(raw, {"step": env_oracle.online_step, "decode": env_oracle.decode_flags,
                                   "get_reward": lambda s, a, s2: tuple(x[0] for x in env_oracle.offline_reward(s, a, s2)),
                                   "get_termination": lambda s2: bool(env_oracle.offline_done(s2)[0])})
    math = hm.HarfangEnvMath()
    p_env = SyntheticHarfang(raw, {"step": hm.step_math, "decode": hm.decode_flags, "get_reward": math.get_reward,
                                   "get_termination": math.get_termination})
    hp = scenarios.HP
    args = (hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], 10 ** 5, B, True, "Harfang_GYM", es_csv, ea_csv, 0.5, True)
    scenarios.seed_all(3)
    orc = OracleAgent(*args)
    np.random.seed(3)
    o_log = hirl_loop(orc, o_env, es_csv, ea_csv, B)
    set_seed(3)
    ag = HIRLAgent(*args)
    np.random.seed(3)
    p_log = hirl_loop(ag, p_env, es_csv, ea_csv, B)

    assert p_log["expert_rows"] == o_log["expert_rows"] > 200
    assert len(p_log["learn"]) == len(o_log["learn"]) >= 50
    assert [s[1:] for s in p_log["stored"]] == [s[1:] for s in o_log["stored"]]                    # done / step_success flags identical
    assert sum(s[1] for s in o_log["stored"]) == 1                                                  # the crash ended one episode early
    assert np.allclose([s[0] for s in p_log["stored"]], [s[0] for s in o_log["stored"]], rtol=1e-6, atol=1e-6)
    assert np.abs(np.array(p_log["actions"]) - np.array(o_log["actions"])).max() <= 2e-4    # fp32 forward of slowly drifting weights
    fed_back = 0
    for k, (p, o) in enumerate(zip(p_log["learn"], o_log["learn"])):
        for name in ("online", "expert", "bc", "noise"):
            assert np.array_equal(np.asarray(p[6][name]), np.asarray(o[6][name])), f"call {k}: {name} draws differ"
        for j in range(5):
            assert abs(p[j] - o[j]) <= 1e-3 * abs(o[j]) + 1e-5, f"call {k} value {j}: {p[j]} vs {o[j]}"
        assert abs(p[5] - o[5]) <= 1.0 / B + 1e-12
        fed_back += o[5] != 100
    assert fed_back == len(o_log["learn"])      # the returned Q-filter weight, not the sentinel, is what the caller passes back in
    print(f"drop-in loop: {len(p_log['learn'])} learn() calls, last = {p_log['learn'][-1][:6]}")

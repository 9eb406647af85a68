"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (the reference checkout is at /root/reference there and does
not exist on the GPU box):

    python -m oracle.gen_golden

For every case it (1) runs the unmodified reference and the oracle restatement side by
side on identical inputs and RNG streams, (2) asserts they agree BIT-FOR-BIT (every
returned scalar, every index drawn, every weight of every network, Adam moments), and
(3) freezes the reference's outputs as small fixtures.  That is what "pins" the oracle;
tests/test_oracle_golden.py re-checks the oracle against the fixtures on any machine.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hirl4ucav_b200 import synth  # noqa: E402
from oracle import env_oracle, ref_harness, scenarios  # noqa: E402
from oracle.hirl_oracle import OracleAgent, OracleBC  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NETS = ("actor", "targetActor", "critic", "targetCritic", "bc_actor")


def probe_indices(n: int, k: int = 1024) -> np.ndarray:
    return np.sort(np.random.default_rng(99).choice(n, size=min(k, n), replace=False))


def ref_flat(agent, which):
    net = getattr(agent, which)
    return np.concatenate([p.detach().double().reshape(-1).numpy() for p in net.parameters()])


def ref_adam(agent, which):
    net = getattr(agent, which)
    st = net.optimizer.state
    m = [st[p]["exp_avg"].double().reshape(-1).numpy() for p in net.parameters()]
    v = [st[p]["exp_avg_sq"].double().reshape(-1).numpy() for p in net.parameters()]
    return np.concatenate(m), np.concatenate(v)


def record_streams(fn):
    """Runs fn() while logging what it draws from random.sample / np.random.choice /
    torch.normal (the three host RNG streams of learn(): buffer.py:45, HIRL.py:249, :265)."""
    log = {"sample": [], "choice": [], "normal": []}
    o_sample, o_choice, o_normal = random.sample, np.random.choice, torch.normal

    def sample(pop, k, **kw):
        # identical stream whether the population is the list or range(len(list))
        idx = o_sample(range(len(pop)), k, **kw)
        log["sample"].append(np.asarray(idx, np.int64))
        return [pop[i] for i in idx]

    def choice(a, size=None, replace=True, p=None):
        out = o_choice(a, size, replace, p)
        log["choice"].append(np.asarray(out, np.int64))
        return out

    def normal(*a, **kw):
        out = o_normal(*a, **kw)
        log["normal"].append(out.detach().numpy().copy())
        return out

    random.sample, np.random.choice, torch.normal = sample, choice, normal
    try:
        res = fn()
    finally:
        random.sample, np.random.choice, torch.normal = o_sample, o_choice, o_normal
    return res, log


def agent_case(name, batch, steps, bcw, expert_num, n_online=3000, n_expert=1500, seed=0, run_seed=1):
    print(f"[agent] {name}: B={batch} steps={steps} bc_weight_now={bcw} expert_num_now={expert_num}")
    sc = scenarios.make_scenario(n_online, n_expert)
    Ref = ref_harness.hirl_agent_cls()
    with ref_harness.quiet():
        ref = scenarios.build_agent(Ref, sc, batch, seed)
    orc = scenarios.build_agent(OracleAgent, sc, batch, seed)
    for w in NETS:
        assert np.array_equal(ref_flat(ref, w), orc.flat(w)), f"init mismatch {w}"
    init_sums = {w: float(ref_flat(ref, w).sum()) for w in NETS}

    outs, draws = [], []
    scenarios.seed_all(run_seed)
    st_ref = (random.getstate(), np.random.get_state(), torch.get_rng_state())
    for t in range(steps):
        res, log = record_streams(lambda: ref.learn(bcw, expert_num, 0))
        outs.append([float(x) for x in res])
        draws.append(log)
    random.setstate(st_ref[0]); np.random.set_state(st_ref[1]); torch.set_rng_state(st_ref[2])
    for t in range(steps):
        res = orc.learn(bcw, expert_num, 0)
        assert [float(x) for x in res] == outs[t], f"step {t}: {res} vs {outs[t]}"
        d = orc.last_draws
        on = draws[t]["sample"][0]
        assert np.array_equal(d["online"], on)
        if expert_num:
            assert np.array_equal(d["expert"], draws[t]["sample"][1])
        assert np.array_equal(d["bc"], draws[t]["choice"][0])
        assert np.array_equal(d["noise"], draws[t]["normal"][0])
    fix = {"outs": np.asarray(outs, np.float64), "batch": batch, "steps": steps, "bc_weight_now": float(bcw),
           "expert_num_now": expert_num, "n_online": n_online, "n_expert": n_expert, "seed": seed, "run_seed": run_seed}
    for w in NETS:
        a, b = ref_flat(ref, w), orc.flat(w)
        assert np.array_equal(a, b), f"final weights differ: {w}"
        pi = probe_indices(a.size)
        fix[f"{w}.probe_idx"], fix[f"{w}.probe"] = pi, a[pi].astype(np.float32)
        fix[f"{w}.sum"], fix[f"{w}.sumsq"], fix[f"{w}.init_sum"] = a.sum(), (a * a).sum(), init_sums[w]
    for w in ("actor", "critic"):
        (m, v), (m2, v2) = ref_adam(ref, w), orc.adam_state(w)
        assert np.array_equal(m, m2) and np.array_equal(v, v2), f"adam state differs: {w}"
        pi = probe_indices(m.size)
        fix[f"{w}.m_probe"], fix[f"{w}.v_probe"] = m[pi].astype(np.float32), v[pi].astype(np.float32)
    fix["idx_online"] = np.stack([d["sample"][0] for d in draws]).astype(np.int32)
    if expert_num:
        fix["idx_expert"] = np.stack([d["sample"][1] for d in draws]).astype(np.int32)
    fix["idx_bc"] = np.stack([d["choice"][0] for d in draws]).astype(np.int32)
    fix["noise"] = np.stack([d["normal"][0] for d in draws]).astype(np.float32)
    # one gathered batch (step 0) as the reference's own float32 tensors would hold it
    on, ex = fix["idx_online"][0], (fix["idx_expert"][0] if expert_num else np.zeros(0, np.int32))
    rows = [ref.buffer.memory[i] for i in on] + [ref.expert_buffer.memory[i] for i in ex]
    fix["batch0.state"] = torch.tensor(np.array([r.state for r in rows]), dtype=torch.float).numpy()
    fix["batch0.action"] = torch.tensor(np.array([r.action for r in rows]), dtype=torch.float).numpy()
    fix["batch0.next_state"] = torch.tensor(np.array([r.next_state for r in rows]), dtype=torch.float).numpy()
    fix["batch0.reward"] = torch.tensor(np.array([r.reward for r in rows]), dtype=torch.float).numpy()
    fix["batch0.done"] = torch.tensor(np.array([r.done for r in rows]), dtype=torch.float).numpy()
    # acting: three public selectors on a fixed state (HIRL.py:192-212)
    scenarios.seed_all(7)
    s0 = sc["s"][5]
    acts_ref = [ref.chooseAction(s0), ref.chooseActionSmallNoise(s0), ref.chooseActionNoNoise(s0)]
    scenarios.seed_all(7)
    acts_orc = [orc.chooseAction(s0), orc.chooseActionSmallNoise(s0), orc.chooseActionNoNoise(s0)]
    for x, y in zip(acts_ref, acts_orc):
        assert np.array_equal(x, y)
    fix["act_state"], fix["acts"] = s0, np.stack(acts_ref)
    np.savez_compressed(os.path.join(OUT, f"agent_{name}.npz"), **fix)
    print(f"    ok: reference == oracle bit-for-bit over {steps} steps; last out = {outs[-1]}")


def td3_bc_case():
    print("[agent] td3 / bc variants")
    sc = scenarios.make_scenario(2000, 800)
    Ref = ref_harness.td3_agent_cls()
    scenarios.seed_all(0)
    hp = scenarios.HP
    with ref_harness.quiet():
        ref = Ref(hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], 10 ** 5, 64, True, "Harfang_GYM")
    scenarios.seed_all(0)
    orc = OracleAgent(hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], 10 ** 5, 64, True,
                      "Harfang_GYM", None, None, 0.0, False, mode="TD3")
    for ag in (ref, orc):
        for i in range(2000):
            ag.store(sc["s"][i], sc["a"][i], sc["s2"][i], float(sc["r"][i]), bool(sc["d"][i]), int(sc["succ"][i]))
    outs = []
    scenarios.seed_all(1)
    for t in range(8):
        outs.append([float(x) for x in ref.learn()])
    scenarios.seed_all(1)
    for t in range(8):
        assert [float(x) for x in orc.learn()] == outs[t]
    for w in ("actor", "targetActor", "critic", "targetCritic"):
        assert np.array_equal(ref_flat(ref, w), orc.flat(w))
    fix = {"td3.outs": np.asarray(outs)}
    for w in ("actor", "critic", "targetActor", "targetCritic"):
        a = ref_flat(ref, w); pi = probe_indices(a.size)
        fix[f"td3.{w}.probe"], fix[f"td3.{w}.sum"] = a[pi].astype(np.float32), a.sum()

    RefBC = ref_harness.bc_agent_cls()
    scenarios.seed_all(0)
    with ref_harness.quiet():
        rbc = RefBC(1e-3, 13, 4, 256, 512, True, "Harfang_GYM", 64, sc["expert_states"], sc["expert_actions"])
    scenarios.seed_all(0)
    obc = OracleBC(1e-3, 13, 4, 256, 512, True, "Harfang_GYM", 64, sc["expert_states"], sc["expert_actions"])
    scenarios.seed_all(1)
    bl = [float(rbc.train_actor()) for _ in range(6)]
    scenarios.seed_all(1)
    for t in range(6):
        assert float(obc.train_actor()) == bl[t]
    a = np.concatenate([p.detach().double().reshape(-1).numpy() for p in rbc.actor.parameters()])
    b = np.concatenate([p.detach().double().reshape(-1).numpy() for p in obc.actor.parameters()])
    assert np.array_equal(a, b)
    pi = probe_indices(a.size)
    fix["bc.losses"], fix["bc.actor.probe"], fix["bc.actor.sum"] = np.asarray(bl), a[pi].astype(np.float32), a.sum()
    np.savez_compressed(os.path.join(OUT, "agent_td3_bc.npz"), **fix)
    print("    ok: TD3 and BC reference == oracle bit-for-bit")


def upsample_case(batch=32, steps=8, expert_num=8):
    """Both up-sampling switches ON (buffer.py:21-29,39-43; HIRL.py:253-257).  The reference hard-codes them to
    False in its constructor (HIRL.py:184-185), so the unmodified class is constructed first and then switched the
    way a user editing those two lines would: expert_upsample attribute + a UniformMemory(capacity, True)."""
    print(f"[agent] upsample32: B={batch} steps={steps} with buffer_upsample and expert_upsample on")
    sc = scenarios.make_scenario(300, 1500)   # small ring: the weighted draw WITH replacement repeats rows
    Ref = ref_harness.hirl_agent_cls()
    RefMem = ref_harness.ref_memory_cls()

    def make(cls, mem_cls):
        scenarios.seed_all(0)
        hp = scenarios.HP
        ag = cls(hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], hp["bufferSize"], batch, True,
                 hp["name"], sc["expert_states"], sc["expert_actions"], hp["bc_weight"], hp["expert_warm_up"])
        ag.expert_upsample = True
        ag.buffer = mem_cls(hp["bufferSize"], True)
        scenarios.fill_agent(ag, sc)
        return ag
    from oracle.hirl_oracle import OracleMemory
    with ref_harness.quiet():
        ref = make(Ref, RefMem)
    orc = make(OracleAgent, OracleMemory)
    outs, on, ex, bc, noise = [], [], [], [], []
    scenarios.seed_all(1)
    for t in range(steps):
        outs.append([float(x) for x in ref.learn(100, expert_num, 0)])
    scenarios.seed_all(1)
    for t in range(steps):
        assert [float(x) for x in orc.learn(100, expert_num, 0)] == outs[t], f"step {t}"
        d = orc.last_draws
        on.append(d["online"]); ex.append(d["expert"]); bc.append(d["bc"]); noise.append(d["noise"])
    fix = {"outs": np.asarray(outs, np.float64), "batch": batch, "steps": steps, "expert_num_now": expert_num,
           "idx_online": np.asarray(on, np.int32), "idx_expert": np.asarray(ex, np.int32), "idx_bc": np.asarray(bc, np.int32),
           "noise": np.asarray(noise, np.float32), "n_online": 300, "n_expert": 1500}
    assert any(len(set(r)) < len(r) for r in on), "weighted draw with replacement should repeat rows in this scenario"
    for w in NETS:
        a, b = ref_flat(ref, w), orc.flat(w)
        assert np.array_equal(a, b), f"final weights differ: {w}"
        pi = probe_indices(a.size)
        fix[f"{w}.probe_idx"], fix[f"{w}.probe"], fix[f"{w}.sum"] = pi, a[pi].astype(np.float32), a.sum()
    np.savez_compressed(os.path.join(OUT, "agent_upsample32.npz"), **fix)
    print(f"    ok: reference == oracle bit-for-bit with up-sampling over {steps} steps; last out = {outs[-1]}")


def env_case(n_check=120_000, n_keep=4096):
    print(f"[env] online step on {n_check} raw states through the unmodified reference class")
    rng = np.random.default_rng(2024)
    raw = synth.raw_states(rng, n_check)
    # hand-made edge rows up front (kept in the fixture)
    edge = raw[:64].copy()
    hv = [0.0, synth.F32_TENTH, np.nextafter(np.float32(0.1), np.float32(0)), np.nextafter(np.float32(0.1), np.float32(1)),
          0.2, -0.0, 1e-30, 0.05]
    av = [499.9999, 500.0, 500.0001, 1999.9999, 2000.0, 2000.0001, 6999.9995, 7000.0, 7000.0005, 9999.999, 10000.0, 10000.001]
    for i in range(64):
        edge[i, 13] = hv[i % len(hv)]
        edge[i, 1] = av[i % len(av)]
        edge[i, 15:16].view(np.uint32)[0] = i % 32
        edge[i, 14] = [0.0, 1e-9, -1e-9, 0.5][i % 4]
    raw[:64] = edge
    for kind in ("HarfangEnv", "HarfangSerpentineInfiniteEnv"):
        obs_r, rew_r, fl_r = ref_harness.run_online_steps(raw, kind)
        obs_o, rew_o, fl_o = env_oracle.online_step(raw)
        assert np.array_equal(fl_r, fl_o), f"{kind}: flags differ at {np.nonzero(fl_r != fl_o)[0][:5]}"
        assert np.array_equal(obs_r, obs_o), f"{kind}: obs differ"
        assert np.array_equal(rew_r, rew_o), f"{kind}: reward differs: max {np.abs(rew_r - rew_o).max()}"
    print("    ok: obs / reward (f64) / flags identical for both env classes")
    sc = scenarios.make_scenario(20000, 4000)
    r_r, s_r, d_r, l_r = ref_harness.run_offline(sc["s"], sc["a"], sc["s2"])
    r_o, s_o = env_oracle.offline_reward(sc["s"], sc["a"], sc["s2"])
    assert np.array_equal(r_r, r_o) and np.array_equal(s_r, s_o)
    assert np.array_equal(d_r, env_oracle.offline_done(sc["s2"]))
    assert np.array_equal(l_r, env_oracle.offline_loc_diff(sc["s2"]))
    print("    ok: offline get_reward / get_termination / get_loc_diff identical on 20000 rows")
    k = n_keep
    np.savez_compressed(os.path.join(OUT, "env_online.npz"), raw=raw[:k], obs=obs_r[:k], reward=rew_r[:k], flags=fl_r[:k])
    np.savez_compressed(os.path.join(OUT, "env_offline.npz"), s=sc["s"][:k].astype(np.float32), a=sc["a"][:k].astype(np.float32),
                        s2=sc["s2"][:k].astype(np.float32), reward=r_r[:k], success=s_r[:k], done=d_r[:k], loc=l_r[:k])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # fixtures must not depend on the intra-op thread count
    only = set(sys.argv[1:])   # e.g. `python -m oracle.gen_golden upsample` regenerates one fixture
    run = lambda tag: not only or tag in only
    if run("env"):
        env_case()
    if run("soft32"):
        agent_case("soft32", 32, 12, 100, 8)
    if run("fixed32"):
        agent_case("fixed32", 32, 12, 0.5, 0)
    if run("soft256"):
        agent_case("soft256", 256, 12, 100, 64, n_online=4000, n_expert=2000)
    if run("td3_bc"):
        td3_bc_case()
    if run("upsample"):
        upsample_case()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()

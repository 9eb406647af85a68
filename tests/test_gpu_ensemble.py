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


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 4 at its real shape: batch 256 selects the tcgen05 path of hirl_ensemble_kernel (8-CTA teams, every
# contraction 3xTF32).  Teacher-forced against the ORACLE (bit-identical to the unmodified reference), per replica,
# with the same bars as tests/test_gpu_update.py::test_teacher_forced_parity.
# ---------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _check(tag, got, ref32, ref64, tol=1e-5):
    e_c32, e_c64, e_o64 = _rel(got, ref32), _rel(got, ref64), _rel(ref32, ref64)
    if e_c32 <= tol:
        return e_c32
    assert e_o64 > 3e-6, f"{tag}: rel-L2 {e_c32:.2e} vs torch fp32 (> {tol}) on an ordinary step"
    assert e_c64 <= 4 * e_o64 + 2e-6, f"{tag}: outside the fp32 envelope: cuda-f64 {e_c64:.2e} vs torch32-f64 {e_o64:.2e}"
    return e_c32


def test_ensemble_tensor_core_path_at_batch_256_matches_oracle_per_replica():
    import ctypes as C

    import torch

    from oracle import scenarios
    from oracle.hirl_oracle import OracleAgent

    B256, EN, STEPS = 256, 64, 6
    seeds = [11, 12, 13, 14, 15]
    scs = [scenarios.make_scenario(1200, 700, data_seed=500 + s) for s in seeds]
    es, ea = scs[0]["expert_states"], scs[0]["expert_actions"]          # the expert (s, a) table is shared by the replicas
    for sc in scs[1:]:
        sc.update({k: scs[0][k] for k in ("expert_states", "expert_actions", "ex_use", "ex_r", "ex_d", "ex_succ")})
    orcs = [scenarios.build_agent(OracleAgent, sc, B256, s) for sc, s in zip(scs, seeds)]
    o64s = [scenarios.build_agent(OracleAgent, sc, B256, s, dtype=torch.float64) for sc, s in zip(scs, seeds)]
    ens = AgentEnsemble(seeds, 1e-3, 1e-3, 0.005, 0.99, 4000, B256, es, ea, max_fused_steps=4)
    lib, h = ens._lib, ens._h
    for r, sc in enumerate(scs):
        n = sc["s"].shape[0]
        rows = np.empty((n, 32), np.float32)
        rows[:, 0:13], rows[:, 13:17], rows[:, 17:30], rows[:, 30], rows[:, 31] = sc["s"], sc["a"], sc["s2"], sc["r"], sc["d"]
        ens.store_rows(r, rows)
        use = np.asarray(sc["ex_use"])
        erows = np.empty((use.size, 32), np.float32)
        erows[:, 0:13], erows[:, 13:17], erows[:, 17:30] = es[use], ea[use], es[use + 1]
        erows[:, 30], erows[:, 31] = sc["ex_r"], sc["ex_d"]
        ens.store_rows(r, erows, expert=True)
        for w, nid in (("actor", 0), ("targetActor", 1), ("critic", 2), ("targetCritic", 3), ("bc_actor", 4)):
            assert np.array_equal(ens.export_flat(r, nid).astype(np.float64), orcs[r].flat(w)), f"replica {r}: seeded init differs: {w}"
    nets4 = (("actor", 0), ("targetActor", 1), ("critic", 2), ("targetCritic", 3))
    worst = 0.0
    for t in range(STEPS):
        ens.draw([(100, EN, 0)])                       # schedule fields of the step records; the draws are overwritten below
        steps = ens._steps_t.numpy().view(_lib.STEP_DTYPE)[: ens.R].reshape(ens.R, 1)
        recs = []
        for r in range(ens.R):
            scenarios.seed_all(100 * t + seeds[r])
            r32, r64 = {}, {}
            actor_step = orcs[r].actorTrainable
            out32 = orcs[r].learn(100, EN, 0, rec=r32)
            d = orcs[r].last_draws
            forced = {"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]}
            if actor_step:
                out64 = o64s[r].learn(100, EN, 0, forced=dict(forced), rec=r64, force_soft_weight=out32[5])
            else:
                out64 = o64s[r].learn(100, EN, 0, forced=forced, rec=r64)
            ens._on_t.numpy()[r, 0, :B256 - EN] = d["online"]
            ens._ex_t.numpy()[r, 0, :EN] = d["expert"]
            ens._bc_t.numpy()[r, 0, :] = d["bc"]
            steps[r, 0]["noise"] = np.asarray(d["noise"], np.float32)
            recs.append((actor_step, out32, out64, r32, r64))
        out = ens.launch(1)
        for r, (actor_step, out32, out64, r32, r64) in enumerate(recs):
            tie = actor_step and r32["soft_count_margin"] < 1e-5
            for j, nm in enumerate(("critic_loss", "actor_loss", "bc_loss", "rl_loss")):
                a, b, c = float(out[r, 0, j]), float(out32[j]), float(out64[j])
                assert abs(a - b) <= 1e-5 * abs(b) + 4 * abs(b - c) + 1e-6, f"replica {r} step {t} {nm}: {a} vs {b}"
            if actor_step and not tie:
                assert out[r, 0, 6] / B256 == out32[5], f"replica {r} step {t}: soft count {out[r, 0, 6]} vs weight {out32[5]}"
            g = np.empty(_lib.CRITIC_PARAMS, np.float32)
            _lib.check(lib.hirl_grads_export(h, r, _lib.NET_CRITIC, _lib.ptr(g), g.size))
            flat = lambda rec, key: np.concatenate([v.double().reshape(-1).numpy() for v in rec[key].values()])
            worst = max(worst, _check(f"replica {r} step {t} critic grads", g, flat(r32, "critic.grads"), flat(r64, "critic.grads")))
            if actor_step and not tie:
                g = np.empty(_lib.ACTOR_PARAMS, np.float32)
                _lib.check(lib.hirl_grads_export(h, r, _lib.NET_ACTOR, _lib.ptr(g), g.size))
                worst = max(worst, _check(f"replica {r} step {t} actor grads", g, flat(r32, "actor.grads"), flat(r64, "actor.grads")))
            if not tie:
                for w, nid in nets4:
                    _check(f"replica {r} step {t} weights {w}", ens.export_flat(r, nid), orcs[r].flat(w), o64s[r].flat(w))
            # teacher forcing: every implementation continues from the fp32 oracle's state
            for w, nid in nets4:
                v = orcs[r].flat(w).astype(np.float32)
                _lib.check(lib.hirl_params_import(h, r, nid, _lib.ptr(v), v.size))
                o64s[r].nets()[w].load_state_dict({k: x.double() for k, x in orcs[r].nets()[w].state_dict().items()})
            for w, nid in (("actor", _lib.NET_ACTOR), ("critic", _lib.NET_CRITIC)):
                net = orcs[r].nets()[w]
                if net.optimizer.state:
                    m, v = orcs[r].adam_state(w)
                    m, v = _lib.f32c(m), _lib.f32c(v)
                    _lib.check(lib.hirl_adam_import(h, r, nid, _lib.ptr(m), _lib.ptr(v), m.size))
                    for p32, p64 in zip(net.parameters(), o64s[r].nets()[w].parameters()):
                        s64 = o64s[r].nets()[w].optimizer.state[p64]
                        s64["exp_avg"].copy_(net.optimizer.state[p32]["exp_avg"].double())
                        s64["exp_avg_sq"].copy_(net.optimizer.state[p32]["exp_avg_sq"].double())
    print(f"ensemble B=256 (tcgen05 path), {len(seeds)} replicas x {STEPS} teacher-forced calls: worst gradient rel-L2 vs torch fp32 {worst:.2e}")

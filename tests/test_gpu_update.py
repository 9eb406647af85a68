"""GPU: the fused update through the C ABI against the oracle (torch fp32, bit-identical to the
unmodified reference), its fp64 twin, and the reference's own frozen outputs.

Tolerances (stated, fp32 path):
  * replay/expert/BC index streams, noise stream, soft count: exact
  * per-step teacher-forced losses / gradients / updated weights: rel-L2 <= 1e-5 against torch fp32
    on ordinary steps.  A step where torch-fp32 itself is more than 3e-6 away from the fp64 twin
    (a ReLU mask / soft-count tie flipped by summation order) is a "tie step": there
    only the envelope  err(cuda, fp64) <= 4 * err(torch32, fp64) + 2e-6  is required.
  * all GEMM back-end choices are exercised: FP32-FMA tiles (HIRL_TC=0), tcgen05 3xTF32 tiles (HIRL_TC=1) and the
    mid-size default where only stage G1 runs on the tensor cores (HIRL_TC=2)
  * free-running drift: within 5x the natural fp32-vs-fp64 drift of the reference, same run.
"""
import os

import numpy as np
import pytest
import torch

from hirl4ucav_b200 import _lib
from hirl4ucav_b200.agents.HIRL import Agent
from hirl4ucav_b200.agents.TD3 import Agent as TD3Agent
from oracle import scenarios
from oracle.hirl_oracle import OracleActor, OracleAgent

pytestmark = pytest.mark.gpu
NETS = (("actor", 0), ("targetActor", 1), ("critic", 2), ("targetCritic", 3))


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def grads_flat(rec, key):
    return np.concatenate([v.double().reshape(-1).numpy() for v in rec[key].values()])


def sync_to_oracle(ag, orc, o64=None):
    """Teacher forcing: every implementation continues from the fp32 oracle's state."""
    for w, _ in NETS + (("bc_actor", 4),):
        if orc.nets().get(w) is None:
            continue
        getattr(ag, w).load_state_dict(orc.nets()[w].state_dict())
        if o64 is not None:
            o64.nets()[w].load_state_dict({k: t.double() for k, t in orc.nets()[w].state_dict().items()})
    ag.push_parameters()
    for w, nid in (("actor", _lib.NET_ACTOR), ("critic", _lib.NET_CRITIC)):
        net = orc.nets()[w]
        if net.optimizer.state:
            m, v = orc.adam_state(w)
            ag.import_adam(nid, m, v, ag._actor_t if w == "actor" else ag._critic_t)
            if o64 is not None:
                for p32, p64 in zip(net.parameters(), o64.nets()[w].parameters()):
                    s32 = net.optimizer.state[p32]
                    s64 = o64.nets()[w].optimizer.state[p64]
                    s64["exp_avg"].copy_(s32["exp_avg"].double())
                    s64["exp_avg_sq"].copy_(s32["exp_avg_sq"].double())


def check_step(tag, got, ref32, ref64, tol=1e-5):
    e_c32, e_c64, e_o64 = rel(got, ref32), rel(got, ref64), rel(ref32, ref64)
    if e_c32 <= tol:            # the stated tolerance against torch fp32
        return e_c32
    # otherwise this must be a tie step (torch fp32 itself is off the fp64 twin) and we must be inside its envelope;
    # + 2e-6: the tensor-core path (3xTF32, truncating accumulator) is fp32-accurate to ~6e-7 per contraction
    assert e_o64 > 3e-6, f"{tag}: rel-L2 {e_c32:.2e} vs torch fp32 (> {tol}) on an ordinary step"
    assert e_c64 <= 4 * e_o64 + 2e-6, f"{tag}: outside the fp32 envelope: cuda-f64 {e_c64:.2e} vs torch32-f64 {e_o64:.2e}"
    return e_c32


def _backend(monkeypatch, mode):
    """'cluster' / 'cluster18': the cluster-resident kernel (row blocks of 16 / 18 rows); '0' / '1' / '2': round 1's
    stage-structured kernel with FP32 tiles / tcgen05 everywhere / tcgen05 for stage G1 only."""
    if mode == "tcc":          # the tensor-core cluster kernel (csrc/tcc_update.cuh)
        monkeypatch.delenv("HIRL_TC", raising=False)
        monkeypatch.setenv("HIRL_CLUSTER", "0")
        monkeypatch.setenv("HIRL_TCC", "1")
    elif mode.startswith("cluster"):
        monkeypatch.delenv("HIRL_TC", raising=False)
        monkeypatch.setenv("HIRL_TCC", "0")
        monkeypatch.setenv("HIRL_CLUSTER", "1")
        monkeypatch.setenv("HIRL_RI", "9" if mode == "cluster18" else "8")
    else:
        monkeypatch.setenv("HIRL_CLUSTER", "0")
        monkeypatch.setenv("HIRL_TCC", "0")
        monkeypatch.setenv("HIRL_TC", mode)


@pytest.mark.parametrize("tc_mode", ["tcc", "cluster", "cluster18", "0", "1", "2"])
@pytest.mark.parametrize("B,bcw,en,steps", [(32, 100, 8, 12), (32, 0.5, 0, 12), (256, 100, 64, 12), (100, 100, 33, 6), (5, 0.3, 2, 6)])
def test_teacher_forced_parity(B, bcw, en, steps, tc_mode, monkeypatch):
    _backend(monkeypatch, tc_mode)
    _teacher_forced(B, bcw, en, steps)


@pytest.mark.parametrize("bcw,en", [(0.5, 1024), (0.5 - 3 / 5000, 0)])
def test_teacher_forced_parity_at_baseline_batch_4096(bcw, en):
    """BASELINE.json configs fixed4096 / linear4096 at their full batch size, default back-end for that size
    (every contraction on tcgen05, dW2 split over the batch); same bars as the small cases."""
    _teacher_forced(4096, bcw, en, 4)


def _teacher_forced(B, bcw, en, steps):
    sc = scenarios.make_scenario(max(1000, 4 * B), max(600, 3 * B))
    orc = scenarios.build_agent(OracleAgent, sc, B, 0)
    o64 = scenarios.build_agent(OracleAgent, sc, B, 0, dtype=torch.float64)
    ag = scenarios.build_agent(Agent, sc, B, 0)
    for w, nid in NETS + (("bc_actor", 4),):
        assert np.array_equal(ag.export_flat(nid).astype(np.float64), orc.flat(w)), f"seeded init differs: {w}"
    scenarios.seed_all(1)
    worst = 0.0
    for t in range(steps):
        r32, r64 = {}, {}
        actor_step = orc.actorTrainable
        out32 = orc.learn(bcw, en, 0, rec=r32)
        d = orc.last_draws
        forced = {"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]}
        tie_count = actor_step and bcw == 100 and r32["soft_count_margin"] < 1e-5
        if actor_step and bcw == 100:
            forced64 = dict(forced)
            out64 = o64.learn(bcw, en, 0, forced=forced64, rec=r64, force_soft_weight=out32[5])
        else:
            out64 = o64.learn(bcw, en, 0, forced=forced, rec=r64)
        out = ag.learn(bcw, en, 0, forced=forced)
        # scalars (HIRL.py:334)
        for j, nm in enumerate(("critic_loss", "actor_loss", "bc_loss", "rl_loss")):
            a, b, c = float(out[j]), float(out32[j]), float(out64[j])
            assert abs(a - b) <= 1e-5 * abs(b) + 4 * abs(b - c) + 1e-6, f"step {t} {nm}: {a} vs {b}"
        assert abs(out[4] - out32[4]) <= 1e-5 * abs(out32[4]) + 1e-6
        if not tie_count:
            assert out[5] == out32[5], f"step {t}: bc_weight {out[5]} vs {out32[5]}"    # soft count is exact
        worst = max(worst, check_step(f"step {t} critic grads", ag.export_grads(_lib.NET_CRITIC), grads_flat(r32, "critic.grads"),
                                      grads_flat(r64, "critic.grads")))
        if actor_step and not tie_count:
            worst = max(worst, check_step(f"step {t} actor grads", ag.export_grads(_lib.NET_ACTOR), grads_flat(r32, "actor.grads"),
                                          grads_flat(r64, "actor.grads")))
        if not tie_count:
            for w, nid in NETS:
                # Adam turns a sign flip of a ~0 gradient into a +-lr step: rel-L2, not max-abs
                check_step(f"step {t} weights {w}", ag.export_flat(nid), orc.flat(w), o64.flat(w), tol=1e-5)
        sync_to_oracle(ag, orc, o64)
    print(f"B={B}: worst gradient rel-L2 vs torch fp32 over {steps} teacher-forced steps: {worst:.2e}")


DB_RLQ, DB_SOFTQ = 14, 15   # csrc/common.cuh: enum DebugBuf


def test_soft_count_is_tie_aware_per_row():
    """the soft BC weight is a COUNT of strict per-row comparisons soft_Q > rl_Q (HIRL.py:303).  Per row the
    kernel's two Q values must match the oracle's to fp32 accuracy, its comparison may differ from the oracle's only where
    the oracle's own margin |soft_Q - rl_Q| is below eps (a tie that summation order decides), and the weight the call
    returns must be exactly the mean of the kernel's own per-row comparisons."""
    B, en = 256, 64
    sc = scenarios.make_scenario(1500, 900)
    orc = scenarios.build_agent(OracleAgent, sc, B, 0)
    ag = scenarios.build_agent(Agent, sc, B, 0)
    scenarios.seed_all(7)
    checked = flips = 0
    for t in range(10):
        r32 = {}
        actor_step = orc.actorTrainable
        out32 = orc.learn(100, en, 0, rec=r32)
        d = orc.last_draws
        out = ag.learn(100, en, 0, forced={"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]})
        if actor_step:
            softq, rlq = ag.debug_buffer(DB_SOFTQ)[:B].astype(np.float64), ag.debug_buffer(DB_RLQ)[:B].astype(np.float64)
            o_s, o_r = r32["soft_q"].double().numpy().ravel(), r32["rl_q"].double().numpy().ravel()
            scale = max(1.0, float(np.abs(o_r).max()), float(np.abs(o_s).max()))
            assert np.abs(softq - o_s).max() <= 1e-5 * scale and np.abs(rlq - o_r).max() <= 1e-5 * scale
            mine, theirs = softq > rlq, o_s > o_r
            eps = 2e-5 * scale
            flipped = mine != theirs
            assert not (flipped & (np.abs(o_s - o_r) >= eps)).any(), f"step {t}: comparison differs on a row with a clear margin"
            assert out[5] == min(np.float32(mine.sum()) / np.float32(B), 1.0), f"step {t}: returned weight {out[5]} vs own count {mine.sum()}"
            assert abs(int(mine.sum()) - int(theirs.sum())) <= int(flipped.sum())
            checked += 1
            flips += int(flipped.sum())
        sync_to_oracle(ag, orc)
    assert checked == 5
    print(f"soft count: {checked} actor calls x {B} rows compared, {flips} tie rows decided differently")


def test_soft_count_exact_ties_count_as_not_greater():
    """bc_actor == actor makes Q1(s, a_bc) and Q1(s, pi(s)) the same number on every row: the reference's strict '>' gives
    weight 0.0.  The kernel evaluates both through the same code on the same inputs, so it must see exact ties too."""
    B = 128
    sc = scenarios.make_scenario(1000, 600)
    orc = scenarios.build_agent(OracleAgent, sc, B, 0)
    ag = scenarios.build_agent(Agent, sc, B, 0)
    orc.bc_actor.load_state_dict(orc.actor.state_dict())
    ag.bc_actor.load_state_dict(orc.actor.state_dict())
    ag.push_parameters()
    scenarios.seed_all(3)
    out32 = orc.learn(100, 16, 0)
    d = orc.last_draws
    out = ag.learn(100, 16, 0, forced={"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]})
    assert out32[5] == 0.0 and out[5] == 0.0
    assert np.array_equal(ag.debug_buffer(DB_SOFTQ)[:B], ag.debug_buffer(DB_RLQ)[:B])


@pytest.mark.parametrize("name", ["soft32", "fixed32", "soft256"])
def test_free_run_matches_reference_fixture(golden_dir, name):
    """Same seeds as the unmodified reference: own RNG draws, own weights, 12 free-running calls."""
    fx = np.load(os.path.join(golden_dir, f"agent_{name}.npz"))
    B, steps, bcw, en = int(fx["batch"]), int(fx["steps"]), float(fx["bc_weight_now"]), int(fx["expert_num_now"])
    sc = scenarios.make_scenario(int(fx["n_online"]), int(fx["n_expert"]))
    ag = scenarios.build_agent(Agent, sc, B, int(fx["seed"]))
    for w, nid in NETS + (("bc_actor", 4),):
        assert float(ag.export_flat(nid).astype(np.float64).sum()) == float(fx[f"{w}.init_sum"])
    scenarios.seed_all(int(fx["run_seed"]))
    for t in range(steps):
        out = ag.learn(bcw if bcw != 100 else 100, en, 0)
        ref = fx["outs"][t]
        assert np.array_equal(ag.last_draws["online"], fx["idx_online"][t]) and np.array_equal(ag.last_draws["noise"], fx["noise"][t])
        for j in range(4):
            assert abs(float(out[j]) - ref[j]) <= 2e-4 * abs(ref[j]) + 1e-5, f"step {t} out[{j}]: {float(out[j])} vs {ref[j]}"
        assert abs(out[5] - ref[5]) <= 1.0 / B + 1e-12
    for w, nid in NETS:
        got = ag.export_flat(nid)[fx[f"{w}.probe_idx"]]
        assert rel(got, fx[f"{w}.probe"]) <= 2e-4, w


def test_upsampling_free_run_matches_reference_fixture(golden_dir):
    """prioritised up-sampling of the online ring (weighted draw WITH replacement, buffer.py:39-43) and the
    firing-row replacement of the cloning batch (HIRL.py:253-257).  Index streams bit-exact, scalars to fp32 tolerance."""
    fx = np.load(os.path.join(golden_dir, "agent_upsample32.npz"))
    B, steps, en = int(fx["batch"]), int(fx["steps"]), int(fx["expert_num_now"])
    sc = scenarios.make_scenario(int(fx["n_online"]), int(fx["n_expert"]))
    ag = scenarios.build_agent(Agent, sc, B, 0, buffer_upsample=True, expert_upsample=True)
    assert ag.buffer.upsample and ag.buffer.priorities_memory[:len(ag.buffer)].max() == 1.0
    scenarios.seed_all(1)
    for t in range(steps):
        out = ag.learn(100, en, 0)
        ref = fx["outs"][t]
        d = ag.last_draws
        assert np.array_equal(d["online"], fx["idx_online"][t]) and np.array_equal(d["expert"], fx["idx_expert"][t])
        assert np.array_equal(d["bc"], fx["idx_bc"][t]) and np.array_equal(d["noise"], fx["noise"][t])
        for j in range(4):
            assert abs(float(out[j]) - ref[j]) <= 2e-4 * abs(ref[j]) + 1e-5, f"step {t} out[{j}]: {float(out[j])} vs {ref[j]}"
        assert abs(out[5] - ref[5]) <= 1.0 / B + 1e-12
    for w, nid in NETS:
        assert rel(ag.export_flat(nid)[fx[f"{w}.probe_idx"]], fx[f"{w}.probe"]) <= 2e-4, w


def test_learn_many_is_bit_identical_to_single_calls_and_deterministic():
    sc = scenarios.make_scenario(2000, 600)
    runs = []
    for mode in ("single", "fused", "fused"):
        ag = scenarios.build_agent(Agent, sc, 64, 3)
        scenarios.seed_all(5)
        if mode == "single":
            outs = [ag.learn(100, 16, 0) for _ in range(12)]
        else:
            outs = ag.learn_many([(100, 16, 0)] * 12)
        runs.append((np.array([[float(x) for x in o] for o in outs]), [ag.export_flat(nid) for _, nid in NETS]))
    for other in runs[1:]:
        assert np.array_equal(runs[0][0], other[0])
        for a, b in zip(runs[0][1], other[1]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("B,en", [(256, 64), (100, 25)])
def test_tcc_results_do_not_depend_on_timing(B, en, monkeypatch):
    """The cluster kernel hands data between CTAs / clusters through barriers, mbarriers and release/acquire flags only.
    HIRL_TCC_JITTER makes every CTA (and warp) sleep a pseudo-random 0-4 us at each synchronisation site: 12 fused calls
    (the persistent launch overlaps consecutive calls where no grid barrier is needed) and 6 single calls must give the
    same bits as the unperturbed run."""
    monkeypatch.setenv("HIRL_TCC", "1")
    monkeypatch.setenv("HIRL_CLUSTER", "0")
    sc = scenarios.make_scenario(max(1000, 4 * B), max(600, 3 * B))
    runs = []
    for seed, ns in (("0", "4000"), ("1", "4000"), ("7", "4000"), ("12345", "4000"), ("3", "300000")):   # the last: up to 0.3 ms per site
        monkeypatch.setenv("HIRL_TCC_JITTER", seed)
        monkeypatch.setenv("HIRL_TCC_JITTER_NS", ns)
        ag = scenarios.build_agent(Agent, sc, B, 3)
        assert ag.backend == "tcc"
        scenarios.seed_all(5)
        outs = ag.learn_many([(100, en, 0)] * 12) + [ag.learn(100, en, 0) for _ in range(6)]
        runs.append((np.array([[float(x) for x in o] for o in outs]), [ag.export_flat(nid) for _, nid in NETS]))
        del ag
    for other in runs[1:]:
        assert np.array_equal(runs[0][0], other[0])
        for a, b in zip(runs[0][1], other[1]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("tc_mode", ["0", "1"])
@pytest.mark.parametrize("ns", ["1", "2", "4"])
def test_split_k_variants_agree(ns, tc_mode, monkeypatch):
    monkeypatch.setenv("HIRL_NSPLIT", ns)        # split-K of the forward / dX tiles (FP32 path)
    monkeypatch.setenv("HIRL_DW_SPLIT", ns)      # split of the batch reduction of dW2 (both paths)
    monkeypatch.setenv("HIRL_TC", tc_mode)
    sc = scenarios.make_scenario(1500, 600)
    orc = scenarios.build_agent(OracleAgent, sc, 48, 0)
    ag = scenarios.build_agent(Agent, sc, 48, 0)
    scenarios.seed_all(2)
    for t in range(6):
        r32 = {}
        ref = orc.learn(100, 10, 0, rec=r32)
        d = orc.last_draws
        out = ag.learn(100, 10, 0, forced={"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]})
        assert abs(float(out[0]) - float(ref[0])) <= 1e-5 * abs(float(ref[0]))
        assert rel(ag.export_grads(_lib.NET_CRITIC), grads_flat(r32, "critic.grads")) <= 1e-5
        if "actor.grads" in r32 and r32.get("soft_count_margin", 1) > 1e-4:
            assert rel(ag.export_grads(_lib.NET_ACTOR), grads_flat(r32, "actor.grads")) <= 1e-5
        sync_to_oracle(ag, orc)


def test_target_update_schedule_matches_oracle():
    sc = scenarios.make_scenario(1000, 600)
    orc = scenarios.build_agent(OracleAgent, sc, 32, 0)
    ag = scenarios.build_agent(Agent, sc, 32, 0)
    scenarios.seed_all(4)
    changed_o, changed_c = [], []
    prev_o, prev_c = orc.flat("targetCritic").copy(), ag.export_flat(3).copy()
    for t in range(12):
        orc.learn(0.5, 4, 0)
        d = orc.last_draws
        ag.learn(0.5, 4, 0, forced={"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]})
        cur_o, cur_c = orc.flat("targetCritic"), ag.export_flat(3)
        changed_o.append(not np.array_equal(cur_o, prev_o)); changed_c.append(not np.array_equal(cur_c, prev_c))
        if changed_c[-1]:
            assert rel(cur_c, cur_o) < 1e-7          # Polyak arithmetic (HIRL.py:11-13) on the same inputs
        sync_to_oracle(ag, orc)
        prev_o, prev_c = cur_o.copy(), ag.export_flat(3).copy()
    assert changed_o == changed_c and sum(changed_c) == 2   # fires on calls 5 and 11 (every 3rd actor update)


@pytest.mark.parametrize("steps,mark_at", [(300, (30, 100, 300)), (1000, (100, 300, 1000))])
def test_drift_envelope(steps, mark_at):
    """Free-running drift against the fp64 twin over 300 and over 1 000 calls (north star: "drift bounded over 1k steps")."""
    B = 256
    sc = scenarios.make_scenario(4000, 2000)
    orc = scenarios.build_agent(OracleAgent, sc, B, 0)
    o64 = scenarios.build_agent(OracleAgent, sc, B, 0, dtype=torch.float64)
    ag = scenarios.build_agent(Agent, sc, B, 0)
    scenarios.seed_all(9)
    forced_all = []
    marks = {}
    for t in range(steps):
        orc.learn(0.5, 64, 0)
        d = orc.last_draws
        f = {"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]}
        o64.learn(0.5, 64, 0, forced=f)
        forced_all.append(f)
        if t + 1 in mark_at:
            marks[t + 1] = (orc.flat("critic").copy(), o64.flat("critic").copy(), orc.flat("actor").copy(), o64.flat("actor").copy())
    done = 0
    for t_mark in mark_at:
        ag.learn_many([(0.5, 64, 0)] * (t_mark - done), forced=forced_all[done:t_mark])
        done = t_mark
        c32, c64, a32, a64 = marks[t_mark]
        nat_c, nat_a = rel(c32, c64), rel(a32, a64)
        got_c, got_a = rel(ag.export_flat(2), c64), rel(ag.export_flat(0), a64)
        print(f"drift @{t_mark}: critic cuda-f64 {got_c:.2e} (torch32-f64 {nat_c:.2e}); actor {got_a:.2e} ({nat_a:.2e})")
        # actor and critic are coupled (each trains through the other), so the envelope is the larger of the
        # two natural fp32-vs-fp64 drifts of the reference at that step count, times 5
        env = 5 * max(nat_c, nat_a, 1e-5)
        assert got_c <= env and got_a <= env


def test_checkpoints_interoperate_with_reference_shaped_modules(tmp_path):
    sc = scenarios.make_scenario(800, 400)
    ag = scenarios.build_agent(Agent, sc, 32, 0)
    scenarios.seed_all(1)
    ag.learn_many([(100, 8, 0)] * 6)
    ag.saveCheckpoints("Agent7_", str(tmp_path))
    sd = torch.load(str(tmp_path) + "\\Agent7_" + "Actor_Harfang_GYM")
    assert list(sd) == ["full1.weight", "full1.bias", "layernorm1.weight", "layernorm1.bias", "full2.weight", "full2.bias",
                        "layernorm2.weight", "layernorm2.bias", "final.weight", "final.bias"]
    assert tuple(sd["full1.weight"].shape) == (256, 13) and tuple(sd["final.weight"].shape) == (4, 512)
    ref_actor = OracleActor(1e-3, 13, 4, 256, 512, True, "x")
    ref_actor.load_state_dict(sd)                      # loads into the reference-shaped module
    s = sc["s"][:7]
    want = ref_actor(torch.tensor(s, dtype=torch.float)).detach().numpy()
    got = ag.act_batch(s)
    assert np.abs(got - want).max() <= 2e-6
    # round trip: a second agent restored from the files acts identically
    ag2 = scenarios.build_agent(Agent, sc, 32, 1)
    ag2.loadCheckpoints("Agent7_", str(tmp_path))
    assert np.array_equal(ag2.act_batch(s), got)
    for nid in range(4):
        assert np.array_equal(ag2.export_flat(nid), ag.export_flat(nid))
    # bc_actor file name = prefix + name (HIRL.py:180, 348)
    torch.save(ref_actor.state_dict(), str(tmp_path) + "\\BC_" + "Harfang_GYM")
    ag2.load_bc_actor("BC_", str(tmp_path))
    assert np.array_equal(ag2.export_flat(4), ag.export_flat(0))
    ag.save_training_state(str(tmp_path / "extra"))
    ag2.load_training_state(str(tmp_path / "extra"))
    assert ag2.update_count == ag.update_count and ag2._critic_t == ag._critic_t


def test_choose_action_matches_oracle():
    sc = scenarios.make_scenario(500, 300)
    orc = scenarios.build_agent(OracleAgent, sc, 32, 0)
    ag = scenarios.build_agent(Agent, sc, 32, 0)
    s0 = sc["s"][11]
    scenarios.seed_all(3)
    want = [orc.chooseAction(s0), orc.chooseActionSmallNoise(s0), orc.chooseActionNoNoise(s0)]
    scenarios.seed_all(3)
    got = [ag.chooseAction(s0), ag.chooseActionSmallNoise(s0), ag.chooseActionNoNoise(s0)]
    for g, w in zip(got, want):
        assert g.shape == (4,) and g.dtype == np.float32 and np.abs(g - w).max() <= 1e-6
    assert Agent.select_action is Agent.chooseAction and Agent.train is Agent.learn


def test_td3_mode_matches_oracle():
    sc = scenarios.make_scenario(2000, 300)
    hp = scenarios.HP
    scenarios.seed_all(0)
    orc = OracleAgent(hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], 10 ** 5, 64, True, "Harfang_GYM",
                      None, None, 0.0, False, mode="TD3")
    scenarios.seed_all(0)
    ag = TD3Agent(hp["actorLR"], hp["criticLR"], 13, 4, 256, 512, hp["tau"], hp["gamma"], 10 ** 5, 64, True, "Harfang_GYM")
    for a in (orc, ag):
        for i in range(2000):
            a.store(sc["s"][i], sc["a"][i], sc["s2"][i], float(sc["r"][i]), bool(sc["d"][i]), int(sc["succ"][i]))
    for w, nid in NETS:
        assert np.array_equal(ag.export_flat(nid).astype(np.float64), orc.flat(w))
    scenarios.seed_all(1)
    for t in range(8):
        r32 = {}
        ref = orc.learn(rec=r32)
        d = orc.last_draws
        out = ag.learn(forced={"online": d["online"], "noise": d["noise"]})
        assert abs(float(out[0]) - float(ref[0])) <= 1e-5 * abs(float(ref[0])) + 1e-6
        assert abs(float(out[1]) - float(ref[1])) <= 1e-5 * abs(float(ref[1])) + 1e-6
        assert rel(ag.export_grads(_lib.NET_CRITIC), grads_flat(r32, "critic.grads")) <= 1e-5
        sync_to_oracle(ag, orc)


def test_bc_mode_matches_oracle_and_reference_fixture(golden_dir):
    from hirl4ucav_b200.agents.BC import Agent as BCAgent
    from oracle.hirl_oracle import OracleBC
    fx = np.load(os.path.join(golden_dir, "agent_td3_bc.npz"))
    sc = scenarios.make_scenario(2000, 800)
    scenarios.seed_all(0)
    obc = OracleBC(1e-3, 13, 4, 256, 512, True, "Harfang_GYM", 64, sc["expert_states"], sc["expert_actions"])
    scenarios.seed_all(0)
    ag = BCAgent(1e-3, 13, 4, 256, 512, True, "Harfang_GYM", 64, sc["expert_states"], sc["expert_actions"])
    want0 = np.concatenate([p.detach().double().reshape(-1).numpy() for p in obc.actor.parameters()])
    assert np.array_equal(ag.export_flat(0).astype(np.float64), want0)           # same seeded init as the reference's single Actor
    scenarios.seed_all(1)
    for t in range(6):
        ref = float(obc.train_actor())
        got = float(ag.train_actor(forced_bc=obc.last_bc))
        assert ref == float(fx["bc.losses"][t])                                    # oracle == unmodified reference (frozen)
        assert abs(got - ref) <= 1e-5 * abs(ref) + 1e-7, (t, got, ref)
        ag.actor.load_state_dict(obc.actor.state_dict()); ag.push_parameters()    # teacher forcing
        st = obc.actor.optimizer.state
        m = np.concatenate([st[p]["exp_avg"].reshape(-1).numpy() for p in obc.actor.parameters()])
        v = np.concatenate([st[p]["exp_avg_sq"].reshape(-1).numpy() for p in obc.actor.parameters()])
        ag.import_adam(_lib.NET_ACTOR, m, v, ag._actor_t)
    # free-running from the same seed: the drop-in draws the same np.random.choice stream as the reference
    scenarios.seed_all(0)
    ag2 = BCAgent(1e-3, 13, 4, 256, 512, True, "Harfang_GYM", 64, sc["expert_states"], sc["expert_actions"])
    scenarios.seed_all(1)
    losses = [float(ag2.train_actor()) for _ in range(6)]
    assert np.allclose(losses, fx["bc.losses"], rtol=2e-4)
    assert rel(ag2.export_flat(0)[np.sort(np.random.default_rng(99).choice(138756, 1024, replace=False))], fx["bc.actor.probe"]) < 2e-4


def test_error_behaviour_matches_reference():
    sc = scenarios.make_scenario(100, 80)
    ag = scenarios.build_agent(Agent, sc, 32, 0)
    with pytest.raises(TypeError):       # reference: Transition(*zip(*[])) when no online rows are drawn
        ag.learn(0.5, 32, 0)
    with pytest.raises(ValueError):      # reference: random.sample larger than population
        scenarios.build_agent(Agent, scenarios.make_scenario(10, 80), 32, 0).learn(0.5, 0, 0)
    with pytest.raises(NotImplementedError):
        Agent(1e-3, 1e-3, 13, 4, 128, 128, 0.005, 0.99, 1000, 32, True, "x", sc["expert_states"], sc["expert_actions"], 0.5, True)
    with pytest.raises(ValueError):      # np.random.choice(N, B, replace=False) with N < B
        Agent(1e-3, 1e-3, 13, 4, 256, 512, 0.005, 0.99, 1000, 128, True, "x", sc["expert_states"][:50], sc["expert_actions"][:50],
              0.5, True).learn(0.5, 0, 0)

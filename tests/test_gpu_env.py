"""GPU: env kernels through the C ABI against the oracle, the golden fixtures and
size-independent properties at the BASELINE size (10 M states)."""
import os

import numpy as np
import pytest
import torch

from hirl4ucav_b200 import synth
from hirl4ucav_b200.environments import harfang_math as hm
from oracle import env_oracle

pytestmark = pytest.mark.gpu


def ulp_diff_f32(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def check_against_oracle(raw, obs, rew, fl, obs_ref, rew_ref, fl_ref):
    assert np.array_equal(fl, fl_ref), "flags must be bit-exact"
    assert np.array_equal(obs, obs_ref.astype(np.float32)), "observations must be bit-exact after the fp32 cast"
    # reward: the reference's pow(x, .5) differs from IEEE sqrt by <= 1 f64 ulp on 0.08 % of inputs; after
    # rounding to fp32 this is invisible except on rounding boundaries: tolerance 1 ulp (fp32), stated.
    u = ulp_diff_f32(rew, rew_ref.astype(np.float32))
    assert u.max() <= 1
    assert (u == 0).mean() > 0.9999


def test_golden_fixture(golden_dir):
    fx = np.load(os.path.join(golden_dir, "env_online.npz"))
    obs, rew, fl = hm.step_math(fx["raw"])
    check_against_oracle(fx["raw"], obs, rew, fl, fx["obs"], fx["reward"], fx["flags"])


@pytest.mark.parametrize("n", [0, 1, 31, 255, 256, 257, 100_003])
def test_random_and_ragged_sizes(n):
    raw = synth.raw_states(np.random.default_rng(n + 5), max(n, 1))[:n]
    obs, rew, fl = hm.step_math(raw)
    assert obs.shape == (n, 13) and rew.shape == (n,) and fl.shape == (n,)
    if n:
        o, r, f = env_oracle.online_step(raw)
        check_against_oracle(raw, obs, rew, fl, o, r, f)


def test_threshold_edges():
    raw = synth.raw_states(np.random.default_rng(0), 4096)
    hv = np.array([0.0, -0.0, synth.F32_TENTH, np.nextafter(np.float32(0.1), np.float32(0)),
                   np.nextafter(np.float32(0.1), np.float32(1)), 0.2, 1e-30, 0.05], np.float32)
    av = np.array([499.99997, 500.0, 500.00003, 1999.9999, 2000.0, 2000.0001, 6999.9995, 7000.0, 7000.0005,
                   9999.999, 10000.0, 10000.001], np.float32)
    i = np.arange(4096)
    raw[:, 13] = hv[i % len(hv)]
    raw[:, 1] = av[(i // len(hv)) % len(av)]
    raw[:, 15] = (i % 32).astype(np.uint32).view(np.float32)
    raw[:, 14] = np.array([0.0, 1e-9, -1e-9, 0.5], np.float32)[(i // 7) % 4]
    obs, rew, fl = hm.step_math(raw)
    o, r, f = env_oracle.online_step(raw)
    check_against_oracle(raw, obs, rew, fl, o, r, f)


def test_device_path_and_counters_10m():
    """BASELINE config 5 size: device-resident path, checked on a slice + global properties."""
    n = 10_000_000
    rng = np.random.default_rng(77)
    raw = synth.raw_states(rng, 1 << 20)
    raw = np.tile(raw, (10, 1))[:n]
    d_raw = torch.from_numpy(raw).cuda()
    obs, rew, fl, cnt = hm.step_math(d_raw, counters=True)
    torch.cuda.synchronize()
    fl_h = fl.cpu().numpy()
    sl = slice(3_000_000, 3_200_000)
    o, r, f = env_oracle.online_step(raw[sl])
    check_against_oracle(raw[sl], obs[sl].cpu().numpy(), rew[sl].cpu().numpy(), fl_h[sl], o, r, f)
    # checksum of checksums: counters equal the flag reductions; flags over the whole 10 M equal the oracle's
    fire, succ, done = env_oracle.infinite_counters(fl_h)
    assert cnt.cpu().tolist() == [fire, succ, done]
    if env_oracle.c_lib() is not None:
        _, _, f_all = env_oracle.online_step_c(raw)
        assert np.array_equal(fl_h, f_all)
    # idempotence / tiling property: the 1 Mi-state pattern repeats
    assert torch.equal(fl[: 1 << 20], fl[1 << 20: 2 << 20])
    assert torch.equal(obs[: 1 << 20], obs[(1 << 20) * 5: (1 << 20) * 6])


def test_offline_functions(golden_dir):
    fo = np.load(os.path.join(golden_dir, "env_offline.npz"))
    r, su, dn = hm.offline_reward(fo["s"], fo["a"], fo["s2"])
    assert np.array_equal(su.astype(np.int32), fo["success"]) and np.array_equal(dn, fo["done"])
    assert ulp_diff_f32(r, fo["reward"].astype(np.float32)).max() <= 1
    env = hm.HarfangEnvMath()
    rr, ss = env.get_reward(fo["s"][3], fo["a"][3], fo["s2"][3])
    assert ss == fo["success"][3] and abs(rr - fo["reward"][3]) <= 1e-6 * max(1, abs(fo["reward"][3]))
    assert env.get_termination(fo["s2"][3]) == bool(fo["done"][3])


def test_offline_thresholds_use_float64_inputs():
    """the reference evaluates get_reward / get_termination on the float64 CSV values; health == 0.1 (double)
    satisfies next_state[12] <= 0.1 (GYM.py:334) while float32(0.1) = 0.10000000149 does not."""
    s = np.zeros((4, 13)); a = np.zeros((4, 4)); s2 = np.zeros((4, 13))
    s2[:, 12] = [0.1, float(np.float32(0.1)), np.nextafter(0.1, 0.0), 0.2]
    s[:, 8] = 1.0; s[:, 7] = [1.0, -1.0, 1e-50, 1.0]; a[:, 3] = [1e-60, 1.0, 1.0, -1.0]
    r, su, dn = hm.offline_reward(s, a, s2)
    o_r, o_su = env_oracle.offline_reward(s, a, s2)
    assert list(dn) == [True, False, True, False] == list(env_oracle.offline_done(s2))
    assert np.array_equal(su.astype(np.int32), o_su) and list(o_su) == [1, -1, 1, 0]
    assert np.array_equal(r, o_r.astype(np.float32))
    env = hm.HarfangEnvMath()
    assert env.get_termination(s2[0]) is True and env.get_termination(s2[0].astype(np.float32)) is False
    # the expert-buffer walk (train_all.py:287-306) follows the float64 done flags
    es = np.zeros((6, 13)); ea = np.zeros((6, 4))
    es[:, 12] = [0.5, 0.1, 0.5, float(np.float32(0.1)), 0.5, 0.5]
    es[:, 0] = np.arange(6) * 0.01
    rows, succ = hm.expert_scan(es, ea)
    use, r_o, d_o, s_o = env_oracle.expert_scan(es, ea)
    assert list(use) == [0, 2, 3, 4] and len(rows) == 4
    assert np.array_equal(rows[:, 0], es[use, 0].astype(np.float32)) and np.array_equal(rows[:, 31] != 0, d_o)


def test_expert_fire_rule_is_a_flag_bit():
    """the expert fires iff obs[7] > 0 and obs[8] > 0 (data/straight_line/ai_data_col.py:63-64) -- emitted
    by env_step_kernel as flag bit 4 next to the policy's rule a[3] > 0 (bit 3, GYM.py:150); both bit-exact."""
    raw = synth.raw_states(np.random.default_rng(12), 200_003)
    obs, rew, fl = hm.step_math(raw)
    d = hm.decode_flags(fl)
    assert np.array_equal(d["expert_fire"], (obs[:, 7] > 0) & (obs[:, 8] > 0))
    assert np.array_equal(d["expert_fire"], hm.HarfangEnvMath.expert_fires(obs))
    assert np.array_equal(d["fired"], hm.HarfangEnvMath.agent_fires(raw[:, 11:15]))
    od = env_oracle.decode_flags(env_oracle.online_step(raw)[2])
    assert np.array_equal(d["expert_fire"], od["expert_fire"]) and 0.05 < d["expert_fire"].mean() < 0.5


def test_expert_scan_matches_oracle():
    rng = np.random.default_rng(3)
    es, ea = synth.env_states(rng, 5000), synth.env_actions(rng, 5000, expert=True)
    es[100:104, 12] = 0.0     # a run of consecutive terminal rows
    rows, succ = hm.expert_scan(es, ea)
    use, r, d, s = env_oracle.expert_scan(es, ea)
    assert len(rows) == len(use)
    assert np.array_equal(rows[:, :13], es[use].astype(np.float32)) and np.array_equal(rows[:, 17:30], es[use + 1].astype(np.float32))
    assert np.array_equal(rows[:, 13:17], ea[use].astype(np.float32))
    assert np.array_equal(rows[:, 31] != 0, d) and np.array_equal(succ.astype(np.int32), s)
    assert ulp_diff_f32(rows[:, 30], r.astype(np.float32)).max() <= 1

"""GPU: device-resident collection -- batched acting and on-device store.

Bit-exactness bar: the ring rows written by the device path must equal, bit for bit, the rows the
reference-shaped host path (Agent.store per transition, buffer.py:20-36) writes for the same data; the
batched actions must equal the single-row chooseAction* path of the same kernel and stay within fp32
tolerance of the oracle's torch forward.
"""
import numpy as np
import pytest
import torch

from hirl4ucav_b200 import _lib, synth
from hirl4ucav_b200.agents.HIRL import Agent
from hirl4ucav_b200.environments import harfang_math as hm
from oracle import env_oracle, scenarios
from oracle.hirl_oracle import OracleAgent

pytestmark = pytest.mark.gpu


def _agents(B=32):
    sc = scenarios.make_scenario(600, 400)
    return sc, scenarios.build_agent(Agent, sc, B, 0), scenarios.build_agent(Agent, sc, B, 0)


def test_device_store_equals_host_store_bit_for_bit():
    sc, dev_ag, host_ag = _agents()
    rng = np.random.default_rng(5)
    n = 257                                               # ragged: not a multiple of the 8 rows a CTA packs
    raw_prev, raw_next = synth.raw_states(rng, n), synth.raw_states(rng, n)
    d = torch.device("cuda", 0)
    obs0, _, _ = hm.step_math(torch.from_numpy(raw_prev).to(d))
    act = dev_ag.act_batch_dev(obs0, noise=torch.tensor([0.05, -0.02, 0.3, -0.6]))
    raw_next[:, 14] = act[:, 3].cpu().numpy()             # the policy's fire channel feeds the reward (GYM.py:150)
    obs1, rew, fl = hm.step_math(torch.from_numpy(raw_next).to(d))
    before = len(dev_ag.buffer)
    dev_ag.store_batch_dev(obs0, act, obs1, rew, fl)
    assert len(dev_ag.buffer) == before + n and dev_ag.buffer.position == (before + n) % dev_ag.buffer.capacity

    # reference-shaped path: one store() per transition from host copies of the same values
    o0, a, o1, r, f = (t.cpu().numpy() for t in (obs0, act, obs1, rew, fl))
    dec = hm.decode_flags(f)
    for i in range(n):
        host_ag.store(o0[i], a[i], o1[i], float(r[i]), bool(dec["done"][i]), int(dec["step_success"][i]))
    host_ag.buffer.flush()
    idx = np.arange(before, before + n, dtype=np.int32)
    lib = _lib.load()
    got, want = np.empty((n, 32), np.float32), np.empty((n, 32), np.float32)
    torch.cuda.synchronize()
    _lib.check(lib.hirl_gather(dev_ag._h, 0, _lib.ptr(idx), n, None, 0, _lib.ptr(got)))
    _lib.check(lib.hirl_gather(host_ag._h, 0, _lib.ptr(idx), n, None, 0, _lib.ptr(want)))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the env outputs themselves against the CPU oracle (flags bit-exact, obs/reward = fp32 rounding of the f64 result)
    o_obs, o_rew, o_fl = env_oracle.online_step(raw_next)
    assert np.array_equal(f, o_fl) and np.array_equal(o1, o_obs.astype(np.float32)) and np.array_equal(r, o_rew.astype(np.float32))
    # lazily refreshed host mirror: memory[i] reads like the reference's list of Transitions
    tr = dev_ag.buffer.memory[before + 3]
    assert np.array_equal(tr.state, o0[3]) and np.array_equal(tr.action, a[3]) and tr.reward == float(r[3])
    assert tr.done == bool(dec["done"][3]) and tr.step_success == int(dec["step_success"][3])


def test_device_store_wraps_and_orders_with_host_stores():
    sc = scenarios.make_scenario(40, 400)
    ag = scenarios.build_agent(Agent, sc, 16, 0, hp={"bufferSize": 64})
    d = torch.device("cuda", 0)
    rng = np.random.default_rng(9)
    n = 40                                                 # 40 stored + 40 appended > capacity 64: wraps
    s, a, s2 = (torch.from_numpy(x.astype(np.float32)).to(d) for x in (synth.env_states(rng, n), synth.env_actions(rng, n), synth.env_states(rng, n)))
    r = torch.arange(n, dtype=torch.float32, device=d)
    fl = torch.zeros(n, dtype=torch.int32, device=d)
    ag.store(sc["s"][0], sc["a"][0], sc["s2"][0], -7.0, False, 0)   # host store BEFORE the device append lands first
    ag.store_batch_dev(s, a, s2, r, fl)
    assert len(ag.buffer) == 64 and ag.buffer.position == (41 + n) % 64
    assert ag.buffer.memory[40].reward == -7.0
    assert ag.buffer.memory[41].reward == 0.0 and ag.buffer.memory[63].reward == 22.0 and ag.buffer.memory[0].reward == 23.0
    out = ag.learn(0.5, 0, 0)                                # the update sees the appended rows (stream ordering)
    assert all(np.isfinite(float(x)) for x in out)


def test_batched_acting_matches_single_row_path_and_oracle():
    sc, ag, _ = _agents()
    orc = scenarios.build_agent(OracleAgent, sc, 32, 0)
    states = sc["s"][:200].astype(np.float32)
    d = torch.device("cuda", 0)
    got = ag.act_batch_dev(torch.from_numpy(states).to(d)).cpu().numpy()
    assert np.array_equal(got, ag.act_batch(states))                                      # same kernel, host-pointer entry
    assert np.array_equal(got[7], ag.chooseActionNoNoise(states[7]))
    want = np.stack([orc.chooseActionNoNoise(s) for s in states])
    assert np.abs(got - want).max() <= 2e-6                                               # fp32 forward, tanh output in [-1,1]
    # per-row noise and clamp (HIRL.py:196-198)
    nz = torch.randn(200, 4) * 2
    noisy = ag.act_batch_dev(torch.from_numpy(states).to(d), noise=nz).cpu().numpy()
    assert np.array_equal(noisy, np.clip(got + nz.numpy(), -1, 1).astype(np.float32))
    with pytest.raises(ValueError):
        ag.act_batch_dev(torch.from_numpy(states).to(d), noise=torch.zeros(3, 4))


def test_upsampling_priorities_follow_device_appends():
    sc = scenarios.make_scenario(100, 400)
    ag = scenarios.build_agent(Agent, sc, 16, 0, buffer_upsample=True)
    d = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    n = 24
    s, a, s2 = (torch.from_numpy(x.astype(np.float32)).to(d) for x in (synth.env_states(rng, n), synth.env_actions(rng, n), synth.env_states(rng, n)))
    succ = np.array([1, -1, 0] * 8, np.int32)
    fl = torch.from_numpy(((succ & 0xFF) << 8).astype(np.int32)).to(d)
    ag.store_batch_dev(s, a, s2, torch.zeros(n, device=d), fl)
    pr = ag.buffer.priorities_memory[100:100 + n]
    assert np.array_equal(pr, np.where(succ == 1, 1.0, 0.1))                              # buffer.py:21-29
    assert ag.buffer.memory[101].step_success == -1


def test_device_append_bookkeeping_is_bounded_and_respects_later_host_stores():
    """a long act -> env -> store_batch_dev -> learn loop must not accumulate per-append state, a ring
    wrap must not replay stale ranges, and a host store() issued AFTER a device append into the same slot must survive the
    lazy mirror refresh."""
    sc = scenarios.make_scenario(40, 400)
    ag = scenarios.build_agent(Agent, sc, 16, 0, hp={"bufferSize": 64})
    buf = ag.buffer
    d = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    n = 24
    for it in range(40):                                    # 960 device-appended rows through a 64-slot ring
        s, a, s2 = (torch.from_numpy(x.astype(np.float32)).to(d) for x in (synth.env_states(rng, n), synth.env_actions(rng, n), synth.env_states(rng, n)))
        r = torch.full((n,), float(it), device=d)
        fl = torch.full((n,), ((it % 3 - 1) & 0xFF) << 8, dtype=torch.int32, device=d)
        ag.store_batch_dev(s, a, s2, r, fl)
        if it % 10 == 9:
            assert all(np.isfinite(float(x)) for x in ag.learn(0.5, 0, 0))
    assert not hasattr(buf, "_dev_ranges") and buf._n_stale <= buf.capacity
    assert buf._rows_stale.shape == (64,) and buf._dev_success.shape == (64,)
    pos = buf.position
    last = (pos - 1) % 64
    # newest row: iteration 39; the slot 24 places behind it: iteration 38; 48 behind: iteration 37 (partly overwritten ring)
    ag.store(sc["s"][1], sc["a"][1], sc["s2"][1], -5.0, True, 1)          # host store lands at `pos`, over a device-appended slot
    assert buf.memory[last].reward == 39.0 and buf.memory[last].step_success == (39 % 3 - 1)
    assert buf.memory[(last - 24) % 64].reward == 38.0 and buf.memory[(last - 24) % 64].step_success == (38 % 3 - 1)
    tr = buf.memory[pos]
    assert tr.reward == -5.0 and tr.done is True and tr.step_success == 1   # not clobbered by the device gather
    assert buf._n_stale == 0
    # and the device sees the host row too
    got = np.empty((1, 32), np.float32)
    idx = np.array([pos], np.int32)
    buf.flush()
    torch.cuda.synchronize()
    _lib.check(_lib.load().hirl_gather(ag._h, 0, _lib.ptr(idx), 1, None, 0, _lib.ptr(got)))
    assert got[0, 30] == -5.0 and got[0, 31] == 1.0

"""CPU: host-side mirrors of the reference interface that need no GPU."""
import random

import numpy as np
import pytest
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


# ---- exact fast index draws + speculative pre-drawing (utils/fast_rng.py, csrc/host_rng.h) ---------------------------
def test_fast_draws_equal_the_interpreter_level_originals_and_leave_the_same_state():
    import random
    from hirl4ucav_b200.utils import fast_rng as fr
    assert fr.py_fast() and fr.np_fast(), "generator state access self-check failed on this interpreter / numpy"
    rng = np.random.default_rng(0)
    for seed in range(40):
        n1, n2, nb = int(rng.integers(1, 40000)), int(rng.integers(1, 3000)), int(rng.integers(1, 25000))
        k1, k2, kb = int(rng.integers(0, min(n1, 4096) + 1)), int(rng.integers(0, min(n2, 300) + 1)), int(rng.integers(0, min(nb, 4096) + 1))
        random.seed(seed); np.random.seed(seed)
        want = (random.sample(range(n1), k1), random.sample(range(n2), k2), np.random.choice(nb, kb, replace=False),
                random.random(), random.getrandbits(17), np.random.rand(), np.random.randn(), np.random.randint(1 << 20))
        random.seed(seed); np.random.seed(seed)
        got = (fr.sample_indices(n1, k1), fr.sample_indices(n2, k2), fr.choice_noreplace(nb, kb),
               random.random(), random.getrandbits(17), np.random.rand(), np.random.randn(), np.random.randint(1 << 20))
        assert want[0] == got[0] and want[1] == got[1] and np.array_equal(want[2], got[2]), (seed, n1, k1, n2, k2, nb, kb)
        assert want[3:] == got[3:], "generator state after the draws differs"
    # both branches of Random.sample around the set/pool threshold (k=192: setsize 21 + 4**5 = 1045)
    for n in (192, 193, 1044, 1045, 1046, 5000):
        random.seed(n); a = random.sample(range(n), 192); random.seed(n); b = fr.sample_indices(n, 192)
        assert a == b
    with pytest.raises(ValueError):
        fr.sample_indices(10, 11)
    with pytest.raises(ValueError):
        fr.choice_noreplace(10, 11)


def test_predraw_is_exact_under_reseeding_foreign_draws_and_wrong_guesses():
    import random
    from hirl4ucav_b200.utils import fast_rng as fr

    def reference(script):
        out = []
        for op in script:
            if op[0] == "seed":
                random.seed(op[1]); np.random.seed(op[1])
            elif op[0] == "foreign":
                out.append((random.randint(-100, 100), float(np.random.rand())))     # e.g. env reset (GYM.py:74)
            else:
                _, n, k, nb, kb = op
                out.append((random.sample(range(n), k), np.random.choice(nb, kb, replace=False).tolist()))
        return out

    def with_predraw(script):
        py, npd = fr.Predraw("py"), fr.Predraw("np")
        out = []
        last = None
        for op in script:
            if op[0] == "seed":
                random.seed(op[1]); np.random.seed(op[1])
            elif op[0] == "foreign":
                out.append((random.randint(-100, 100), float(np.random.rand())))
            else:
                _, n, k, nb, kb = op
                r = py.take([(n, k)])
                s = r[0].tolist() if r is not None else fr.sample_indices(n, k)
                r = npd.take([(nb, kb)])
                c = r[0].tolist() if r is not None else fr.choice_noreplace(nb, kb).tolist()
                out.append((s, c))
                last = op
                py.prepare([(n + 1, k)])          # guess: the ring grows by one row per call
                npd.prepare([(nb, kb)])
        return out, py, npd

    script = [("seed", 1)] + [("learn", 5000 + i, 64, 2000, 32) for i in range(6)]          # guesses right: hits
    script += [("foreign",), ("learn", 5006, 64, 2000, 32)]                                  # foreign draw: discard
    script += [("seed", 9), ("learn", 5007, 64, 2000, 32), ("learn", 5007, 64, 2000, 32)]    # reseed, then a wrong size guess
    script += [("learn", 5008, 48, 2000, 32), ("learn", 5009, 48, 2000, 16)]                 # argument changes
    want = reference(script)
    got, py, npd = with_predraw(script)
    assert want == got
    assert py.hits >= 5 and py.misses >= 4 and npd.hits >= 6


def test_threaded_ensemble_draws_equal_per_replica_sequential_draws():
    """hirl_host_draw_many: the whole sweep of index draws of R replicas on worker threads == the reference's draws made
    replica by replica on private generators (random.Random / np.random.RandomState), for any thread count, and the
    generators end in the same state."""
    import ctypes as C
    import random
    from hirl4ucav_b200 import _lib
    from hirl4ucav_b200.utils import fast_rng as fr
    assert fr.py_fast() and fr.np_fast()
    lib = _lib.load()
    R, K, B, n_bc = 9, 5, 64, 3000
    split = [(48, 16), (64, 0), (40, 24), (48, 16), (63, 1)]
    lens = [(500 + 37 * r, 100 + r) for r in range(R)]
    want_on = np.zeros((R, K, B), np.int64); want_ex = np.zeros((R, K, B), np.int64); want_bc = np.zeros((R, K, B), np.int64)
    tails = []
    for r in range(R):
        py, rs = random.Random(100 + r), np.random.RandomState(100 + r)
        for k, (n_on, n_ex) in enumerate(split):
            want_on[r, k, :n_on] = py.sample(range(lens[r][0]), n_on)
            if n_ex:
                want_ex[r, k, :n_ex] = py.sample(range(lens[r][1]), n_ex)
            want_bc[r, k] = rs.choice(n_bc, B, replace=False)
        tails.append((py.random(), rs.rand()))
    for threads in (1, 4, 32):
        pys = [random.Random(100 + r) for r in range(R)]
        rss = [np.random.RandomState(100 + r) for r in range(R)]
        states = (fr.MTState * R)()
        jobs = (_lib.DrawJob * R)()
        for r in range(R):
            live = fr._py_live(pys[r])
            C.memmove(states[r].key, live.state, 624 * 4)
            states[r].pos = live.index
            jobs[r].py, jobs[r].np = C.addressof(states[r]), rss[r]._bit_generator.ctypes.state_address
            jobs[r].online_len, jobs[r].expert_len = lens[r]
        n_on = np.array([s[0] for s in split], np.int32); n_ex = np.array([s[1] for s in split], np.int32)
        o = np.zeros((R, K, B), np.int32); e = np.zeros((R, K, B), np.int32); b = np.zeros((R, K, B), np.int32)
        _lib.check(lib.hirl_host_draw_many(jobs, R, K, _lib.ptr(n_on), _lib.ptr(n_ex), n_bc, B, _lib.ptr(o), _lib.ptr(e), _lib.ptr(b), threads))
        for r in range(R):
            fr._py_write(states[r], pys[r])
            for k, (a, c) in enumerate(split):
                assert np.array_equal(o[r, k, :a], want_on[r, k, :a]) and np.array_equal(e[r, k, :c], want_ex[r, k, :c])
            assert np.array_equal(b[r], want_bc[r])
            assert (pys[r].random(), rss[r].rand()) == tails[r], "generator state after the sweep differs"
    bad = np.array([600] + [0] * (K - 1), np.int32)     # sample larger than the population -> ValueError like random.sample
    with pytest.raises(ValueError):
        _lib.check(lib.hirl_host_draw_many(jobs, R, K, _lib.ptr(bad), _lib.ptr(n_ex), n_bc, B, _lib.ptr(o), _lib.ptr(e), _lib.ptr(b), 2))


def test_noise_draw_shortcut_is_the_reference_stream():
    """torch.normal(mean=zeros(4), std=ones(4) * 0.2) (HIRL.py:265) vs the in-place form the agents use (buf.normal_() * std + mean):
    bit-identical values and generator state, on the global generator and on a private one."""
    mean, std = torch.zeros(4), torch.ones(4) * 0.2
    for gen in (None, torch.Generator()):
        kw = {} if gen is None else {"generator": gen}
        (torch.manual_seed if gen is None else gen.manual_seed)(11)
        want = [torch.normal(mean=mean, std=std, **kw).numpy().copy() for _ in range(200)]
        tail_w = torch.rand(1, **kw).item()
        (torch.manual_seed if gen is None else gen.manual_seed)(11)
        buf = torch.empty(4)
        got = []
        for _ in range(200):
            buf.normal_(**kw)
            got.append(buf.numpy() * std.numpy() + mean.numpy())
        tail_g = torch.rand(1, **kw).item()
        assert all(np.array_equal(a, b) for a, b in zip(want, got)) and tail_w == tail_g


def test_predraw_in_flight_survives_its_owner():
    """An agent can be garbage-collected while its np pre-draw still runs on the library's worker thread: the job's buffers
    are held at module level and the destructor joins the worker (the worker must never write into freed memory)."""
    import gc
    from hirl4ucav_b200 import _lib
    from hirl4ucav_b200.utils import fast_rng as fr
    assert fr.np_fast()
    np.random.seed(5)
    want = np.random.choice(20000, 256, replace=False)
    np.random.seed(5)
    for _ in range(20):
        p = fr.Predraw("np")
        p.prepare([(20000, 256)])
        del p                                   # owner gone while the shuffle may still be running
        gc.collect()
        assert fr._INFLIGHT is None             # joined by the destructor
    q = fr.Predraw("np")
    q.prepare([(20000, 256)])
    got = q.take([(20000, 256)])                # the live state was never touched by the abandoned jobs
    assert got is not None and np.array_equal(got[0], want)
    assert _lib.load().hirl_host_async_wait() == 0

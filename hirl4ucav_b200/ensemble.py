"""R independent HIRL agents ("replicas", BASELINE.json config 4) updated by ONE persistent launch.

Each replica is a complete agent of the reference (hirl/agents/HIRL.py:148-190): its own five
networks, Adam state, replay rings and its own three RNG streams, seeded exactly like a single run
with ``set_seed(seed)`` (hirl/utils/seed.py:5-10) -- replica i has the weights of an ``Agent`` constructed
under that seed and, after ``reseed()``, consumes the same three streams (see ``_Streams``).  Only the read-only
expert (s, a) table is shared.  On the device the grid is cut
into teams of CTAs; each team runs whole replicas with team-local barriers
(csrc/hirl_b200.cu: hirl_ensemble_kernel).  No collective: replicas shard over GPUs with
``parallel.shard_range``.
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from . import _lib, nets
from .utils import fast_rng
from .agents.HIRL import SOFT_SENTINEL, _adam_scalars


class _Streams:
    """The reference's three host RNG streams for one replica: private generators seeded like set_seed(seed).  The py / np
    streams reproduce a single run's exactly; the torch stream is a FRESH generator, whereas a single agent's global torch
    generator has already been advanced by the weight initialisation -- call ``AgentEnsemble.reseed()`` (what a run does with
    set_seed before its update loop, and what the parity tests do) to put all three streams in a single run's state."""

    def __init__(self, seed):
        self.py = random.Random(seed)
        self.np = np.random.RandomState(seed)
        self.torch = torch.Generator().manual_seed(seed)

    def reseed(self, seed):
        self.py.seed(seed)
        self.np.seed(seed)
        self.torch.manual_seed(seed)


class AgentEnsemble(object):
    def __init__(self, seeds, actorLR, criticLR, tau, gamma, bufferSize, batchSize, expert_states, expert_actions,
                 expert_warm_up=True, name="Harfang_GYM", device=None, max_fused_steps=64):
        self.R = len(seeds)
        self.seeds = list(seeds)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.batchSize, self.bufferSize = int(batchSize), int(bufferSize)
        self.actorLR, self.criticLR, self.tau, self.gamma = actorLR, criticLR, tau, gamma
        self.expert_warm_up = expert_warm_up
        self.loss_lambda = 10000
        self.target_update_freq = 3
        self.actionDim = 4
        self.TD3LearningNoise, self.TD3LearningNoiseClamp = 0.2, 0.5
        self.actorTrainable = True
        self.update_count = 0
        self._critic_t = self._actor_t = 0
        self._max_steps = int(max_fused_steps)
        self.n_expert = len(expert_states)
        cfg = _lib.Config(device=self.device.index, batch=self.batchSize, online_capacity=self.bufferSize,
                          expert_capacity=self.n_expert + 10, bc_rows=self.n_expert, max_steps=self._max_steps,
                          activation=_lib.ACT_RELU, replicas=self.R, gamma=gamma, tau=tau,
                          noise_clamp=self.TD3LearningNoiseClamp, loss_lambda=float(self.loss_lambda))
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.hirl_agent_create(C.byref(cfg), C.byref(self._h)))
        es, ea = _lib.f32c(expert_states), _lib.f32c(expert_actions)
        _lib.check(self._lib.hirl_bc_table_set(self._h, _lib.ptr(es), _lib.ptr(ea), self.n_expert))
        self.streams = []
        self.online_len = [0] * self.R
        self.expert_len = [0] * self.R
        self._online_pos = [0] * self.R      # ring write positions (buffer.py:12,31-36: position wraps, len saturates)
        self._expert_pos = [0] * self.R
        for r, seed in enumerate(self.seeds):
            # initial weights: the reference's construction order under torch.manual_seed(seed)
            with torch.random.fork_rng(devices=[]):
                torch.manual_seed(seed)
                actor = nets.make_actor(13, 4, 256, 512, "Actor_" + name)
                nets.make_actor(13, 4, 256, 512, "TargetActor_" + name)      # drawn, then overwritten (hard_update)
                critic = nets.make_critic(13, 4, 256, 512, "Critic_" + name)
                nets.make_critic(13, 4, 256, 512, "TargetCritic_" + name)
                bc = nets.make_actor(13, 4, 256, 512, name)
            for net_id, mod in ((_lib.NET_ACTOR, actor), (_lib.NET_TARGET_ACTOR, actor), (_lib.NET_CRITIC, critic),
                                (_lib.NET_TARGET_CRITIC, critic), (_lib.NET_BC_ACTOR, bc)):
                v = mod.flat()
                _lib.check(self._lib.hirl_params_import(self._h, r, net_id, _lib.ptr(v), v.size))
            self.streams.append(_Streams(seed))
        K, B, R = self._max_steps, self.batchSize, self.R
        pin = torch.cuda.is_available()
        mk = lambda shape, dt: torch.zeros(shape, dtype=dt, pin_memory=pin)
        self._steps_t = mk((R * K * _lib.STEP_DTYPE.itemsize,), torch.uint8)
        self._on_t, self._ex_t, self._bc_t = mk((R, K, B), torch.int32), mk((R, K, B), torch.int32), mk((R, K, B), torch.int32)
        self._out_t = mk((R, K, _lib.OUT_SCALARS), torch.float32)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.hirl_agent_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- data ---------------------------------------------------------------------------------
    def store_rows(self, replica, rows32, expert=False):
        """Appends pre-packed 32-word rows to one replica's online / expert ring at its write position (buffer.py:31-36:
        the position wraps around the capacity, the length saturates)."""
        rows32 = _lib.f32c(rows32).reshape(-1, 32)
        cap = self.n_expert + 10 if expert else self.bufferSize
        lens, pos = (self.expert_len, self._expert_pos) if expert else (self.online_len, self._online_pos)
        which = _lib.BUF_EXPERT if expert else _lib.BUF_ONLINE
        rows32 = rows32[-cap:] if rows32.shape[0] > cap else rows32           # only the last `cap` rows survive a longer append
        n, first = rows32.shape[0], pos[replica]
        head = min(n, cap - first)
        _lib.check(self._lib.hirl_store_rows(self._h, replica, which, first, head, _lib.ptr(np.ascontiguousarray(rows32[:head]))))
        if n > head:                                                           # wrapped: the rest goes to the start of the ring
            _lib.check(self._lib.hirl_store_rows(self._h, replica, which, 0, n - head, _lib.ptr(np.ascontiguousarray(rows32[head:]))))
        pos[replica] = (first + n) % cap
        lens[replica] = min(cap, lens[replica] + n)

    def reseed(self, offset=1):
        """set_seed(seed + offset) for every replica's streams (what bench/tests do before the timed loop)."""
        for s, seed in zip(self.streams, self.seeds):
            s.reseed(seed + offset)

    def export_flat(self, replica, net_id):
        n = _lib.CRITIC_PARAMS if net_id in (_lib.NET_CRITIC, _lib.NET_TARGET_CRITIC) else _lib.ACTOR_PARAMS
        out = np.empty(n, np.float32)
        _lib.check(self._lib.hirl_params_export(self._h, replica, net_id, _lib.ptr(out), n))
        return out

    # ---- update -------------------------------------------------------------------------------
    def draw(self, calls):
        """Pre-draws the index / noise streams of len(calls) consecutive learn() calls for every replica
        (reference order per call: online sample, expert sample, BC choice, noise)."""
        K, B, R = len(calls), self.batchSize, self.R
        if K > self._max_steps:
            raise ValueError("more calls than max_fused_steps")
        steps = self._steps_t.numpy().view(_lib.STEP_DTYPE)[: R * K].reshape(R, K)
        on, ex, bc = self._on_t.numpy(), self._ex_t.numpy(), self._bc_t.numpy()
        flags = []
        trainable, upd, ct, at = self.actorTrainable, self.update_count, self._critic_t, self._actor_t
        sched = []
        for _ in range(K):
            do_actor = trainable
            ct += 1
            if do_actor:
                at += 1
                upd += 1
            sched.append((do_actor, do_actor and upd % self.target_update_freq == 0, _adam_scalars(self.criticLR, ct),
                          _adam_scalars(self.actorLR, at) if do_actor else (0.0, 0.0)))
            trainable = not trainable
        self.actorTrainable, self.update_count, self._critic_t, self._actor_t = trainable, upd, ct, at
        mean, std = torch.zeros(self.actionDim), torch.ones(self.actionDim) * self.TD3LearningNoise
        split = []
        for bcw, en, warm in calls:
            n_ex = int(en) if (self.expert_warm_up and en) else 0
            if B - n_ex <= 0:
                raise TypeError("expert_num_now == batchSize leaves no online rows")
            split.append((B - n_ex, n_ex))
        if fast_rng.py_fast() and fast_rng.np_fast():
            # every index draw of the sweep in ONE library call, replicas spread over worker threads (their generators are
            # independent: exact C replicas of the reference's draws on each replica's private state, utils/fast_rng.py)
            self._draw_indices_threaded(split, on, ex, bc)
        else:
            for r in range(R):
                st = self.streams[r]
                for k, (n_on, n_ex) in enumerate(split):
                    on[r, k, :n_on] = fast_rng.sample_indices_on(st.py, self.online_len[r], n_on)
                    if n_ex:
                        ex[r, k, :n_ex] = fast_rng.sample_indices_on(st.py, self.expert_len[r], n_ex)
                    bc[r, k, :] = fast_rng.choice_noreplace_on(st.np, self.n_expert, B)
        # the step records: the schedule is the same for every replica, only the noise differs
        i32 = self._steps_t.numpy().view(np.int32)[: R * K * 16].reshape(R, K, 16)
        f32 = self._steps_t.numpy().view(np.float32)[: R * K * 16].reshape(R, K, 16)
        for k, ((bcw, en, warm), (n_on, n_ex)) in enumerate(zip(calls, split)):
            i32[:, k, 0:4] = (n_on, n_ex, int(sched[k][0]), int(sched[k][1]))
            f32[:, k, 4:7] = (bcw, warm, -1.0)
            f32[:, k, 11:15] = sched[k][2] + sched[k][3]
            i32[:, k, 15] = 0
        buf = torch.empty(self.actionDim)      # torch.normal(mean, std) == out.normal_(0, 1).mul_(std).add_(mean): same stream, 6x cheaper
        bn, sn, mn = buf.numpy(), std.numpy(), mean.numpy()
        for r in range(R):
            g = self.streams[r].torch
            for k in range(K):
                buf.normal_(generator=g)
                f32[r, k, 7:11] = bn * sn + mn
        return K

    def _draw_indices_threaded(self, split, on, ex, bc):
        import os
        R, K, B = self.R, len(split), self.batchSize
        Kmax = on.shape[1]
        py_states = (fast_rng.MTState * R)()
        jobs = (_lib.DrawJob * R)()
        for r in range(R):
            st = self.streams[r]
            live = fast_rng._py_live(st.py)
            C.memmove(py_states[r].key, live.state, 624 * 4)
            py_states[r].pos = live.index
            jobs[r].py = C.addressof(py_states[r])
            jobs[r].np = st.np._bit_generator.ctypes.state_address
            jobs[r].online_len, jobs[r].expert_len = int(self.online_len[r]), int(self.expert_len[r])
        n_on = np.array([s_[0] for s_ in split], np.int32)
        n_ex = np.array([s_[1] for s_ in split], np.int32)
        o = np.empty((R, K, B), np.int32); e = np.empty((R, K, B), np.int32); b = np.empty((R, K, B), np.int32)
        _lib.check(self._lib.hirl_host_draw_many(jobs, R, K, _lib.ptr(n_on), _lib.ptr(n_ex), self.n_expert, B, _lib.ptr(o), _lib.ptr(e),
                                                 _lib.ptr(b), min(32, os.cpu_count() or 1)))
        for r in range(R):
            fast_rng._py_write(py_states[r], self.streams[r].py)
        on[:, :K] = o
        ex[:, :K] = e
        bc[:, :K] = b

    def launch(self, K):
        """hirl_learn over the staged streams: [R][K] everything, host buffers."""
        R, B = self.R, self.batchSize
        st = np.ascontiguousarray(self._steps_t.numpy().view(_lib.STEP_DTYPE)[: R * K])
        on = np.ascontiguousarray(self._on_t.numpy()[:, :K]); ex = np.ascontiguousarray(self._ex_t.numpy()[:, :K])
        bc = np.ascontiguousarray(self._bc_t.numpy()[:, :K])
        out = np.empty((R, K, _lib.OUT_SCALARS), np.float32)
        _lib.check(self._lib.hirl_learn(self._h, K, st.ctypes.data_as(C.c_void_p), _lib.ptr(on), _lib.ptr(ex), _lib.ptr(bc),
                                        _lib.ptr(out)))
        return out

    def learn_many(self, calls):
        """All replicas execute the same call sequence; returns scalars [R][K][8] (critic_loss, actor_loss,
        bc_loss, rl_loss, fire_mse, bc_weight, soft_count, -)."""
        K = self.draw(list(calls))
        return self.launch(K)

    def device_streams(self, K):
        """Device copies of the staged streams + a launcher (bench: inputs resident in HBM)."""
        R = self.R
        dev = self.device
        st = torch.from_numpy(np.ascontiguousarray(self._steps_t.numpy()[: R * K * 64])).to(dev)
        on = self._on_t[:, :K].contiguous().to(dev); ex = self._ex_t[:, :K].contiguous().to(dev); bc = self._bc_t[:, :K].contiguous().to(dev)
        out = torch.zeros((R, K, 8), dtype=torch.float32, device=dev)

        def launch():
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(self._lib.hirl_learn_dev(self._h, K, st.data_ptr(), on.data_ptr(), ex.data_ptr(), bc.data_ptr(), out.data_ptr(), stream))
        launch.keep = (st, on, ex, bc, out)
        return launch, out

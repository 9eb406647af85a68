"""Drop-in for ``hirl.agents.HIRL.Agent`` (reference hirl/agents/HIRL.py:148-350) whose update
runs as ONE persistent fused sm_100a kernel (csrc/update.cuh) behind the C ABI.

What stays on the host, exactly as in the reference: the three RNG streams
(``random.sample`` for replay indices, buffer.py:45; ``np.random.choice`` for the BC batch,
HIRL.py:249; ``torch.normal`` for the target-policy noise, HIRL.py:265), the
actorTrainable / update_count schedule (HIRL.py:291, 327-332) and Adam's bias corrections
(computed in Python floats like torch.optim.Adam).  Everything numerical is on the GPU; there
is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .. import _lib, nets
from ..utils import fast_rng
from ..utils.buffer import UniformMemory

SOFT_SENTINEL = 100
_BETA1, _BETA2 = 0.9, 0.999


def _adam_scalars(lr: float, step: int):
    """-step_size and sqrt(bias_correction2) as torch's single-tensor Adam computes them."""
    bc1 = 1 - _BETA1 ** step
    bc2 = 1 - _BETA2 ** step
    return -(lr / bc1), bc2 ** 0.5


_POOL = None


def _draw_pool():
    """One helper thread for learn_many's np.random draws (created on first use)."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=1, thread_name_prefix="hirl-draw")
    return _POOL


class Agent(object):
    _MODE = 0            # 0 HIRL, 1 TD3 (see agents/TD3.py)
    _GAIN = "relu"
    _ACT = _lib.ACT_RELU
    # HIRL.py:184-185 hard-codes both switches to False inside the constructor; a maintainer who flips those
    # source lines flips these class attributes instead (or passes the keyword arguments).
    BUFFER_UPSAMPLE = False   # online ring: weighted draw with replacement (buffer.py:21-29, 39-43)
    EXPERT_UPSAMPLE = False   # cloning batch: one row replaced by a firing expert row (HIRL.py:253-257)
    FUSE_SUB = 8              # learn_many: calls per queued launch (host draws of the next ones overlap the GPU)
    PREDRAW = True            # overlap the next call's host index draws with the GPU (exact: utils/fast_rng.Predraw)

    def __init__(self, actorLR, criticLR, stateDim, actionDim, full1Dim, full2Dim, tau, gamma, bufferSize, batchSize,
                 layerNorm, name, expert_states, expert_actions, bc_weight, expert_warm_up, device=None,
                 max_fused_steps=64, buffer_upsample=None, expert_upsample=None, data_parallel=None):
        """``data_parallel=(rank, world)``: data-parallel large-batch mode (include/hirl_b200.h).  Every rank
        constructs the same agent under the same seeds and stores the same transitions; ``batchSize`` is the GLOBAL batch,
        this process executes ``batchSize // world`` of every call's rows and the gradients are all-reduced (NCCL) at the
        reference's two optimizer steps (HIRL.py:286-288, 323-325).  Needs an initialised torch.distributed group for
        world > 1 (only to broadcast the 128-byte NCCL id)."""
        if (stateDim, actionDim, full1Dim, full2Dim) != (13, 4, 256, 512):
            raise NotImplementedError("the fused kernel is specialised for the reference driver's geometry "
                                      "13/4/256/512 (hirl/train_all.py:200-203)")
        if not layerNorm:
            raise NotImplementedError("layerNorm=False is outside the fused path (train_all.py:204 hard-codes True)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.tau, self.gamma = tau, gamma
        self.bufferSize, self.batchSize = int(bufferSize), int(batchSize)
        self._dp = None if data_parallel is None else (int(data_parallel[0]), int(data_parallel[1]))
        if self._dp is not None and (self.batchSize % self._dp[1] or not 0 <= self._dp[0] < self._dp[1]):
            raise ValueError("data_parallel=(rank, world): batchSize must be a multiple of world and 0 <= rank < world")
        self._dev_batch = self.batchSize if self._dp is None else self.batchSize // self._dp[1]   # rows this process executes per call
        self.actorLR, self.criticLR = actorLR, criticLR
        self.actorTrainable = True
        self.actionDim = actionDim
        self.actionNoise = 0.1
        self.TD3LearningNoise = 0.2
        self.TD3LearningNoiseClamp = 0.5
        self.expert_states = expert_states
        self.expert_actions = expert_actions
        self.bc_weight = bc_weight
        self.expert_warm_up = expert_warm_up
        self.update_count = 0
        self.loss_lambda = 10000
        self.target_update_freq = 3
        self.buffer_upsample = bool(self.BUFFER_UPSAMPLE if buffer_upsample is None else buffer_upsample)
        self.expert_upsample = bool(self.EXPERT_UPSAMPLE if expert_upsample is None else expert_upsample)
        self._pre_py, self._pre_np = fast_rng.Predraw("py"), fast_rng.Predraw("np")
        self._n_online_prev = 0
        self._critic_t = 0      # Adam step counters (torch keeps them in optimizer.state)
        self._actor_t = 0
        self._max_steps = int(max_fused_steps)
        has_expert = expert_states is not None and len(expert_states) > 0
        if has_expert:
            self.target_indices = np.where(np.asarray(expert_actions)[:, 3] == 1)[0]
            print(f"upsample expert data num is {np.size(self.target_indices)}")

        # host parameter containers, created in the reference's order (RNG stream!)
        self._make_nets(stateDim, actionDim, full1Dim, full2Dim, name)
        if self._MODE == 0:
            print(f"buffer upsample is {self.buffer_upsample}, expert upsample is {self.expert_upsample}, "
                  f"loss lambda is {self.loss_lambda}")

        n_exp = len(expert_states) if has_expert else 0
        cfg = _lib.Config(device=self.device.index, batch=self._dev_batch, online_capacity=self.bufferSize,
                          expert_capacity=n_exp + 10, bc_rows=n_exp, max_steps=self._max_steps, activation=self._ACT,
                          replicas=1, gamma=gamma, tau=tau, noise_clamp=self.TD3LearningNoiseClamp,
                          loss_lambda=float(self.loss_lambda))
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.hirl_agent_create(C.byref(cfg), C.byref(self._h)))
        self._nets = {_lib.NET_ACTOR: self.actor, _lib.NET_TARGET_ACTOR: self.targetActor, _lib.NET_CRITIC: self.critic,
                      _lib.NET_TARGET_CRITIC: self.targetCritic}
        if self.bc_actor is not None:
            self._nets[_lib.NET_BC_ACTOR] = self.bc_actor
        self._host_stale = False
        self.push_parameters()
        if has_expert:
            es, ea = _lib.f32c(expert_states), _lib.f32c(expert_actions)
            _lib.check(self._lib.hirl_bc_table_set(self._h, _lib.ptr(es), _lib.ptr(ea), n_exp))

        self.buffer = UniformMemory(self.bufferSize, self.buffer_upsample)
        self.buffer._attach(self._h, _lib.BUF_ONLINE)
        self.expert_buffer = UniformMemory(n_exp + 10, False)
        self.expert_buffer._attach(self._h, _lib.BUF_EXPERT)

        # pinned staging for the step streams
        K, B = self._max_steps, self.batchSize
        pin = torch.cuda.is_available()
        mk = lambda shape, dt: torch.zeros(shape, dtype=dt, pin_memory=pin)
        self._st_steps_t = mk((K * _lib.STEP_DTYPE.itemsize,), torch.uint8)
        self._st_steps = self._st_steps_t.numpy().view(_lib.STEP_DTYPE)
        self._st_i32 = self._st_steps_t.numpy().view(np.int32).reshape(K, 16)      # the same records, word by word (hirl_step_t)
        self._st_f32 = self._st_steps_t.numpy().view(np.float32).reshape(K, 16)
        self._ld = self._ld_cache = None
        self._st_on_t, self._st_ex_t, self._st_bc_t = mk((K, B), torch.int32), mk((K, B), torch.int32), mk((K, B), torch.int32)
        self._st_on, self._st_ex, self._st_bc = self._st_on_t.numpy(), self._st_ex_t.numpy(), self._st_bc_t.numpy()
        self._st_out_t = mk((K, _lib.OUT_SCALARS), torch.float32)
        self._st_out = self._st_out_t.numpy()
        self._st_ptrs = (self._st_steps.ctypes.data_as(C.c_void_p), _lib.ptr(self._st_on), _lib.ptr(self._st_ex),
                         _lib.ptr(self._st_bc), _lib.ptr(self._st_out))   # the staging arrays never move
        self._st_addr = tuple(a.ctypes.data for a in (self._st_steps, self._st_on, self._st_ex, self._st_bc, self._st_out))
        self._ret = (np.float32(0), np.float32(0), np.float32(0), np.float32(0), 0.0, bc_weight)
        self._noise_mean = torch.zeros(self.actionDim)
        self._noise_std = torch.ones(self.actionDim) * self.TD3LearningNoise
        self._noise_buf = torch.empty(self.actionDim)
        self._noise_np, self._noise_mean_np, self._noise_std_np = self._noise_buf.numpy(), self._noise_mean.numpy(), self._noise_std.numpy()
        self.last_draws = {}
        if self._dp is not None:
            Bl = self._dev_batch
            self._sh_steps_t = mk((K * _lib.STEP_DTYPE.itemsize,), torch.uint8)
            self._sh_steps = self._sh_steps_t.numpy().view(_lib.STEP_DTYPE)
            self._sh_on_t, self._sh_ex_t, self._sh_bc_t = mk((K, Bl), torch.int32), mk((K, Bl), torch.int32), mk((K, Bl), torch.int32)
            self._sh_on, self._sh_ex, self._sh_bc = self._sh_on_t.numpy(), self._sh_ex_t.numpy(), self._sh_bc_t.numpy()
            self._sh_addr = tuple(a.ctypes.data for a in (self._sh_steps, self._sh_on, self._sh_ex, self._sh_bc))
            self._attach_data_parallel()

    def _attach_data_parallel(self):
        """Fused variant first (all-reduce inside the kernel over IPC-mapped NVLink peer memory: a call stays one launch);
        the NCCL variant (stage-range launches + ncclAllReduce) if CUDA IPC is unavailable or HIRL_DP_NCCL=1."""
        rank, world = self._dp
        self.dp_variant = None
        if os.environ.get("HIRL_DP_NCCL") != "1":
            h64 = np.zeros(64, np.uint8)
            rc = self._lib.hirl_dp_fused_export(self._h, rank, world, self.batchSize, _lib.ptr(h64))
            ok = rc == 0 and os.environ.get("HIRL_DP_FUSED_FAIL") != "1"      # (test hook: pretend CUDA IPC is unavailable)
            handles = h64
            if world > 1:
                import torch.distributed as dist
                box = [None] * world
                dist.all_gather_object(box, (ok, h64.tobytes()))
                ok = all(b[0] for b in box)
                handles = np.frombuffer(b"".join(b[1] for b in box), np.uint8).copy()
            if ok:
                ok = self._lib.hirl_dp_fused_attach(self._h, _lib.ptr(handles), world) == 0
                if world > 1:
                    box = [None] * world
                    dist.all_gather_object(box, ok)
                    ok = all(box)
            if ok:
                self.dp_variant = "fused"
                return
            # every rank saw the same verdict (all-gathered above), so they all fall back together
            import warnings
            warnings.warn("fused data-parallel attach failed on some rank (" + self._lib.hirl_last_error().decode(errors="replace")
                          + "): falling back to the NCCL variant")
        uid = np.zeros(128, np.uint8)
        if rank == 0:
            _lib.check(self._lib.hirl_nccl_unique_id(_lib.ptr(uid)))
        if world > 1:
            import torch.distributed as dist
            box = [uid.tobytes()]
            dist.broadcast_object_list(box, src=0)
            uid = np.frombuffer(box[0], np.uint8).copy()
        _lib.check(self._lib.hirl_allreduce_init(self._h, _lib.ptr(uid), rank, world, self.batchSize))
        self.dp_variant = "nccl"

    @property
    def backend(self):
        """Name of the update kernel this agent launches."""
        return ("stage", "tcc", "cluster", "ensemble")[self._lib.hirl_agent_backend(self._h)]

    def _shard_slots(self, lo, n):
        """Cuts staging slots lo .. lo+n-1 (drawn over the global batch) down to this rank's rows (parallel.shard_step)."""
        from .. import parallel
        rank, world = self._dp
        for k in range(lo, lo + n):
            st = self._st_steps[k]
            n_on, n_ex, on, ex, bc = parallel.shard_step(int(st["n_online"]), int(st["n_expert"]), self._st_on[k], self._st_ex[k],
                                                         self._st_bc[k], rank, world)
            self._sh_steps[k] = st
            self._sh_steps[k]["n_online"], self._sh_steps[k]["n_expert"] = n_on, n_ex
            self._sh_on[k, :n_on] = on
            self._sh_ex[k, :n_ex] = ex
            self._sh_bc[k, :] = bc

    def _submit(self, lo, n):
        """Enqueues staging slots lo .. lo+n-1.  Data-parallel agents first cut every slot's index streams (drawn over the
        global batch, identically on every rank) down to this rank's rows (parallel.shard_step)."""
        if self._dp is None:
            _lib.check(self._lib.hirl_learn_async(self._h, n, *(self._st_ptrs if lo == 0 else self._slot_ptrs(lo))))
            return
        self._shard_slots(lo, n)
        Bl = self._dev_batch
        a = self._sh_addr
        _lib.check(self._lib.hirl_learn_async(self._h, n, C.c_void_p(a[0] + lo * _lib.STEP_DTYPE.itemsize), C.c_void_p(a[1] + lo * Bl * 4),
                                              C.c_void_p(a[2] + lo * Bl * 4), C.c_void_p(a[3] + lo * Bl * 4),
                                              C.c_void_p(self._st_addr[4] + lo * _lib.OUT_SCALARS * 4)))

    def _make_nets(self, stateDim, actionDim, full1Dim, full2Dim, name):
        mk_a = lambda tag: nets.make_actor(stateDim, actionDim, full1Dim, full2Dim, tag + name, self._GAIN)
        mk_c = lambda tag: nets.make_critic(stateDim, actionDim, full1Dim, full2Dim, tag + name, self._GAIN)
        self.actor = mk_a("Actor_")
        self.targetActor = mk_a("TargetActor_")
        self.targetActor.load_state_dict(self.actor.state_dict())
        self.critic = mk_c("Critic_")
        self.targetCritic = mk_c("TargetCritic_")
        self.targetCritic.load_state_dict(self.critic.state_dict())
        self.bc_actor = mk_a("") if self._MODE == 0 else None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.hirl_agent_destroy(h)
            except Exception:
                pass
            self._h = None

    def __str__(self):
        return f"{type(self).__name__}(fused sm_100a update, batch={self.batchSize}, device={self.device})"

    # ---- parameter traffic ---------------------------------------------------------------
    def push_parameters(self):
        """host containers -> device arena (after construction / checkpoint load)."""
        for net_id, mod in self._nets.items():
            v = mod.flat()
            _lib.check(self._lib.hirl_params_import(self._h, 0, net_id, _lib.ptr(v), v.size))
        self._host_stale = False

    def pull_parameters(self):
        """device arena -> host containers (before a checkpoint save / state_dict read)."""
        for net_id, mod in self._nets.items():
            mod.load_flat(self.export_flat(net_id))
        self._host_stale = False

    def export_flat(self, net_id) -> np.ndarray:
        n = _lib.CRITIC_PARAMS if net_id in (_lib.NET_CRITIC, _lib.NET_TARGET_CRITIC) else _lib.ACTOR_PARAMS
        out = np.empty(n, np.float32)
        _lib.check(self._lib.hirl_params_export(self._h, 0, net_id, _lib.ptr(out), n))
        return out

    def export_grads(self, net_id) -> np.ndarray:
        n = _lib.CRITIC_PARAMS if net_id == _lib.NET_CRITIC else _lib.ACTOR_PARAMS
        out = np.empty(n, np.float32)
        _lib.check(self._lib.hirl_grads_export(self._h, 0, net_id, _lib.ptr(out), n))
        return out

    def export_adam(self, net_id):
        n = _lib.CRITIC_PARAMS if net_id == _lib.NET_CRITIC else _lib.ACTOR_PARAMS
        m, v = np.empty(n, np.float32), np.empty(n, np.float32)
        _lib.check(self._lib.hirl_adam_export(self._h, 0, net_id, _lib.ptr(m), _lib.ptr(v), n))
        return m, v

    def import_adam(self, net_id, m, v, step):
        m, v = _lib.f32c(m), _lib.f32c(v)
        _lib.check(self._lib.hirl_adam_import(self._h, 0, net_id, _lib.ptr(m), _lib.ptr(v), m.size))
        if net_id == _lib.NET_CRITIC:
            self._critic_t = int(step)
        else:
            self._actor_t = int(step)

    # ---- acting (HIRL.py:192-212) --------------------------------------------------------
    def _actor_forward(self, state, net_id=_lib.NET_ACTOR) -> np.ndarray:
        s = _lib.f32c(state).reshape(-1, 13)
        out = np.empty((s.shape[0], 4), np.float32)
        _lib.check(self._lib.hirl_act(self._h, 0, net_id, _lib.ptr(s), s.shape[0], _lib.ptr(out)))
        return out

    def _noisy(self, state, sigma):
        a = torch.from_numpy(self._actor_forward(state)[0])
        noise = torch.normal(mean=torch.zeros(self.actionDim), std=torch.ones(self.actionDim) * sigma)
        return (a + noise).clamp(-1, +1).numpy()

    def chooseAction(self, state):
        return self._noisy(state, self.actionNoise)

    def chooseActionSmallNoise(self, state):
        return self._noisy(state, self.actionNoise / 10)

    def chooseActionNoNoise(self, state):
        return self._actor_forward(state)[0]

    def act_batch(self, states, net="actor") -> np.ndarray:
        """Noise-free actions for n states in one call (batched acting)."""
        ids = {"actor": _lib.NET_ACTOR, "targetActor": _lib.NET_TARGET_ACTOR, "bc_actor": _lib.NET_BC_ACTOR}
        return self._actor_forward(states, ids[net])

    def act_batch_dev(self, states, noise=None, net="actor"):
        """Device-resident batched acting: states (n,13) f32 CUDA tensor -> actions (n,4) f32 CUDA
        tensor on the caller's current stream, no host round trip.  ``noise``: None, a (4,) tensor shared by all rows
        (the reference's per-call draw, HIRL.py:196) or an (n,4) tensor; actions = clamp(pi(s) + noise, -1, 1)."""
        ids = {"actor": _lib.NET_ACTOR, "targetActor": _lib.NET_TARGET_ACTOR, "bc_actor": _lib.NET_BC_ACTOR}
        if not states.is_cuda or states.dtype != torch.float32 or not states.is_contiguous() or states.shape[-1] != 13:
            raise ValueError("act_batch_dev takes a contiguous (n,13) float32 CUDA tensor")
        n = states.numel() // 13
        out = torch.empty((n, 4), dtype=torch.float32, device=states.device)
        nz, nrows = None, 0
        if noise is not None:
            noise = noise.to(device=states.device, dtype=torch.float32).contiguous()
            nrows = noise.numel() // 4
            nz = noise.data_ptr()
        stream = torch.cuda.current_stream(states.device)
        _lib.check(self._lib.hirl_act_dev(self._h, 0, ids[net], states.data_ptr(), n, out.data_ptr(), nz, nrows, stream.cuda_stream))
        # the act kernel READS the actor on the caller's stream while learn() WRITES it on the agent's own stream:
        # the next update must wait for this launch (acting interleaved with training without a store in between)
        self._act_event = torch.cuda.Event()
        self._act_event.record(stream)
        return out

    def store_batch_dev(self, states, actions, next_states, rewards, flags):
        """Device-resident store of n transitions (see UniformMemory.store_batch_dev)."""
        self.buffer.store_batch_dev(states, actions, next_states, rewards, flags)

    def _update_stream(self):
        if getattr(self, "_ext_stream", None) is None:
            p = C.c_void_p()
            _lib.check(self._lib.hirl_agent_stream(self._h, C.byref(p)))
            self._ext_stream = torch.cuda.ExternalStream(p.value, device=self.device)
        return self._ext_stream

    def store(self, *args):
        self.buffer.store(*args)

    def upsample_expert_data(self, BCActions):
        return np.where(BCActions[:, 3] == 1)[0]

    # ---- the update ------------------------------------------------------------------------
    def _draw(self, k, bc_weight_now, expert_num_now, bc_warm_up_weight, forced, defer_bc=False):
        """Fills staging slot k with one call's index/noise draws, in the reference's RNG order.  ``defer_bc``: the
        cloning batch of this call is being drawn (from np.random, in call order) by learn_many's helper thread and is
        written into the slot by the caller."""
        B = self.batchSize
        st = self._st_steps[k]
        forced = forced or {}
        warm = bool(self.expert_warm_up and expert_num_now)
        n_ex = int(expert_num_now) if warm else 0
        pre = None
        if k == 0 and not forced and not self.buffer.upsample:
            pre = self._pre_py.take(self._py_requests(len(self.buffer), n_ex))   # pre-drawn while the GPU ran the last call
        if pre is not None:
            idx_on, idx_ex = pre[0], (pre[1] if n_ex else [])
        elif warm:
            idx_on = forced.get("online") if "online" in forced else self.buffer.sample_indices_array(B - n_ex)
            idx_ex = forced.get("expert") if "expert" in forced else self.expert_buffer.sample_indices_array(n_ex)
        else:
            idx_on = forced.get("online") if "online" in forced else self.buffer.sample_indices_array(B)
            idx_ex = []
        n_on = B - n_ex
        if len(idx_on) != n_on or len(idx_ex) != n_ex:
            raise ValueError("index streams do not match the batch split")
        if n_on == 0:
            # the reference fails on random.sample(memory, 0) + zip(*[]) (HIRL.py:226, buffer.py:47)
            raise TypeError("expert_num_now == batchSize leaves no online rows (Transition(*zip(*[])) in the reference)")
        self._st_on[k, :n_on] = idx_on
        if n_ex:
            self._st_ex[k, :n_ex] = idx_ex
        if self._MODE == 0:
            n_bc = len(self.expert_states)
            if "bc" in forced:
                idx_bc = forced["bc"]
            elif defer_bc:
                idx_bc = self._st_bc[k]          # placeholder: the row is overwritten after the helper thread is joined
            else:
                pre = self._pre_np.take([(n_bc, B)]) if k == 0 and not self.buffer.upsample else None
                idx_bc = pre[0] if pre is not None else fast_rng.choice_noreplace(n_bc, B)   # == np.random.choice(n_bc, B, replace=False)
            self._st_bc[k, :] = idx_bc
            if self.expert_upsample and "bc" not in forced:
                # HIRL.py:253-257: one random row of the cloning batch is replaced by a firing expert row
                selected = np.random.choice(self.target_indices)
                self._st_bc[k, np.random.randint(B)] = selected
        else:
            idx_bc = None
        if "noise" in forced:
            noise = np.asarray(forced["noise"], np.float32)
        else:
            # HIRL.py:265 torch.normal(mean=zeros, std=ones * 0.2): at::normal(Tensor, Tensor) is out.normal_(0, 1).mul_(std).add_(mean);
            # the same three steps on a preallocated buffer consume the global generator identically (bit-identical values and
            # generator state, tests/test_host_logic.py) at 2.4 us instead of 14 us per call
            self._noise_buf.normal_()
            noise = self._noise_np * self._noise_std_np + self._noise_mean_np
        # schedule + Adam scalars
        do_actor = bool(self.actorTrainable)
        self._critic_t += 1
        c_neg, c_bs = _adam_scalars(self.criticLR, self._critic_t)
        a_neg = a_bs = 0.0
        do_polyak = False
        if do_actor:
            self._actor_t += 1
            a_neg, a_bs = _adam_scalars(self.actorLR, self._actor_t)
            self.update_count += 1
            do_polyak = self.update_count % self.target_update_freq == 0
        self.actorTrainable = not self.actorTrainable
        # one record assignment (field order = hirl_step_t)
        wi, wf = self._st_i32[k], self._st_f32[k]
        wi[0:4] = (n_on, n_ex, do_actor, do_polyak)
        wf[4:7] = (bc_weight_now, bc_warm_up_weight, forced.get("soft_weight", -1.0))
        wf[7:11] = noise
        wf[11:15] = (c_neg, c_bs, a_neg, a_bs)
        wi[15] = self._MODE
        self._ld, self._ld_cache = (k, n_on, n_ex, idx_bc is not None, noise), None
        return do_actor

    @property
    def last_draws(self):
        """The index / noise streams of the last drawn call (copies, built on first access)."""
        if self._ld_cache is None and self._ld is not None:
            k, n_on, n_ex, has_bc, noise = self._ld
            self._ld_cache = {"online": self._st_on[k, :n_on].copy(), "expert": self._st_ex[k, :n_ex].copy(),
                              "bc": self._st_bc[k].copy() if has_bc else None, "noise": np.array(noise, np.float32)}
        return self._ld_cache

    @last_draws.setter
    def last_draws(self, value):
        self._ld, self._ld_cache = None, value

    def _finish(self, k, do_actor, bc_weight_now, bc_warm_up_weight):
        """Builds the reference's return tuple (HIRL.py:334) from the kernel's scalars."""
        o = self._st_out[k]
        critic_loss = np.float32(o[0])
        if do_actor:
            if self._MODE == 0:
                w = bc_weight_now
                if bc_weight_now == SOFT_SENTINEL:
                    forced = float(self._st_steps[k]["force_soft_weight"])
                    mean = np.float32(forced) if forced >= 0 else np.float32(o[6]) / np.float32(self.batchSize)
                    w = float(mean) + bc_warm_up_weight
                if w > 1:
                    w = 1
                self.bc_weight = w
                fire = float(o[4]) * self.loss_lambda
            else:
                fire = 0.0
            self._ret = (critic_loss, np.float32(o[1]), np.float32(o[2]), np.float32(o[3]), fire, self.bc_weight)
        else:
            self._ret = (critic_loss,) + self._ret[1:]
        return self._ret

    def _py_requests(self, n_online_rows, n_ex):
        """The random.sample calls one learn() makes, as (population, k) pairs (HIRL.py:224-233)."""
        B = self.batchSize
        return [(n_online_rows, B - n_ex), (len(self.expert_buffer), n_ex)] if n_ex else [(n_online_rows, B)]

    def _predraw_next(self, expert_num_now):
        """Runs while the GPU executes the current call: draws the NEXT call's index streams on copies of the
        generator states (fast_rng.Predraw).  Prediction: same expert_num_now, online ring grown by as many rows as
        between the last two calls.  A wrong guess or any foreign use of the generators just discards the copy."""
        if not self.PREDRAW:
            return
        n_now = len(self.buffer)
        n_next = min(self.buffer.capacity, n_now + max(0, n_now - self._n_online_prev))
        self._n_online_prev = n_now
        n_ex = int(expert_num_now) if (self.expert_warm_up and expert_num_now) else 0
        if not self.buffer.upsample:
            if self._MODE != 1:
                # a full 20 000-element shuffle (~60 us): handed to the library's worker thread, collected by the next call's take()
                self._pre_np.prepare([(len(self.expert_states), self.batchSize)])
            self._pre_py.prepare(self._py_requests(n_next, n_ex))

    def _pre_launch(self):
        self.buffer.flush()
        self.expert_buffer.flush()
        if getattr(self.buffer, "_dev_event", None) is not None:   # device-side appends run on the caller's stream
            self.buffer.wait_device_appends(self._update_stream())
        if getattr(self, "_act_event", None) is not None:           # so do device-side acting launches (they read the actor)
            self._update_stream().wait_event(self._act_event)
            self._act_event = None

    def _launch(self, k_steps, predraw_for=None):
        self._pre_launch()
        self._submit(0, k_steps)
        try:
            if predraw_for is not None:
                self._predraw_next(predraw_for)
        finally:
            _lib.check(self._lib.hirl_learn_wait(self._h))
        self._host_stale = True

    def learn(self, bc_weight_now, expert_num_now, bc_warm_up_weight=0, forced=None):
        do_actor = self._draw(0, bc_weight_now, expert_num_now, bc_warm_up_weight, forced)
        self._launch(1, predraw_for=None if forced else expert_num_now)
        return self._finish(0, do_actor, bc_weight_now, bc_warm_up_weight)

    def learn_many(self, calls, forced=None):
        """Runs len(calls) consecutive learn() calls in ONE kernel launch.

        ``calls`` is a sequence of (bc_weight_now, expert_num_now, bc_warm_up_weight).  The index and
        noise streams are pre-drawn in call order, which is legal because they depend only on the
        buffer lengths, and those do not change between fused calls.  Returns the list of
        per-call return tuples.
        """
        calls = list(calls)
        out = []
        for base in range(0, len(calls), self._max_steps):
            chunk = calls[base:base + self._max_steps]
            flags = []
            self._pre_launch()
            try:
                # pipelined: while the GPU runs sub-chunk j the host draws sub-chunk j+1 (launches queue on one stream,
                # so the device still executes the calls back to back, in order)
                # The first sub-chunks are short (1, 2, 4 calls): nothing overlaps the very first draw, so it should be brief.
                bounds, lo, step = [], 0, 1
                if self.FUSE_SUB >= len(chunk):        # the caller asked for ONE launch (profilers set FUSE_SUB = max_fused_steps)
                    bounds, lo = [(0, len(chunk))], len(chunk)
                while lo < len(chunk):
                    hi = min(len(chunk), lo + min(step, self.FUSE_SUB))
                    bounds.append((lo, hi))
                    lo, step = hi, step * 2
                for lo, hi in bounds:
                    # The np.random stream (one cloning batch per call: the longest draw) is advanced by a helper
                    # thread while this thread draws the random / torch streams of the same calls.  The three
                    # generators are independent and each is still consumed in call order; the C draws release the
                    # GIL.  Not with forced streams or the up-sampling switches (those use np.random elsewhere too).
                    fut = pend = None
                    if self._MODE == 0 and not forced and not self.buffer.upsample and not self.expert_upsample:
                        n_bc, B = len(self.expert_states), self.batchSize
                        if fast_rng.np_fast() and 0 < B <= n_bc < (1 << 31):
                            # all cloning batches of the sub-chunk in ONE library call on the library's worker thread, straight
                            # on numpy's live MT19937 state (nothing else touches np.random until it is joined below)
                            self._pre_np._join()
                            buf = np.empty((hi - lo, B), np.int32)
                            live = fast_rng._np_live()
                            if self._lib.hirl_host_choice_many_async(C.byref(live), n_bc, B, hi - lo, _lib.ptr(buf)) == 0:
                                pend = buf
                        if pend is None:
                            fut = _draw_pool().submit(lambda m=hi - lo: [fast_rng.choice_noreplace(n_bc, B) for _ in range(m)])
                    try:
                        for k in range(lo, hi):
                            c = chunk[k]
                            flags.append(self._draw(k, c[0], c[1], c[2] if len(c) > 2 else 0, forced[base + k] if forced else None,
                                                    defer_bc=fut is not None or pend is not None))
                    finally:                                                  # always joined: np.random must be quiescent afterwards
                        ready = fut.result() if fut is not None else None
                        if pend is not None:
                            self._lib.hirl_host_async_wait()
                            ready = pend
                    if ready is not None:
                        for k in range(lo, hi):
                            self._st_bc[k, :] = ready[k - lo]
                        self.last_draws["bc"] = self._st_bc[hi - 1].copy()
                    self._submit(lo, hi - lo)
            finally:
                _lib.check(self._lib.hirl_learn_wait(self._h))
            self._host_stale = True
            out += [self._finish(k, flags[k], c[0], c[2] if len(c) > 2 else 0) for k, c in enumerate(chunk)]
        return out

    def _slot_ptrs(self, k):
        """Staging pointers of step slot k (steps, idx_online, idx_expert, idx_bc, out_scalars)."""
        B = self.batchSize
        p = self._st_addr
        return (C.c_void_p(p[0] + k * _lib.STEP_DTYPE.itemsize), C.c_void_p(p[1] + k * B * 4), C.c_void_p(p[2] + k * B * 4),
                C.c_void_p(p[3] + k * B * 4), C.c_void_p(p[4] + k * _lib.OUT_SCALARS * 4))

    # aliases named by the north star
    select_action = chooseAction
    train = learn

    # ---- test hooks --------------------------------------------------------------------------
    def gather(self, idx_online, idx_expert=()):
        self.buffer.flush()
        self.expert_buffer.flush()
        io, ie = _lib.i32c(idx_online), _lib.i32c(idx_expert)
        out = np.empty((io.size + ie.size, 32), np.float32)
        _lib.check(self._lib.hirl_gather(self._h, 0, _lib.ptr(io), io.size, _lib.ptr(ie), ie.size, _lib.ptr(out)))
        return out

    def gather_bc(self, idx_bc):
        ib = _lib.i32c(idx_bc)
        s, a = np.empty((ib.size, 13), np.float32), np.empty((ib.size, 4), np.float32)
        _lib.check(self._lib.hirl_gather_bc(self._h, _lib.ptr(ib), ib.size, _lib.ptr(s), _lib.ptr(a)))
        return s, a

    def debug_stop_after(self, stage):
        _lib.check(self._lib.hirl_debug_stop_after(self._h, int(stage)))

    def debug_buffer(self, buffer_id, index=0):
        n = C.c_size_t()
        _lib.check(self._lib.hirl_debug_export(self._h, buffer_id, index, None, 0, C.byref(n)))
        out = np.empty(n.value, np.float32)
        _lib.check(self._lib.hirl_debug_export(self._h, buffer_id, index, _lib.ptr(out), out.size, None))
        return out

    # ---- checkpoints (HIRL.py:99-103, 142-146, 336-350) -----------------------------------
    @staticmethod
    def _ckpt_path(ajan, model_name, net_name):
        return model_name + "\\{}".format(ajan) + net_name   # the reference's literal backslash

    def saveCheckpoints(self, ajan, model_name):
        self.pull_parameters()
        for mod in (self.critic, self.actor, self.targetCritic, self.targetActor):
            torch.save(mod.state_dict(), self._ckpt_path(ajan, model_name, mod.name))

    def loadCheckpoints(self, ajan, model_name):
        if self._host_stale:
            self.pull_parameters()
        for mod in (self.critic, self.actor, self.targetCritic, self.targetActor):
            mod.load_state_dict(torch.load(self._ckpt_path(ajan, model_name, mod.name), map_location="cpu"))
        self.push_parameters()

    def load_bc_actor(self, ajan, model_name):
        if self._host_stale:
            self.pull_parameters()
        self.bc_actor.load_state_dict(torch.load(self._ckpt_path(ajan, model_name, self.bc_actor.name), map_location="cpu"))
        self.push_parameters()

    def save_training_state(self, path):
        """Checkpoint extras the reference does not keep: Adam moments, counters."""
        am, av = self.export_adam(_lib.NET_ACTOR)
        cm, cv = self.export_adam(_lib.NET_CRITIC)
        np.savez(path, actor_m=am, actor_v=av, critic_m=cm, critic_v=cv, actor_t=self._actor_t, critic_t=self._critic_t,
                 update_count=self.update_count, actorTrainable=self.actorTrainable)

    def load_training_state(self, path):
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        self.import_adam(_lib.NET_ACTOR, z["actor_m"], z["actor_v"], int(z["actor_t"]))
        self.import_adam(_lib.NET_CRITIC, z["critic_m"], z["critic_v"], int(z["critic_t"]))
        self.update_count, self.actorTrainable = int(z["update_count"]), bool(z["actorTrainable"])

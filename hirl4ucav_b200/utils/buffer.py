"""Replay ring with a device-resident copy (mirrors reference hirl/utils/buffer.py:8-54).

Rows live in a pinned host array of 32 fp32 words (s[13] a[4] s'[13] r done) that is
flushed to the GPU table lazily, right before the next fused update needs it.  Sampling is
index-only: ``random.sample`` draws positions from the same Mersenne stream the reference
consumes (buffer.py:45) and the gather happens inside the CUDA kernel.
"""
from __future__ import annotations

import random
from collections import namedtuple

import numpy as np
import torch

from .. import _lib
from . import fast_rng

Transition = namedtuple("Transition", ("state", "action", "next_state", "reward", "done", "step_success"))
S, A, ROW = 13, 4, 32


def _pinned(shape, dtype):
    t = torch.zeros(shape, dtype=dtype)
    if torch.cuda.is_available():
        t = t.pin_memory()
    return t


class _MemoryView:
    """len()/indexing view that stands in for the reference's ``memory`` list."""

    def __init__(self, ring):
        self._r = ring

    def __len__(self):
        return self._r._size

    def __getitem__(self, i):
        r = self._r
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += r._size
        if not 0 <= i < r._size:
            raise IndexError("list index out of range")
        if r._n_stale:
            r.sync_host()
        row = r._rows[i]
        return Transition(row[0:S].copy(), row[S:S + A].copy(), row[S + A:S + A + S].copy(), float(row[30]),
                          bool(row[31] != 0), int(r._success[i]))

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class UniformMemory(object):
    def __init__(self, capacity, upsample=False):
        self.capacity = int(capacity)
        self.position = 0
        self.upsample = upsample
        # buffer.py:18: per-slot sampling weight, only with up-sampling (off in the reference's constructor, HIRL.py:184)
        self.priorities_memory = np.zeros(self.capacity) if upsample else None
        self._size = 0
        self._rows_t = _pinned((self.capacity, ROW), torch.float32)
        self._rows = self._rows_t.numpy()
        self._success = np.zeros(self.capacity, np.int8)
        self._dirty_from = None   # first dirty slot (ring order) and count
        self._dirty_n = 0
        self._backend = None      # (handle, buffer id) once attached to an agent
        # Device-side appends (store_batch_dev) are tracked PER SLOT, so the bookkeeping is bounded by the ring whatever
        # the number of appends: which slots hold rows / step_success values the host mirror has not seen yet, and one
        # device array of the step_success values (no reference to the caller's flag tensors is kept).
        self._rows_stale = None   # np.bool_[capacity], allocated on the first device append
        self._succ_stale = None
        self._dev_success = None  # torch.int8[capacity] on the agent's device
        self._n_stale = 0         # slots with a stale host row or step_success
        self.memory = _MemoryView(self)

    # ---- device attachment -------------------------------------------------------------
    def _attach(self, handle, which):
        self._backend = (handle, which)
        if self._size:
            self._dirty_from, self._dirty_n = 0, self._size

    def flush(self):
        """Copies rows stored since the last flush to the GPU table."""
        if self._backend is None or not self._dirty_n:
            return 0
        handle, which = self._backend
        n = min(self._dirty_n, self.capacity)
        first = self._dirty_from if self._dirty_n <= self.capacity else self.position
        lib = _lib.load()
        if first + n <= self.capacity:
            _lib.check(lib.hirl_store_rows(handle, 0, which, first, n, _lib.ptr(self._rows[first:first + n])))
        else:
            k = self.capacity - first
            _lib.check(lib.hirl_store_rows(handle, 0, which, first, k, _lib.ptr(self._rows[first:])))
            _lib.check(lib.hirl_store_rows(handle, 0, which, 0, n - k, _lib.ptr(self._rows[:n - k])))
        self._dirty_from, self._dirty_n = None, 0
        return n

    # ---- device-resident append -------------------------------------------------
    def store_batch_dev(self, states, actions, next_states, rewards, flags):
        """buffer.py:20-36 for n transitions that already live on the GPU (CUDA tensors: (n,13) f32, (n,4) f32,
        (n,13) f32, (n,) f32 and the (n,) int32 flag words hirl_env_step_dev writes: bit 0 done, byte 1
        step_success).  One kernel on the caller's current stream packs the 128-byte rows straight into the ring:
        no H2D/D2H.  The host mirror (``memory[i]``) is refreshed lazily by sync_host()."""
        if self._backend is None:
            raise _lib.HirlError("store_batch_dev needs a ring attached to an agent")
        handle, which = self._backend
        ts = [states, actions, next_states, rewards, flags]
        want = [torch.float32] * 4 + [torch.int32]
        if any((not t.is_cuda) or t.dtype != d or not t.is_contiguous() for t, d in zip(ts, want)):
            raise ValueError("store_batch_dev takes contiguous CUDA tensors: f32 states/actions/next_states/rewards, int32 flags")
        n = int(states.shape[0])
        if n > self.capacity or tuple(states.shape) != (n, S) or tuple(actions.shape) != (n, A) or \
                tuple(next_states.shape) != (n, S) or rewards.numel() != n or flags.numel() != n:
            raise ValueError("store_batch_dev: inconsistent shapes")
        self.flush()   # rows stored on the host before this call are older: they must land first
        lib = _lib.load()
        stream = torch.cuda.current_stream(states.device)
        _lib.check(lib.hirl_store_transitions_dev(handle, 0, which, self.position, n, states.data_ptr(), actions.data_ptr(),
                                                  next_states.data_ptr(), rewards.data_ptr(), flags.data_ptr(),
                                                  stream.cuda_stream))
        if self._rows_stale is None:
            self._rows_stale = np.zeros(self.capacity, np.bool_)
            self._succ_stale = np.zeros(self.capacity, np.bool_)
            self._dev_success = torch.zeros(self.capacity, dtype=torch.int8, device=states.device)
        idx = (self.position + np.arange(n)) % self.capacity
        succ = ((flags >> 8) & 0xFF).to(torch.int8)          # byte 1 of the flag word = step_success (two's complement)
        if self.position + n <= self.capacity:
            self._dev_success[self.position:self.position + n] = succ
        else:
            k = self.capacity - self.position
            self._dev_success[self.position:] = succ[:k]
            self._dev_success[:n - k] = succ[k:]
        self._rows_stale[idx] = True     # a slot overwritten by ring wrap is simply marked again: nothing accumulates
        self._succ_stale[idx] = True
        self._n_stale = int(np.count_nonzero(self._rows_stale | self._succ_stale))
        self._dev_event = torch.cuda.Event()
        self._dev_event.record(stream)
        if self.upsample:   # the weighted draw needs step_success on the host (buffer.py:21-29)
            self.sync_host(rows=False)
        self._size = min(self.capacity, self._size + n)
        self.position = (self.position + n) % self.capacity

    def wait_device_appends(self, handle_stream):
        """Orders the agent's update stream after the pending device-side appends."""
        ev = getattr(self, "_dev_event", None)
        if ev is not None:
            handle_stream.wait_event(ev)
            self._dev_event = None

    def sync_host(self, rows=True):
        """Brings step_success (and with rows=True the row contents) of device-appended slots to the host mirror."""
        if not self._n_stale:
            return
        handle, which = self._backend
        lib = _lib.load()
        sidx = np.flatnonzero(self._succ_stale)
        if sidx.size:
            torch.cuda.synchronize()                          # appends may have been queued on any stream
            succ = self._dev_success.cpu().numpy()
            self._success[sidx] = succ[sidx]
            if self.upsample:
                self.priorities_memory[sidx] = np.where(succ[sidx] == 1, 1.0, 0.1)
            self._succ_stale[:] = False
        if rows:
            ridx = np.flatnonzero(self._rows_stale)
            if ridx.size:
                self.flush()                                  # host-dirty rows first (they never share a slot with a stale one)
                torch.cuda.synchronize()
                idx32 = _lib.i32c(ridx)
                n = int(ridx.size)
                out = np.empty((n, ROW), np.float32)
                on, ex = (idx32, None) if which == _lib.BUF_ONLINE else (None, idx32)
                _lib.check(lib.hirl_gather(handle, 0, None if on is None else _lib.ptr(on), 0 if on is None else n,
                                           None if ex is None else _lib.ptr(ex), 0 if ex is None else n, _lib.ptr(out)))
                self._rows[ridx] = out
                self._rows_stale[:] = False
        self._n_stale = int(np.count_nonzero(self._rows_stale)) if self._rows_stale is not None else 0

    def _host_wrote(self, p):
        """A host-side store into slot p supersedes whatever the device appended there."""
        if self._n_stale and (self._rows_stale[p] or self._succ_stale[p]):
            self._rows_stale[p] = self._succ_stale[p] = False
            self._n_stale -= 1

    # ---- reference surface ---------------------------------------------------------------
    def store(self, state, action, next_state, reward, done, step_success):
        self._host_wrote(self.position)
        row = self._rows[self.position]
        row[0:S] = state            # float64 -> float32, the cast the reference does at sample time
        row[S:S + A] = action
        row[S + A:S + A + S] = next_state
        row[30] = reward
        row[31] = 1.0 if done else 0.0
        self._success[self.position] = step_success
        if self.upsample:   # buffer.py:21-29
            self.priorities_memory[self.position] = 1.0 if step_success == 1 else 0.1
        if self._dirty_n == 0:
            self._dirty_from = self.position
        self._dirty_n += 1
        if self._size < self.capacity:
            self._size += 1
        self.position = (self.position + 1) % self.capacity

    def store_rows(self, rows32: np.ndarray, success=None):
        """Bulk append of pre-packed rows (expert-buffer construction, batched acting)."""
        rows32 = np.ascontiguousarray(rows32, np.float32).reshape(-1, ROW)
        for i in range(rows32.shape[0]):
            p = self.position
            self._host_wrote(p)
            self._rows[p] = rows32[i]
            self._success[p] = 0 if success is None else success[i]
            if self.upsample:
                self.priorities_memory[p] = 1.0 if self._success[p] == 1 else 0.1
            if self._dirty_n == 0:
                self._dirty_from = p
            self._dirty_n += 1
            if self._size < self.capacity:
                self._size += 1
            self.position = (p + 1) % self.capacity

    def sample_indices(self, batchSize):
        """The positions ``random.sample(self.memory, batchSize)`` would pick (buffer.py:45); with up-sampling the
        weighted numpy draw WITH replacement of buffer.py:39-43.  Same RNG stream consumption as the reference."""
        if self.upsample:
            pr = self.priorities_memory[:self._size]
            return [int(i) for i in np.random.choice(self._size, batchSize, p=pr / np.sum(pr))]
        return fast_rng.sample_indices(self._size, batchSize)   # == random.sample(range(self._size), batchSize)

    def sample_indices_array(self, batchSize):
        """sample_indices as an int32 array (the update path: no Python list between the draw and the staging row)."""
        if self.upsample:
            return np.asarray(self.sample_indices(batchSize), np.int32)
        return fast_rng.sample_indices_array(self._size, batchSize)

    def sample(self, batchSize):
        idx = self.sample_indices(batchSize)
        batch = Transition(*zip(*[self.memory[i] for i in idx]))
        return batch.state, batch.action, batch.next_state, batch.reward, batch.done

    def fullEnough(self, batchSize):
        return self._size >= batchSize

    def __len__(self):
        return self._size

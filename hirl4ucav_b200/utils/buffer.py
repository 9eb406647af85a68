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
        row = r._rows[i]
        return Transition(row[0:S].copy(), row[S:S + A].copy(), row[S + A:S + A + S].copy(), float(row[30]),
                          bool(row[31] != 0), int(r._success[i]))

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class UniformMemory(object):
    def __init__(self, capacity, upsample=False):
        if upsample:
            # disabled in the reference (HIRL.py:184: buffer_upsample = False)
            raise NotImplementedError("prioritised up-sampling is outside the fused path")
        self.capacity = int(capacity)
        self.position = 0
        self.upsample = upsample
        self.priorities_memory = None
        self._size = 0
        self._rows_t = _pinned((self.capacity, ROW), torch.float32)
        self._rows = self._rows_t.numpy()
        self._success = np.zeros(self.capacity, np.int8)
        self._dirty_from = None   # first dirty slot (ring order) and count
        self._dirty_n = 0
        self._backend = None      # (handle, buffer id) once attached to an agent
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

    # ---- reference surface ---------------------------------------------------------------
    def store(self, state, action, next_state, reward, done, step_success):
        row = self._rows[self.position]
        row[0:S] = state            # float64 -> float32, the cast the reference does at sample time
        row[S:S + A] = action
        row[S + A:S + A + S] = next_state
        row[30] = reward
        row[31] = 1.0 if done else 0.0
        self._success[self.position] = step_success
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
            self._rows[p] = rows32[i]
            self._success[p] = 0 if success is None else success[i]
            if self._dirty_n == 0:
                self._dirty_from = p
            self._dirty_n += 1
            if self._size < self.capacity:
                self._size += 1
            self.position = (p + 1) % self.capacity

    def sample_indices(self, batchSize):
        """The positions ``random.sample(self.memory, batchSize)`` would pick (buffer.py:45)."""
        return random.sample(range(self._size), batchSize)

    def sample(self, batchSize):
        idx = self.sample_indices(batchSize)
        batch = Transition(*zip(*[self.memory[i] for i in idx]))
        return batch.state, batch.action, batch.next_state, batch.reward, batch.done

    def fullEnough(self, batchSize):
        return self._size >= batchSize

    def __len__(self):
        return self._size

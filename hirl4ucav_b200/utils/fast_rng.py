"""Exact, faster forms of the two host-side index draws of ``learn()``, plus speculative pre-drawing.

The reference owns three RNG streams (hirl/utils/seed.py:5-10); a drop-in must consume them identically or
seeded runs diverge.  Everything here returns the SAME indices as the interpreter-level originals and leaves the
global generators in the SAME state:

* ``sample_indices(n, k)``   == ``random.sample(range(n), k)``                 (buffer.py:45)
* ``choice_noreplace(n, k)`` == ``np.random.choice(n, k, replace=False)``      (HIRL.py:249, BC.py:167)

Both are restatements of the generators' own algorithms in C (csrc/host_rng.h, hirl_host_* in the C ABI)
running directly on the generators' MT19937 state, reached through ``bit_generator.ctypes.state_address``
(numpy) and the ``_random.Random`` object layout (CPython).  Each access path is self-checked once per process
against the real generator; a mismatch (different numpy / CPython build) switches that fast path off and the
plain ``random.sample`` / ``np.random.choice`` are called instead -- slower, never different.

Speculation (class ``Predraw``): while the GPU executes call t, the host draws call t+1's indices on a COPY of the
generator state.  At call t+1 the live state is compared with the snapshot the copy was taken from: equal means
nobody consumed or reseeded the generator in between, so committing the advanced copy and using the pre-drawn
indices is indistinguishable from drawing now; different means the speculation is discarded and the draw is made
from the live state.  Exact in both cases.
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np

from .. import _lib


class MTState(C.Structure):          # include/hirl_b200.h: hirl_mt19937_t
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int32)]


class _PyRandomObject(C.Structure):  # CPython Modules/_randommodule.c: RandomObject (after PyObject_HEAD)
    _fields_ = [("ob_refcnt", C.c_ssize_t), ("ob_type", C.c_void_p), ("index", C.c_int), ("state", C.c_uint32 * 624)]


_ok = {"np": None, "py": None}


# ---- live-state access -------------------------------------------------------------------------------------------
_live_cache = {}      # state address -> ctypes view (the generators' state blocks never move while the objects live)


def _np_live(rs=None):
    rs = np.random.mtrand._rand if rs is None else rs
    addr = rs._bit_generator.ctypes.state_address
    v = _live_cache.get(addr)
    if v is None:
        v = _live_cache[addr] = C.cast(addr, C.POINTER(MTState)).contents
    return v


def _py_live(inst=None):
    inst = random._inst if inst is None else inst
    if inst is random._inst:
        v = _live_cache.get("py")
        if v is None or v[0] is not inst:
            v = _live_cache["py"] = (inst, _PyRandomObject.from_address(id(inst)))
        return v[1]
    return _PyRandomObject.from_address(id(inst))


def _py_read(inst=None) -> MTState:
    live, st = _py_live(inst), MTState()
    C.memmove(st.key, live.state, 624 * 4)
    st.pos = live.index
    return st


def _py_write(st: MTState, inst=None):
    live = _py_live(inst)
    C.memmove(live.state, st.key, 624 * 4)
    live.index = st.pos


def _copy(st: MTState) -> MTState:
    out = MTState()
    C.memmove(C.byref(out), C.byref(st), C.sizeof(MTState))
    return out


_memcmp = C.CDLL(None).memcmp
_memcmp.restype, _memcmp.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]


def _same(a: MTState, b: MTState) -> bool:
    return a.pos == b.pos and _memcmp(C.byref(a), C.byref(b), 624 * 4) == 0


def _live_equals(kind: str, snap: MTState) -> bool:
    """Compares the live generator state with a snapshot in place (no copy of the 2.5 KB state)."""
    if kind == "py":
        live = _py_live()
        return live.index == snap.pos and _memcmp(C.byref(live, _PyRandomObject.state.offset), C.byref(snap), 624 * 4) == 0
    live = _np_live()
    return live.pos == snap.pos and _memcmp(C.byref(live), C.byref(snap), 624 * 4) == 0


# ---- the C draws ---------------------------------------------------------------------------------------------------
def _c_sample(st: MTState, n: int, k: int):
    out = np.empty(k, np.int32)
    _lib.check(_lib.load().hirl_host_sample_indices(C.byref(st), n, k, _lib.ptr(out)))
    return out


def _c_choice(st: MTState, n: int, k: int):
    out = np.empty(k, np.int32)
    _lib.check(_lib.load().hirl_host_choice_noreplace(C.byref(st), n, k, _lib.ptr(out)))
    return out


# ---- self checks against the real generators -----------------------------------------------------------------------
def _check_np() -> bool:
    try:
        if np.random.get_state()[0] != "MT19937":
            return False
        for n, k in ((20000, 256), (2000, 64), (257, 256), (5, 5), (1, 1)):
            a, b = np.random.RandomState(777 + n), np.random.RandomState(777 + n)
            x = a.choice(n, k, replace=False)
            y = _c_choice(_np_live(b), n, k)
            sa, sb = a.get_state(), b.get_state()
            if not np.array_equal(x, y) or sa[2] != sb[2] or not np.array_equal(sa[1], sb[1]):
                return False
            if a.random_sample() != b.random_sample():
                return False
        return True
    except Exception:
        return False


def _check_py() -> bool:
    try:
        for n, k in ((30000, 192), (1047, 192), (1045, 192), (77, 3), (100000, 4096), (40, 40), (7, 0)):
            a, b = random.Random(12345 + n), random.Random(12345 + n)
            live = _py_live(b)
            ver, internal, gauss = b.getstate()
            if live.index != internal[624] or list(live.state[:8]) != list(internal[:8]):
                return False                        # object layout is not what we think it is
            x = a.sample(range(n), k)
            st = _py_read(b)
            y = _c_sample(st, n, k)
            _py_write(st, b)
            if x != y.tolist() or a.getstate() != b.getstate() or a.random() != b.random():
                return False
        return True
    except Exception:
        return False


def np_fast() -> bool:
    if _ok["np"] is None:
        _ok["np"] = _check_np()
    return _ok["np"]


def py_fast() -> bool:
    if _ok["py"] is None:
        _ok["py"] = _check_py()
    return _ok["py"]


# ---- public draws ------------------------------------------------------------------------------------------------
def sample_indices_array(n: int, k: int):
    """== random.sample(range(n), k) on the global generator, as an int32 array (no Python list in between: the update's staging
    rows are int32 arrays; at batch 4096 the list round trip cost more than the draw)."""
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")     # random.sample's own error
    if n >= 1 << 31 or not py_fast():
        return np.asarray(random.sample(range(n), k), np.int32)
    st = _py_read()
    out = _c_sample(st, n, k)
    _py_write(st)
    return out


def sample_indices(n: int, k: int):
    """== random.sample(range(n), k) on the global generator."""
    return sample_indices_array(n, k).tolist()


def choice_noreplace(n: int, k: int) -> np.ndarray:
    """== np.random.choice(n, k, replace=False) on the global RandomState."""
    if k > n:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")   # numpy's message
    if n < 1 or n >= 1 << 31 or not np_fast():
        return np.random.choice(n, k, replace=False)
    return _c_choice(_np_live(), n, k).astype(np.int64)


def sample_indices_on(inst, n: int, k: int):
    """== inst.sample(range(n), k) for a private ``random.Random`` instance (ensemble replicas)."""
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")
    if n >= 1 << 31 or not py_fast():
        return inst.sample(range(n), k)
    st = _py_read(inst)
    out = _c_sample(st, n, k)
    _py_write(st, inst)
    return out


def choice_noreplace_on(rs, n: int, k: int):
    """== rs.choice(n, k, replace=False) for a private ``np.random.RandomState`` (ensemble replicas)."""
    if k > n or n < 1 or n >= 1 << 31 or not np_fast():
        return rs.choice(n, k, replace=False)
    return _c_choice(_np_live(rs), n, k)


_exit_join_registered = False
# The buffers of the job the library's worker thread is running (generator-state copy, output array).  Held HERE, not only by the
# Predraw that started it: an agent may be garbage-collected while its pre-draw is still in flight, and the worker must never
# write into freed memory.  One job at a time library-wide, so one slot is enough; cleared when the job is joined.
_INFLIGHT = None


def _register_exit_join():
    """The worker thread writes into Python-owned buffers: never let the interpreter tear them down under it."""
    global _exit_join_registered
    if not _exit_join_registered:
        import atexit
        atexit.register(lambda: _lib.load().hirl_host_async_wait())
        _exit_join_registered = True


class Predraw:
    """Speculative draws for the next call (see the module docstring).  ``kind`` is "py" (random.sample stream:
    a list of (n, k) draws made back to back) or "np" (np.random.choice stream: one (n, k) draw)."""

    def __init__(self, kind: str):
        self.kind = kind
        self._snap = None
        self._pending = False
        self.hits = self.misses = 0

    def _live(self) -> MTState:
        return _py_read() if self.kind == "py" else _copy(_np_live())

    def enabled(self) -> bool:
        return py_fast() if self.kind == "py" else np_fast()

    def _join(self):
        """A draw handed to the library's worker thread is finished (its state copy and output array are ours again)."""
        global _INFLIGHT
        if self._pending:
            _lib.load().hirl_host_async_wait()
            self._pending = False
            _INFLIGHT = None

    def __del__(self):
        try:
            self._join()
        except Exception:
            pass

    def prepare(self, requests):
        """requests: list of (n, k) expected at the next call.  The np stream (one full shuffle, the longest thing the host
        does per call) runs on the library's worker thread; take() collects it."""
        global _INFLIGHT
        self._join()
        self._snap = None
        if not self.enabled() or any(not (0 <= k <= n and 0 < n < (1 << 31)) for n, k in requests):
            return
        snap = self._live()
        work = _copy(snap)
        if self.kind == "np" and len(requests) == 1:
            n, k = requests[0]
            out = np.empty(k, np.int32)
            if _lib.load().hirl_host_choice_noreplace_async(C.byref(work), n, k, _lib.ptr(out)) == 0:
                _register_exit_join()
                _INFLIGHT = (work, out)
                self._pending = True
                self._snap, self._after, self._req, self._outs = snap, work, list(requests), [out]
                return
        fn = _c_sample if self.kind == "py" else _c_choice
        outs = [fn(work, n, k) for n, k in requests]
        self._snap, self._after, self._req, self._outs = snap, work, list(requests), outs

    def ready(self) -> bool:
        return self._snap is not None

    def take(self, requests):
        """Returns the pre-drawn index arrays if they are exactly what drawing now would give, else None."""
        self._join()
        snap, self._snap = self._snap, None
        if snap is None or list(requests) != self._req or not _live_equals(self.kind, snap):
            self.misses += 1
            return None
        if self.kind == "py":
            _py_write(self._after)
        else:
            live = _np_live()
            C.memmove(C.byref(live), C.byref(self._after), C.sizeof(MTState))
        self.hits += 1
        return self._outs

"""CPU: the C-ABI library builds/loads and exports exactly what include/hirl_b200.h declares
(no compute calls here -- there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

from hirl4ucav_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "hirl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hirl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_lib.lib_path())
    for name in header_symbols():
        assert hasattr(raw, name), f"{name} declared in include/hirl_b200.h but not exported"
    assert lib.hirl_abi_version() == 1


def test_no_torch_or_cxx_types_cross_the_boundary():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [s for s in exported if s.startswith("hirl_")]
    assert sorted(ours) == header_symbols()


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Step) == 64 and _lib.STEP_DTYPE.itemsize == 64
    assert ctypes.sizeof(_lib.Config) == 48


def test_errors_are_loud_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.Config(device=0, batch=32, online_capacity=100, expert_capacity=100, bc_rows=50, max_steps=1,
                      activation=0, replicas=1, gamma=0.99, tau=0.005, noise_clamp=0.5, loss_lambda=1e4)
    h = ctypes.c_void_p()
    rc = lib.hirl_agent_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc < 0 and lib.hirl_last_error()          # no device -> error code + message, never a CPU path
    with pytest.raises(Exception):
        _lib.check(rc)

"""GPU: data-parallel large-batch mode (include/hirl_b200.h), both variants: "fused" (all-reduces inside the
update kernel over NVLink peer memory, CUDA IPC) and "nccl" (stage-range launches + ncclAllReduce, HIRL_DP_NCCL=1).  The
reference has no such path; the bar is "k shards == one GPU within summation-order tolerance", at the reference's two optimizer boundaries
(/root/reference/hirl/agents/HIRL.py:286-288 and :323-325) and the batch mean of the soft Q-filter (:303)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_pair(B, bcw, en, steps, dp):
    """The same seeded scenario on a plain stage-kernel agent and on a world-1 data-parallel agent."""
    from hirl4ucav_b200 import _lib
    from hirl4ucav_b200.agents.HIRL import Agent
    from oracle import scenarios
    sc = scenarios.make_scenario(max(1000, 4 * B), max(600, 3 * B))
    ag = scenarios.build_agent(Agent, sc, B, 0, **({"data_parallel": dp} if dp else {}))
    scenarios.seed_all(1)
    outs = [ag.learn(bcw, en, 0) for _ in range(steps)]
    return ag, outs, [ag.export_flat(i) for i in range(4)], ag.export_grads(_lib.NET_CRITIC), ag.export_grads(_lib.NET_ACTOR)


@pytest.mark.parametrize("variant", ["fused", "nccl"])
@pytest.mark.parametrize("B,bcw,en", [(64, 100, 16), (64, 0.5, 0), (512, 100, 128)])
def test_world_1_data_parallel_equals_the_plain_stage_kernel(B, bcw, en, variant, monkeypatch):
    """One rank: every reduction point runs (fused: exchange block, flag, rank-ordered sum over one rank inside the kernel;
    nccl: five launches over stage ranges with trivial all-reduces between them); the arithmetic is the stage kernel's, so
    the result must be bit-identical to the plain single launch."""
    monkeypatch.setenv("HIRL_TCC", "0")
    monkeypatch.setenv("HIRL_CLUSTER", "0")
    monkeypatch.setenv("HIRL_DP_NCCL", "1" if variant == "nccl" else "0")
    a0, o0, w0, gc0, ga0 = _run_pair(B, bcw, en, 6, None)
    assert a0.backend == "stage"
    a1, o1, w1, gc1, ga1 = _run_pair(B, bcw, en, 6, (0, 1))
    assert a1.dp_variant == variant
    for x, y in zip(o0, o1):
        assert [float(v) for v in x] == [float(v) for v in y]
    for x, y in zip(w0, w1):
        assert np.array_equal(x, y)
    assert np.array_equal(gc0, gc1) and np.array_equal(ga0, ga1)


def test_fused_variant_falls_back_to_nccl_when_ipc_is_unavailable(monkeypatch):
    """If the exchange blocks cannot be shared (CUDA IPC unavailable on some rank) every rank falls back to the NCCL variant."""
    monkeypatch.setenv("HIRL_TCC", "0")
    monkeypatch.setenv("HIRL_DP_NCCL", "0")
    monkeypatch.setenv("HIRL_DP_FUSED_FAIL", "1")
    with pytest.warns(UserWarning, match="falling back to the NCCL variant"):
        a1, o1, w1, gc1, ga1 = _run_pair(64, 100, 16, 4, (0, 1))
    assert a1.dp_variant == "nccl"
    monkeypatch.delenv("HIRL_DP_FUSED_FAIL")
    a0, o0, w0, gc0, ga0 = _run_pair(64, 100, 16, 4, None)
    for x, y in zip(w0, w1):
        assert np.array_equal(x, y)


def test_data_parallel_agent_moves_off_the_cluster_kernel():
    """An agent created at a batch the tensor-core cluster kernel would take is switched to the stage kernel on attach."""
    os.environ.pop("HIRL_TCC", None)
    from hirl4ucav_b200.agents.HIRL import Agent
    from oracle import scenarios
    sc = scenarios.make_scenario(1000, 600)
    ag = scenarios.build_agent(Agent, sc, 64, 0, data_parallel=(0, 1))
    assert ag.backend == "stage"
    out = ag.learn(100, 16, 0)
    assert np.isfinite([float(v) for v in out[:4]]).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("variant", ["fused", "nccl"])
def test_two_shards_match_one_gpu(variant):
    env = dict(os.environ, HIRL_TCC="0", HIRL_CLUSTER="0", HIRL_DP_NCCL="1" if variant == "nccl" else "0")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "dp_worker.py")],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "dp ok" in res.stdout

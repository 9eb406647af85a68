"""TEST / BENCH INFRASTRUCTURE -- times the UNMODIFIED reference ``hirl.agents.HIRL.Agent.learn`` (hirl/agents/HIRL.py:221-334,
including its own host-side sampling, hirl/utils/buffer.py:38-48) on the host cores.

Run as a subprocess by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg) with the files that
``make -C oracle _ref`` staged under ``oracle/_ref`` (git-ignored copy of the reference's own agent files; nothing is
read from /root/reference at run time).  CUDA is hidden: the reference would otherwise pick cuda:0 by itself
(HIRL.py:52,125,153).  Protocol = BASELINE.md section 3: >= 6 warm-up calls, whole 6-call cycles, median of the repeats.

    python oracle/ref_bench.py --batch 256 --bcw 100 --expert 64 --calls 300 --repeats 3
prints one JSON line.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

os.environ["CUDA_VISIBLE_DEVICES"] = ""
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_ref"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--bcw", type=float, default=100.0)
    ap.add_argument("--expert", type=int, default=64)
    ap.add_argument("--calls", type=int, default=300)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--warm", type=int, default=6)
    ap.add_argument("--online", type=int, default=30000)
    ap.add_argument("--experts", type=int, default=20000)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    import torch
    if a.threads:
        torch.set_num_threads(a.threads)
    from hirl.agents.HIRL import Agent          # the reference's own class (oracle/_ref)
    import hirl.agents.HIRL as ref_mod
    assert os.path.realpath(ref_mod.__file__).startswith(os.path.realpath(os.path.join(HERE, "_ref"))), ref_mod.__file__
    from oracle import scenarios
    sc = scenarios.make_scenario(a.online, a.experts)
    with contextlib.redirect_stdout(io.StringIO()):
        ag = scenarios.build_agent(Agent, sc, a.batch, 0)
    scenarios.seed_all(1)
    bcw = int(a.bcw) if a.bcw == 100 else a.bcw
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(a.warm):
            ag.learn(bcw, a.expert, 0)
        rates = []
        for _ in range(a.repeats):
            t0 = time.perf_counter()
            for _ in range(a.calls):
                ag.learn(bcw, a.expert, 0)
            rates.append(a.calls / (time.perf_counter() - t0))
    rates.sort()
    print(json.dumps({"updates_per_s": rates[len(rates) // 2], "all": rates, "threads": torch.get_num_threads(), "calls": a.calls,
                      "repeats": a.repeats, "batch": a.batch, "nproc": os.cpu_count(), "torch": torch.__version__}))


if __name__ == "__main__":
    main()

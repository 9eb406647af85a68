"""GPU: per-stage timeline of the persistent update kernel from in-kernel SM-clock stamps.

    python tools/stage_profile.py [--batch 256] [--bcw 100] [--expert 64] [--mhz 1965]
Prints, per stage: the slowest CTA's work time, the mean work time, the number of CTAs that had work,
and the barrier cost seen by the last arriver; per learn() call totals.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hirl4ucav_b200 import _lib, synth  # noqa: E402
from hirl4ucav_b200.agents.HIRL import Agent  # noqa: E402
from hirl4ucav_b200.utils.seed import set_seed  # noqa: E402

NAMES = ["start", "R0", "G1", "R1", "G2", "R2", "G3", "R3", "G4", "R4", "G5", "R5", "G6", "R6", "G7", "R7", "G8"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--bcw", type=float, default=100.0)
    ap.add_argument("--expert", type=int, default=64)
    ap.add_argument("--mhz", type=float, default=1965.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ranks", action="store_true", help="also print the slowest CTA ranks of every stage")
    a = ap.parse_args()
    es, ea = synth.expert_table(20000)
    set_seed(0)
    ag = Agent(1e-3, 1e-3, 13, 4, 256, 512, 0.005, 0.99, 10 ** 5, a.batch, True, "Harfang_GYM", es, ea, 0.5, True, max_fused_steps=64)
    ag.FUSE_SUB = 64   # one launch for all calls: the in-kernel stamps cover the first 12 steps of a launch
    synth.fill_agent(ag, 30000, expert_states=es, expert_actions=ea)
    lib = _lib.load()
    calls = [(a.bcw, a.expert, 0)] * 12
    ag.learn_many(calls)            # warm-up
    _lib.check(lib.hirl_debug_profile(ag._h, 1, None, 0))
    ag.learn_many(calls)
    n = 12 * 17 * 160 * 2
    buf = np.zeros(n, np.int64)
    _lib.check(lib.hirl_debug_profile(ag._h, 0, buf.ctypes.data_as(C.c_void_p), n))
    prof = buf.reshape(12, 17, 160, 2)
    grid = lib.hirl_agent_sm_count(ag._h)
    us = lambda cyc: cyc / a.mhz
    lines = []
    for step in (6, 7):          # one actor step, one critic-only step (steady state)
        p = prof[step, :, :grid, :]
        active = [s for s in range(1, 17) if p[s, 0, 0] != 0 and p[s, 0, 1] > p[s - 1 if s > 1 else 0, 0, 1] - 10 ** 12]
        lines.append(f"step {step}: stages present = {[NAMES[s] for s in active]}")
        prev = p[0, :, 1].copy()   # step start per CTA
        tot = 0.0
        for s in active:
            work = p[s, :, 0] - prev
            bar = p[s, :, 1] - p[s, :, 0]
            stage_total = (p[s, :, 1] - prev).max()
            busy = int((work > 0.25 * work.max()).sum())
            lines.append(f"  {NAMES[s]:3s} total {us(stage_total):7.2f} us | work max {us(work.max()):7.2f} mean {us(work.mean()):7.2f} "
                         f"busy CTAs {busy:3d} | barrier min {us(bar.min()):5.2f} us")
            if a.ranks:   # the slowest CTAs of the stage: which ranks are the stragglers
                top = np.argsort(-work)[:4]
                lines.append("        slowest ranks: " + ", ".join(f"{int(r)}: {us(work[r]):.2f}" for r in top)
                             + f" | median {us(np.median(work)):.2f}")
            tot += us(stage_total)
            prev = p[s, :, 1].copy()
        lines.append(f"  sum of stages: {tot:.1f} us")
    print("\n".join(lines))
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

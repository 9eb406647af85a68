"""GPU: phase timeline of the cluster-resident update kernel from in-kernel SM-clock stamps of CTA 0.

    python tools/cluster_profile.py [--batch 256] [--bcw 100] [--expert 64] [--mhz 1965]
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

LABEL = {0: "step start", 1: "A: gather + first layers + cluster sync", 2: "A: segment-1 GEMMs (TA [A AB BC])", 3: "A: LayerNorm2/heads of segment 1 (exchanges)",
         4: "A: first layers of target critics + cluster sync", 5: "A: segment-2 GEMMs (TC1 TC2 C1 C2)", 6: "A: LayerNorm2/heads/TD target (exchanges)",
         7: "A: head+LN2 backward, column sums, dz2 store", 8: "A: dH1 x2 + reduce-scatter", 24: "A: LayerNorm1 backward, flush -> phase end", 25: "grid barrier",
         10: "B: small second-layer params (partials + Adam)", 11: "B: first-layer columns + Adam", 26: "B: dW2 tiles + Adam -> phase end", 27: "grid barrier",
         13: "C: gather, first layers, sync, Q GEMMs", 14: "C: LN2/heads, count, LN2 backward", 15: "C: dH1 of Q1(s,pi) + reduce-scatter", 16: "C: LN1 backward, d pi exchange",
         17: "C: actor backward (both branches)", 28: "C: flush -> phase end", 29: "grid barrier",
         20: "D: small params", 21: "D: first-layer columns, Polyak(critic), scalars", 30: "D: dW2 tiles + Adam + Polyak -> phase end", 31: "grid barrier"}
ORDER = [0, 1, 2, 3, 4, 5, 6, 7, 8, 24, 25, 10, 11, 26, 27, 13, 14, 15, 16, 17, 28, 29, 20, 21, 30, 31]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--bcw", type=float, default=100.0)
    ap.add_argument("--expert", type=int, default=64)
    ap.add_argument("--mhz", type=float, default=1965.0)
    a = ap.parse_args()
    es, ea = synth.expert_table(20000)
    set_seed(0)
    ag = Agent(1e-3, 1e-3, 13, 4, 256, 512, 0.005, 0.99, 10 ** 5, a.batch, True, "Harfang_GYM", es, ea, 0.5, True, max_fused_steps=64)
    ag.FUSE_SUB = 64
    synth.fill_agent(ag, 30000, expert_states=es, expert_actions=ea)
    lib = _lib.load()
    calls = [(a.bcw, a.expert, 0)] * 12
    ag.learn_many(calls)
    _lib.check(lib.hirl_debug_profile(ag._h, 1, None, 0))
    ag.learn_many(calls)
    n = 12 * 17 * 160 * 2
    buf = np.zeros(n, np.int64)
    _lib.check(lib.hirl_debug_profile(ag._h, 0, buf.ctypes.data_as(C.c_void_p), n))
    prof = buf[:12 * 40].reshape(12, 40)
    for step in (6, 7):
        p = prof[step]
        print(f"step {step}:")
        prev = p[0]
        for i in ORDER[1:]:
            if p[i] == 0 or p[i] < p[0]:
                continue
            print(f"  {LABEL[i]:58s} {(p[i] - prev) / a.mhz:8.2f} us   (t = {(p[i] - p[0]) / a.mhz:8.2f})")
            prev = p[i]
    g = buf[12 * 40:12 * 40 + 12 * 4 * 16].reshape(12, 4, 16)
    for step in (6, 7):
        for ph in range(4):
            t = g[step, ph]
            t = t[t > 0]
            if t.size:
                print(f"step {step} phase {'ABCD'[ph]} end per cluster (us after the first): " + " ".join(f"{(x - t.min()) / 1e3:.1f}" for x in t))
    print("grid (CTAs):", lib.hirl_agent_sm_count(ag._h))


if __name__ == "__main__":
    main()

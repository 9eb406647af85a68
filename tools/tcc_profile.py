"""GPU: timeline of the tensor-core cluster kernel (csrc/tcc_update.cuh) from SM-clock stamps of cluster 0 / CTA 0.

    HIRL_TCC=1 python tools/tcc_profile.py [--batch 256] [--bcw 100] [--expert 64] [--mhz 1965]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HIRL_TCC", "1")
from hirl4ucav_b200 import _lib, synth  # noqa: E402
from hirl4ucav_b200.agents.HIRL import Agent  # noqa: E402
from hirl4ucav_b200.utils.seed import set_seed  # noqa: E402

ONLINE = [(17, "A/tower unit: gather"), (4, "A/tower unit: critic forward pass"), (38, "A/tower unit: park state, wait for a'"),
          (39, "A/tower unit: target critic pass -> tq published"), (36, "A/tower unit: wait for the partner tower's tq"),
          (37, "A/tower unit: TD target, loss, head/LN2 backward, small-param column sums"), (5, "A/tower unit: backward GEMMs, LN1 backward, partials")]
MAIN = [(1, "A/target-actor unit: gather"), (2, "A/target-actor unit: target actor pass -> a' published"),
        (6, "grid barrier (after the slower chain)"), (7, "B: critic Adam + images"), (8, "grid barrier"),
        (10, "C: gather, parked pi of the phase-A actor unit"), (11, "C: Q1(s,pi) pass"), (12, "C: soft count, Q backward -> d pi"),
        (14, "C: actor backward (RL branch)"), (15, "grid barrier"), (16, "D: actor Adam, Polyak, images + grid barrier")]
FWD = ["stream begin + param cache", "first layer of the own 8 rows -> H1 image in L2", "cluster barrier, multicast, MMA issue", "wait for the accumulator", "TMEM -> registers, local LN stats",
       "exchange 1 (LN statistics)", "normalise, head partial, exchange 2"]
BWD = ["dz2 tile", "publish + dH1 MMA issue", "H1^T recompute", "wait dH1, dz2^T tile, dW2 issue", "dH1 partial -> L2", "cluster barrier", "own rows: 8 partials, LN1 backward",
       "wait dW2, partial -> L2"]


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
    prof = buf[:12 * 96].reshape(12, 96)
    us = lambda d: d / a.mhz
    tot = [us(prof[i + 1][0] - prof[i][0]) for i in range(4, 10)]
    print("whole calls 4..9 (start to start, cluster 0): " + " ".join(f"{t:.1f}" for t in tot) + f" us   mean {sum(tot) / len(tot):.1f} us")
    for step in (6, 7):
        p = prof[step]
        print(f"step {step} ({'actor' if p[10] else 'critic-only'} call), cluster 0 / CTA 0:")
        print(f"  clusters: {p[19]}")
        prev = p[18]
        for i, label in ONLINE:      # stamped by the cluster of the first online unit (SM clocks of different SMs: offsets of a few hundred ns)
            if p[i] == 0:
                continue
            print(f"  {label:66s} {us(p[i] - prev):8.2f} us   (t = {us(p[i] - p[18]):8.2f})")
            prev = p[i]
        prev = p[0]
        for i, label in MAIN:
            if p[i] == 0:
                continue
            print(f"  {label:66s} {us(p[i] - prev):8.2f} us   (t = {us(p[i] - p[0]):8.2f})")
            prev = p[i]
        if p[55]:
            t0 = p[21]
            print("    target actor pass, after the first layer: cluster barrier +%.2f | H1 tile landed +%.2f | W2 unit c landed (issuer 0 past its wait): %s us" % (
                us(p[55] - t0), us(p[56] - t0), " ".join(f"+{us(p[57 + c] - t0):.2f}" for c in range(8))))
        for base, name, labels, start in ((20, "target actor pass", FWD, p[1]), (28, "critic forward pass", FWD, p[17]), (40, "critic backward", BWD, p[37])):
            print(f"    -- inside the {name}:")
            prev = start
            for j, label in enumerate(labels):
                if p[base + j] == 0:
                    continue
                print(f"       {label:58s} {us(p[base + j] - prev):8.2f} us")
                prev = p[base + j]


if __name__ == "__main__":
    main()

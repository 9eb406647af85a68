"""GPU diagnostic: runs the fused update stage by stage and prints, for every intermediate,
its error against the oracle (fp32) and the fp64 twin.  Not a test (pytest does not collect it); a bug locator.  It lives under tests/ because it executes the oracle.

    python tests/diag_update.py [--batch 32] [--steps 4]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hirl4ucav_b200 import _lib  # noqa: E402
from hirl4ucav_b200.agents.HIRL import Agent  # noqa: E402
from oracle import scenarios  # noqa: E402
from oracle.hirl_oracle import OracleAgent  # noqa: E402

P = dict(TA=0, C1=1, C2=2, TC1=3, TC2=4, A=5, AB=6, BC=7, QP=8, QS=9)
DB = dict(X=0, XB=1, H1=2, XH1=3, RS1=4, Z2=5, DH1=6, APRIME=7, PI=8, PIB=9, ABC=10, Y=11, Q=12, TQ=13, RLQ=14, SOFTQ=15, GPI=16, LP=17)


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def show(name, got, ref32, ref64=None):
    line = f"  {name:28s} rel(cuda,o32)={rel(got, ref32):.2e} maxabs={np.abs(np.asarray(got, np.float64).ravel() - np.asarray(ref32, np.float64).ravel()).max():.2e}"
    if ref64 is not None:
        line += f"  rel(cuda,f64)={rel(got, ref64):.2e} rel(o32,f64)={rel(ref32, ref64):.2e}"
    print(line)


def t2n(t):
    return t.detach().double().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--bcw", type=float, default=100)
    ap.add_argument("--expert", type=int, default=8)
    args = ap.parse_args()
    B = args.batch
    sc = scenarios.make_scenario(max(1000, 4 * B), max(600, 2 * B))
    orc = scenarios.build_agent(OracleAgent, sc, B, 0)
    o64 = scenarios.build_agent(OracleAgent, sc, B, 0, dtype=torch.float64)
    ag = scenarios.build_agent(Agent, sc, B, 0)
    for w, nid in (("actor", 0), ("targetActor", 1), ("critic", 2), ("targetCritic", 3), ("bc_actor", 4)):
        assert np.array_equal(ag.export_flat(nid).astype(np.float64), orc.flat(w)), f"init/import/export mismatch {w}"
    print("init + import/export round trip: bit-exact")
    zs = None
    scenarios.seed_all(1)
    for step in range(args.steps):
        rec32, rec64 = {}, {}
        actor_step = orc.actorTrainable
        out32 = orc.learn(args.bcw, args.expert, 0, rec=rec32)
        d = orc.last_draws
        forced = {"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]}
        out64 = o64.learn(args.bcw, args.expert, 0, forced=forced, rec=rec64)
        print(f"== step {step} (actor_step={actor_step}) oracle out32={[float(x) for x in out32]}")
        out = ag.learn(args.bcw, args.expert, 0, forced=forced)
        print(f"   cuda out   ={[float(x) for x in out]}")
        print(f"   fp64 out   ={[float(x) for x in out64]}")
        buf = lambda name, idx=0: ag.debug_buffer(DB[name], idx)
        nsz = lambda a, width: a.reshape(-1, B, width).sum(0)  # sum split-K partials
        X = buf("X").reshape(B, 32)
        show("X.state", X[:, :13], t2n(rec32["S"]))
        show("X.next_state", X[:, 17:30], t2n(rec32["S2"]))
        show("X.reward", X[:, 30], t2n(rec32["R"]))
        for nm, tag in (("TA", "ta."), ("C1", "critic.q1."), ("C2", "critic.q2."), ("TC1", "tc.q1."), ("TC2", "tc.q2.")):
            show(f"H1[{nm}]", buf("H1", P[nm]).reshape(B, 256), t2n(rec32[tag + "h1"]), t2n(rec64[tag + "h1"]))
        show("APRIME", buf("APRIME").reshape(B, 4), t2n(rec32["a_next"]), t2n(rec64["a_next"]))
        tq = buf("TQ").reshape(2, B)
        show("TQ1", tq[0], t2n(rec32["tq1"]), t2n(rec64["tq1"]))
        show("TQ2", tq[1], t2n(rec32["tq2"]), t2n(rec64["tq2"]))
        show("Y", buf("Y"), t2n(rec32["y"]), t2n(rec64["y"]))
        q = buf("Q").reshape(2, B)
        show("Q1", q[0], t2n(rec32["q1"]), t2n(rec64["q1"]))
        show("Q2", q[1], t2n(rec32["q2"]), t2n(rec64["q2"]))
        if not actor_step and os.environ.get("HIRL_CLUSTER") == "0":
            # Z2 of critic passes now hold dz2 (grad of z2); target passes still hold z2
            show("Z2[TC1]=z2", nsz(buf("Z2", P["TC1"]), 512), t2n(rec32["tc.q1.z2"]), t2n(rec64["tc.q1.z2"]))
            show("Z2[C1]=dz2 (split0)", buf("Z2", P["C1"]).reshape(-1, B, 512)[0], t2n(rec32["critic.q1.z2"].grad), t2n(rec64["critic.q1.z2"].grad))
            show("DH1[C1]=dh1", nsz(buf("DH1", 0), 256), t2n(rec32["critic.q1.h1"].grad), t2n(rec64["critic.q1.h1"].grad))
        g = ag.export_grads(_lib.NET_CRITIC)
        g32 = np.concatenate([v.double().reshape(-1).numpy() for v in rec32["critic.grads"].values()])
        g64 = np.concatenate([v.double().reshape(-1).numpy() for v in rec64["critic.grads"].values()])
        show("critic grads (all)", g, g32, g64)
        off = 0
        for k, v in rec32["critic.grads"].items():
            n = v.numel()
            show(f"   grad {k}", g[off:off + n], g32[off:off + n], g64[off:off + n])
            off += n
        if actor_step:
            show("PI", buf("PI").reshape(B, 4), t2n(rec32["pi"]), t2n(rec64["pi"]))
            show("PIB", buf("PIB").reshape(B, 4), t2n(rec32["pi_bc"]), t2n(rec64["pi_bc"]))
            if "soft_q" in rec32:
                show("ABC", buf("ABC").reshape(B, 4), t2n(rec32["a_bc"]), t2n(rec64["a_bc"])) if "a_bc" in rec32 else None
            show("RLQ", buf("RLQ"), t2n(rec32["rl_q_all"]), t2n(rec64["rl_q_all"]))
            if "soft_q" in rec32:
                show("SOFTQ", buf("SOFTQ"), t2n(rec32["soft_q"]), t2n(rec64["soft_q"]))
                print(f"   soft margin min|softQ-rlQ| = {rec32['soft_count_margin']:.3e}")
            if os.environ.get("HIRL_CLUSTER") != "0":
                rs2 = ag.debug_buffer(20, 0); print("   RS2[A] head", rs2[:6], " RS2[AB]", ag.debug_buffer(20, 1)[:6])
                xh2 = ag.debug_buffer(19, 0).reshape(-1, 512); print("   XH2[A] row0 mean/var", xh2[0].mean(), xh2[0].var(), " row20", xh2[20].mean(), xh2[20].var())
                xh2b = ag.debug_buffer(19, 1).reshape(-1, 512); print("   XH2[AB] row0 mean/var", xh2b[0].mean(), xh2b[0].var())
                for sl in (2, 3):
                    dz = ag.debug_buffer(21, sl); print(f"   DZ2S[{sl}] absmax {np.abs(dz).max():.3e} nonzero {np.count_nonzero(dz)}/{dz.size}")
                    print(f"   DZ1[{sl}] absmax {np.abs(ag.debug_buffer(22, sl)).max():.3e}  DY1 absmax {np.abs(ag.debug_buffer(23, sl)).max():.3e}")
            ga = ag.export_grads(_lib.NET_ACTOR)
            a32 = np.concatenate([v.double().reshape(-1).numpy() for v in rec32["actor.grads"].values()])
            a64 = np.concatenate([v.double().reshape(-1).numpy() for v in rec64["actor.grads"].values()])
            show("actor grads (all)", ga, a32, a64)
            off = 0
            for k, v in rec32["actor.grads"].items():
                n = v.numel()
                show(f"   grad {k}", ga[off:off + n], a32[off:off + n], a64[off:off + n])
                off += n
        for w, nid in (("actor", 0), ("targetActor", 1), ("critic", 2), ("targetCritic", 3)):
            show(f"weights {w}", ag.export_flat(nid), orc.flat(w), o64.flat(w))
        m, v = ag.export_adam(_lib.NET_CRITIC)
        m32, v32 = orc.adam_state("critic")
        show("adam m critic", m, m32)
        show("adam v critic", v, v32)
        # teacher forcing: continue every implementation from the fp32 oracle's state
        for w, nid in (("actor", 0), ("targetActor", 1), ("critic", 2), ("targetCritic", 3)):
            getattr(ag, w).load_state_dict(orc.nets()[w].state_dict())
            o64.nets()[w].load_state_dict({k: t.double() for k, t in orc.nets()[w].state_dict().items()})
        ag.push_parameters()
        for w, nid in (("actor", _lib.NET_ACTOR), ("critic", _lib.NET_CRITIC)):
            mm, vv = orc.adam_state(w)
            if orc.nets()[w].optimizer.state:
                ag.import_adam(nid, mm, vv, ag._actor_t if w == "actor" else ag._critic_t)
                for p32, p64 in zip(orc.nets()[w].parameters(), o64.nets()[w].parameters()):
                    s32, s64 = orc.nets()[w].optimizer.state[p32], o64.nets()[w].optimizer.state[p64]
                    s64["exp_avg"].copy_(s32["exp_avg"].double())
                    s64["exp_avg_sq"].copy_(s32["exp_avg_sq"].double())


if __name__ == "__main__":
    main()

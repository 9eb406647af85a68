"""torchrun worker of tests/test_gpu_dp.py::test_two_shards_match_one_gpu (also runnable by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dp_worker.py)

Every rank builds the SAME seeded agent with the global batch and data_parallel=(rank, world); rank 0 also runs a plain
one-GPU agent of the global batch on the same streams (both free-running).  Per call the returned scalars must agree to
2e-4 relative and the gradients of both networks to 1e-4 rel-L2 (summation order differs: sum of shard sums, and the
difference feeds back through the weights; the first call agrees to ~1e-6); after the last call every rank must hold
bit-identical weights (they all apply the same all-reduced gradient)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hirl4ucav_b200 import _lib  # noqa: E402
from hirl4ucav_b200.agents.HIRL import Agent  # noqa: E402
from oracle import scenarios  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(1e-30, np.linalg.norm(b.astype(np.float64))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    worst = 0.0
    for B, bcw, en, steps in ((512, 100, 128, 6), (4096, 0.5, 1024, 4), (256, 100, 0, 4)):
        sc = scenarios.make_scenario(max(1000, 2 * B), max(600, 2 * B))
        dp = scenarios.build_agent(Agent, sc, B, 0, data_parallel=(rank, world))
        ref = scenarios.build_agent(Agent, sc, B, 0) if rank == 0 else None
        scenarios.seed_all(1)
        for t in range(steps):
            out = dp.learn(bcw, en, 0)
            d = dp.last_draws
            if rank == 0:
                want = ref.learn(bcw, en, 0, forced={"online": d["online"], "expert": d["expert"], "bc": d["bc"], "noise": d["noise"]})
                for j in range(4):
                    a, b = float(out[j]), float(want[j])
                    assert abs(a - b) <= 2e-4 * abs(b) + 1e-6, (B, t, j, a, b)
                assert abs(float(out[5]) - float(want[5])) <= 1.0 / B + 1e-6, (B, t, out[5], want[5])   # Q-filter weight: one tie at most
                for net in (_lib.NET_CRITIC, _lib.NET_ACTOR):
                    if net == _lib.NET_ACTOR and (t % 2):
                        continue
                    e = rel(dp.export_grads(net), ref.export_grads(net))
                    worst = max(worst, e)
                    assert e <= 1e-4, (B, t, net, e)
        # every rank must end with identical weights
        mine = np.concatenate([dp.export_flat(i) for i in range(4)])
        box = [None] * world
        dist.all_gather_object(box, mine.tobytes())
        assert all(b == box[0] for b in box), f"ranks diverged at B={B}"
        del dp, ref
    if rank == 0:
        print(f"dp ok: world {world}, worst gradient rel-L2 vs one GPU {worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU, world_size 2, gloo: the host logic of the N>1 path (sharding, reductions, record gather)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hirl4ucav_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_replica_seeds_do_not_depend_on_world_size():
    one = parallel.replica_seeds(100, 1024, 1, 0)
    for world in (2, 4, 8):
        many = sum((parallel.replica_seeds(100, 1024, world, r) for r in range(world)), [])
        assert many == one


def test_shard_step_reassembles_the_global_batch():
    rng = np.random.default_rng(3)
    for B, n_ex in ((8, 3), (256, 64), (4096, 1024), (64, 0), (64, 63)):
        n_on = B - n_ex
        on, ex, bc = rng.integers(0, 10 ** 5, n_on), rng.integers(0, 10 ** 4, n_ex), rng.permutation(20000)[:B]
        for world in (1, 2, 4, 8):
            rows, bcs = [], []
            for r in range(world):
                a, b, o, e, c = parallel.shard_step(n_on, n_ex, on, ex, bc, r, world)
                assert a + b == B // world == len(c) and len(o) == a and len(e) == b
                rows += [("on", int(i)) for i in o] + [("ex", int(i)) for i in e]
                bcs += list(c)
            assert rows == [("on", int(i)) for i in on] + [("ex", int(i)) for i in ex]     # the reference's torch.cat order (HIRL.py:229-233)
            assert bcs == list(bc)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.env_rank() == (rank, rank, world)
        # replicas: each rank "runs" its shard; the per-replica records are identical to a 1-rank run
        seeds = parallel.replica_seeds(7, 10, world, rank)
        recs = [{"seed": s, "value": float(np.random.default_rng(s).random())} for s in seeds]
        allr = sum(parallel.gather_records(recs), [])
        t_max = parallel.max_over_ranks(1.0 + rank)
        total = parallel.sum_over_ranks(len(seeds))
        # env shards: contiguous chunks, checksum of checksums equals the global checksum
        x = np.arange(1001, dtype=np.int64)
        a, b = parallel.shard_range(x.size, world, rank)
        chk = parallel.sum_over_ranks(float(x[a:b].sum()))
        # data-parallel mode: every rank draws the same streams (same seed) and keeps its rows; the union is the global batch
        g = np.random.default_rng(11)
        on, ex, bc = g.integers(0, 1000, 48), g.integers(0, 100, 16), g.permutation(500)[:64]
        n_on, n_ex, o, e, c = parallel.shard_step(48, 16, on, ex, bc, rank, world)
        shards = parallel.gather_records((n_on, n_ex, [int(i) for i in o], [int(i) for i in e], [int(i) for i in c]))
        assert sum(s_[0] for s_ in shards) == 48 and sum(s_[1] for s_ in shards) == 16
        assert sum((s_[2] for s_ in shards), []) == [int(i) for i in on] and sum((s_[3] for s_ in shards), []) == [int(i) for i in ex]
        assert sum((s_[4] for s_ in shards), []) == [int(i) for i in bc]
        q.put((rank, [r["seed"] for r in allr], [r["value"] for r in allr], t_max, total, chk))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_run_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single_seeds = parallel.replica_seeds(7, 10, 1, 0)
    single_vals = [float(np.random.default_rng(s).random()) for s in single_seeds]
    for rank, seeds, vals, t_max, total, chk in out:
        assert seeds == single_seeds and vals == single_vals      # replica i is the same wherever it runs
        assert t_max == 2.0 and total == 10.0 and chk == float(np.arange(1001).sum())

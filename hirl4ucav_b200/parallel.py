"""Host-side sharding for multi-GPU runs (one process per GPU, launched by torchrun).

The hot path shards only where it is naturally independent:
  * agent updates: independent replicas / seeds, R/k per GPU, NO data-path collective;
  * env pass: contiguous chunks of N/k states per GPU, no exchange;
  * data-parallel large-batch mode (one agent, global batch cut into equal row shards): every rank draws the same index
    streams and keeps its rows (shard_step); the gradient all-reduces live in the C library (hirl_allreduce_init).
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used for the timing barrier, the
max-over-ranks reduction and gathering small result records -- plumbing, not the product.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_range(n: int, world: int, rank: int):
    """Contiguous [start, stop) of n items for `rank`; sizes differ by at most one, order preserved."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_step(n_online: int, n_expert: int, idx_online, idx_expert, idx_bc, rank: int, world: int):
    """Rows [rank * B/world, (rank + 1) * B/world) of one learn() call's batch of B = n_online + n_expert rows.

    The batch is the reference's torch.cat((online batch, expert batch)) (HIRL.py:229-233): row b < n_online is
    online[idx_online[b]], the others expert[idx_expert[b - n_online]]; the cloning batch (HIRL.py:249-251) is row-aligned
    with it.  Returns (n_online_loc, n_expert_loc, idx_online_loc, idx_expert_loc, idx_bc_loc): concatenating the shards
    of all ranks in rank order gives back the global batch row for row."""
    B = int(n_online) + int(n_expert)
    if B % world:
        raise ValueError("the global batch must be a multiple of the world size")
    lo, hi = shard_range(B, world, rank)
    on_lo, on_hi = min(lo, n_online), min(hi, n_online)
    ex_lo, ex_hi = max(lo, n_online) - n_online, max(hi, n_online) - n_online
    return on_hi - on_lo, ex_hi - ex_lo, idx_online[on_lo:on_hi], idx_expert[ex_lo:ex_hi], idx_bc[lo:hi]


def replica_seeds(base_seed: int, n_replicas: int, world: int, rank: int):
    """Seeds of the replicas this rank owns: replica i always gets base_seed + i, wherever it runs,
    so a k-GPU run reproduces the 1-GPU run replica by replica."""
    a, b = shard_range(n_replicas, world, rank)
    return [base_seed + i for i in range(a, b)]


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_records(record):
    """List of every rank's small Python record on every rank (rank order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [record]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, record)
    return out

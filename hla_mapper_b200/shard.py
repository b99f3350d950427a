"""Pair sharding across the GPUs of one box.

Every read pair is independent through screen -> trim -> locus mask -> score -> assign
(/root/reference/src/map_dna.cpp:917-1029 only touches keys of its own read id), so ranks own
contiguous, BLOCK-aligned shards of the pair stream, the database is replicated per GPU and the
only inter-rank traffic is the host-side gather of the per-pair results.  No data-path collective.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_pairs: int, world: int, block: int = 1):
    """[lo, hi) of every rank: contiguous, block-aligned (except the last), covering [0, n)"""
    nblk = -(-n_pairs // block)
    per = -(-nblk // world)
    out = []
    for r in range(world):
        lo = min(n_pairs, r * per * block)
        hi = min(n_pairs, (r + 1) * per * block)
        out.append((lo, hi))
    return out


def gather_to_rank0(dist, local: dict, bounds, rank: int, world: int):
    """host-side, order-preserving gather of per-pair result arrays (a dict of numpy arrays whose
    first axis is the rank's shard) to rank 0 over torch.distributed (gloo or nccl-backed object
    gather); returns the concatenated dict on rank 0, None elsewhere"""
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0)
    if rank != 0:
        return None
    out = {}
    for k in local:
        out[k] = np.concatenate([parts[r][k] for r in range(world)], axis=0)
        assert out[k].shape[0] == bounds[-1][1]
    return out


def max_over_ranks(dist, value: float, device=None) -> float:
    """a device time is the max over ranks (never wall clock of one rank)"""
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in
the CPU tests).  The eval path shards by impression — impressions are independent, every rank holds
a replica of the news-vector cache — and the only collective is one all-reduce of the five metric
partial sums.  The reference itself evaluates on rank 0 only (trainer.py:342).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from .synth import Impressions


def init_from_env(backend=None):
    """Join the process group torchrun described in the environment (RANK / WORLD_SIZE /
    LOCAL_RANK / MASTER_*).  Returns (rank, world_size, local_rank); world_size 1 -> no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(cand_off, max_history, world_size):
    """Contiguous impression ranges balanced by the work sum(H + C_i) (the bytes an impression makes
    the scoring kernel read).  Returns int64 [world_size + 1] impression boundaries."""
    cand_off = np.asarray(cand_off, np.int64)
    n = cand_off.shape[0] - 1
    work = np.cumsum(np.diff(cand_off) + max_history)
    total = work[-1] if n else 0
    bounds = np.zeros(world_size + 1, np.int64)
    for r in range(1, world_size):
        bounds[r] = np.searchsorted(work, total * r / world_size, side="left")
    bounds[world_size] = n
    return np.maximum.accumulate(bounds)


def shard_impressions(imp: Impressions, rank, world_size):
    """(this rank's impressions, global index of its first pair, total pairs).  The global pair
    index keeps the reference's mini-batch boundaries — hence the GraphSAGE prefix of every pair —
    independent of the sharding."""
    b = shard_bounds(imp.cand_off, imp.hist_news.shape[1], world_size)
    lo, hi = int(b[rank]), int(b[rank + 1])
    return imp.slice(lo, hi), int(imp.cand_off[lo]), int(imp.cand_off[-1])


def all_reduce_sums(sums, group=None):
    """SUM all-reduce of the fp64 [5] metric partial sums; returns the 4 global means."""
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    s = sums.tolist()
    return tuple((x / s[4]) if s[4] > 0 else float("nan") for x in s[:4])

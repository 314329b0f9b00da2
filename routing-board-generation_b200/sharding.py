"""Multi-GPU plumbing: boards / envs are independent, so a batch shards by
contiguous key slices (the reference's pmap layout, rl_training/setup_train.py:397-400)
with no data-path collective.  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is used only to gather benchmark statistics."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [offset, offset+count) of `total` units for `rank`; earlier ranks take the remainder."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def shard_keys(key, total: int, rank: int, world: int) -> torch.Tensor:
    """This rank's rows of jax.random.split(key, total), derived on the device (no host scatter)."""
    from . import engine

    offset, count = shard_bounds(total, rank, world)
    return engine.split(key, total, offset, count)


def gather_stats(local: Dict[str, float], device=None) -> Dict[str, Dict[str, float]]:
    """all_reduce a flat dict of counters: returns {'sum': {...}, 'max': {...}, 'min': {...}}."""
    names = sorted(local)
    vals = torch.tensor([float(local[n]) for n in names], dtype=torch.float64, device=device)
    out = {}
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for opname, op in (("sum", dist.ReduceOp.SUM), ("max", dist.ReduceOp.MAX), ("min", dist.ReduceOp.MIN)):
            t = vals.clone()
            dist.all_reduce(t, op=op)
            out[opname] = dict(zip(names, t.tolist()))
    else:
        out = {k: dict(zip(names, vals.tolist())) for k in ("sum", "max", "min")}
    return out

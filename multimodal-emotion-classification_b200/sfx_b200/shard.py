"""Clip-sharded data parallelism over the GPUs of one box.

Clips are independent units: each rank extracts its own contiguous range with no data-path collective.
The only exchange is the optional all-gather of the [N, 56] float32 feature matrix into the training
feature cache (the in-memory ``X`` of model_training/train_speech_model.py:151), 224 B per clip, done with
``torch.distributed.all_gather_into_tensor`` (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_size(n_total: int, world: int) -> int:
    """Rows per rank when N is padded up to a multiple of the world size."""
    return -(-n_total // world)


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous range [lo, hi) of clips owned by `rank` (equal shards, the last ones may be short/empty)."""
    per = shard_size(n_total, world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def partition_by_samples(lengths, world: int):
    """Contiguous ranges balanced by SAMPLE count (variable-length batches): list of (lo, hi) per rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    csum = np.concatenate(([0], np.cumsum(lengths)))
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * r / world, side="left")))
    cuts.append(len(lengths))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(lengths)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def gather_feature_cache(local_feats: torch.Tensor, n_total: int, group=None, out: torch.Tensor | None = None,
                         async_op: bool = False):
    """All-gather equal-size shards of feature rows into the full [n_total, F] cache on every rank.

    `local_feats` holds this rank's rows of shard_range(n_total, world, rank); shorter trailing shards are
    padded to shard_size rows for the collective and the padding is dropped afterwards.

    `out` ([shard_size * world, F]) reuses a cache buffer.  With ``async_op=True`` the call returns ``(cache, work)`` at
    once: the collective runs on the communication stream underneath the next batch's extraction, and
    ``work.wait()`` orders the current stream after it (call it before `local_feats` or `out` are overwritten)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = shard_size(n_total, world)
    lo, hi = shard_range(n_total, world, rank)
    if local_feats.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: expected {hi - lo} rows, got {local_feats.shape[0]}")
    width = local_feats.shape[1]
    send = local_feats
    if hi - lo < per:
        send = torch.zeros((per, width), dtype=local_feats.dtype, device=local_feats.device)
        send[:hi - lo] = local_feats
    full = out if out is not None else torch.empty((per * world, width), dtype=local_feats.dtype, device=local_feats.device)
    if full.shape != (per * world, width):
        raise ValueError(f"out must have shape {(per * world, width)}")
    work = dist.all_gather_into_tensor(full, send.contiguous(), group=group, async_op=async_op)
    return (full[:n_total], work) if async_op else full[:n_total]


def gather_ragged_feature_cache(local_feats: torch.Tensor, ranges, group=None) -> torch.Tensor:
    """All-gather for unequal shards (partition_by_samples): pads every shard to the largest one."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [hi - lo for lo, hi in ranges]
    per = max(max(sizes), 1)
    width = local_feats.shape[1]
    send = torch.zeros((per, width), dtype=local_feats.dtype, device=local_feats.device)
    send[:sizes[rank]] = local_feats
    full = torch.empty((per * world, width), dtype=local_feats.dtype, device=local_feats.device)
    dist.all_gather_into_tensor(full, send, group=group)
    return torch.cat([full[r * per:r * per + sizes[r]] for r in range(world)], dim=0)


def extract_sharded(extract_fn, waves: torch.Tensor, lengths=None, group=None, gather: bool = True):
    """Run `extract_fn(waves[lo:hi], lengths[lo:hi])` on this rank's shard of a batch every rank can index
    (e.g. a memory-mapped or replicated clip table), then optionally all-gather the feature cache."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_total = waves.shape[0]
    lo, hi = shard_range(n_total, world, rank)
    local = extract_fn(waves[lo:hi], None if lengths is None else lengths[lo:hi])
    return gather_feature_cache(local, n_total, group) if gather else local

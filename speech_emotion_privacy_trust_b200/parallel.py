"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the B200 box, gloo in CPU tests).

Extraction shards by utterance and needs no collective.  Cloak + GRL training is data parallel with one
flat-bucket gradient all-reduce per step; per-speaker statistics need one all-gather of the utterance partials' merge.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_by_length(lengths: Sequence[int], world_size: int, bucket: int = 8000) -> list[np.ndarray]:
    """Length-bucketed, balanced assignment of utterances to ranks.

    Utterances are sorted into buckets of `bucket` samples (0.5 s at 16 kHz), then dealt longest-first to the rank
    with the least samples so far; every rank gets the same mix of lengths and (almost) the same total audio.
    Returns, per rank, the utterance indices in ascending order."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.lexsort((np.arange(len(lengths)), -(lengths // bucket)))
    load = np.zeros(world_size, dtype=np.int64)
    parts: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += lengths[i]
    return [np.sort(np.asarray(p, dtype=np.int64)) for p in parts]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> int:
    """All-reduce the gradients of `params` as ONE flat fp32 bucket (5 MB for the GRL model: latency-bound on NVLink,
    so a single collective beats per-tensor calls).  Parameters without a gradient contribute zeros so every rank
    reduces the same layout.  Returns the number of elements reduced."""
    params = [p for p in params if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        chunk = flat[off:off + n].view_as(g).to(g.dtype)
        if p.grad is None:
            p.grad = chunk.clone()
        else:
            p.grad.copy_(chunk)
        off += n
    return off


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def merge_speaker_partials(count: torch.Tensor, mean: torch.Tensor, m2: torch.Tensor, mn: torch.Tensor, mx: torch.Tensor,
                           group=None):
    """Chan merge of per-rank per-speaker (count, mean, M2, min, max) partials, each (n_spk, F) with the same speaker
    rows on every rank: all-gather, then merge in rank order (deterministic).  Returns merged (count, mean, std, min, max)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        packed = torch.stack([count, mean, m2, mn, mx])
        gathered = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(gathered, packed, group=group)
    else:
        gathered = [torch.stack([count, mean, m2, mn, mx])]
    n, mu, s2, lo, hi = (gathered[0][i].clone() for i in range(5))
    for part in gathered[1:]:
        nb, mb, sb, lb, hb = (part[i] for i in range(5))
        tot = n + nb
        safe = torch.where(tot > 0, tot, torch.ones_like(tot))
        d = mb - mu
        mu = mu + d * (nb / safe)
        s2 = s2 + sb + d * d * (n * nb / safe)
        n = tot
        lo, hi = torch.minimum(lo, lb), torch.maximum(hi, hb)
    std = torch.sqrt(s2 / torch.where(n > 0, n, torch.ones_like(n)))
    return n, mu, std, lo, hi

"""
Contig sharding across GPUs (one process per GPU) -- the role of the reference's thread bin packing
(/root/reference/src/PanDepth.cpp:4415-4503: contigs sorted by length, greedily filled bins, one worker per bin).

Depth is additive and contigs are independent, so there is no data-path collective: every rank accumulates and
reduces only the contigs it owns; the single exchange is an all-gather of the per-contig (Length, CoveredSite,
TotalDepth) uint64 triples so that rank 0 can print the per-chromosome rows in tid order and the
`##RegionLength` summary row (PanDepth.cpp:5119-5122).

Range sharding (second form): contigs rarely divide evenly over GPUs (9 equal contigs on 8 GPUs
leave one GPU with twice the work under any contig assignment), but depth is additive over bases, so the concatenated
genome is cut into world_size equal ranges instead: rank r owns [r*G/N, (r+1)*G/N), which is a suffix of one contig, some
whole contigs and a prefix of another (pd_create_ranges).  Reads overlapping a boundary go to both neighbours and each engine
clips to its range; every rank reduces ALL intervals clipped to what it owns, and one all-reduce (sum) of the cov / sum
arrays gives every rank the full result.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def assign_contigs(lengths: Sequence[int], world_size: int) -> List[int]:
    """LPT bin packing: longest contig first onto the currently lightest rank. Returns owner rank per tid."""
    order = sorted(range(len(lengths)), key=lambda t: (-int(lengths[t]), t))
    load = [0] * world_size
    owner = [0] * len(lengths)
    for t in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[t] = r
        load[r] += int(lengths[t])
    return owner


def owned_mask(owner: Sequence[int], rank: int) -> np.ndarray:
    return (np.asarray(owner) == rank).astype(np.uint8)


def gather_triples(local: np.ndarray, owner: Sequence[int], rank: int, world_size: int, device=None) -> np.ndarray:
    """All-gather the dense [n_contigs, 3] uint64 table (rows of contigs a rank does not own are zero) and
    combine by ownership.  Uses torch.distributed (NCCL on GPUs, gloo on CPU); 24 bytes per contig per rank."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint64).view(np.int64))
    if device is not None:
        t = t.to(device)
    if world_size == 1 or not dist.is_initialized():
        return local.copy()
    out = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(out, t)
    res = np.zeros_like(local, dtype=np.uint64)
    own = np.asarray(owner)
    for r in range(world_size):
        part = out[r].cpu().numpy().view(np.uint64)
        res[own == r] = part[own == r]
    return res


def assign_ranges(lengths: Sequence[int], world_size: int):
    """Equal cut of the concatenated genome: -> (own_beg, own_end), each [world_size, n_contigs] uint32 (end <= beg: not owned)."""
    lens = np.asarray(lengths, dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(lens)])
    G = int(start[-1])
    beg = np.zeros((world_size, lens.size), np.uint32)
    end = np.zeros((world_size, lens.size), np.uint32)
    for r in range(world_size):
        a, z = G * r // world_size, G * (r + 1) // world_size
        for t in range(lens.size):
            lo, hi = max(a, int(start[t])), min(z, int(start[t + 1]))
            if hi > lo:
                beg[r, t] = lo - int(start[t])
                end[r, t] = hi - int(start[t])
    return beg, end


def allreduce_sum_u64(local: np.ndarray, world_size: int, device=None) -> np.ndarray:
    """Sum a uint64 array over ranks (values stay below 2^63: counts of bases / depth sums).  NCCL on GPUs, gloo on CPU."""
    import torch
    import torch.distributed as dist

    if world_size == 1 or not dist.is_initialized():
        return local.copy()
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.uint64).reshape(local.shape)

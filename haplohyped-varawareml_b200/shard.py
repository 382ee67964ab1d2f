"""Multi-GPU sharding of the conversion path.

The path partitions into independent units -- chromosome files -- so ranks never exchange genotype
data.  (Splitting ONE file over GPUs by BGZF-member range is not built: the HDF5 chunk that straddles two
ranks' rows needs a boundary-row exchange)  The only collective is an all_gather of a few int64 of
per-shard index metadata (record counts, first/last position, byte counts), which turns local
record indices into global chunk ranges.  One process per GPU; `torch.distributed` is plumbing
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist

META_FIELDS = ("n_records", "n_lines", "text_bytes", "first_pos", "last_pos", "out_bytes")


def plan_shards(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time bin packing of units (e.g. chromosome files by compressed size) onto
    `world` ranks.  Returns, per rank, the unit indices it owns (each sorted ascending)."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += sizes[i]
    return [sorted(b) for b in bins]


def gather_metadata(meta: Dict[str, int], device=None) -> List[Dict[str, int]]:
    """all_gather of the per-shard metadata; returns one dict per rank (rank order)."""
    vals = torch.tensor([int(meta.get(k, 0)) for k in META_FIELDS], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [dict(zip(META_FIELDS, vals.tolist()))]
    out = [torch.zeros_like(vals) for _ in range(dist.get_world_size())]
    dist.all_gather(out, vals)
    return [dict(zip(META_FIELDS, t.tolist())) for t in out]


def global_row_offsets(gathered: List[Dict[str, int]]) -> List[int]:
    """Exclusive prefix sum of n_records over ranks: where each shard's records start globally."""
    offs, acc = [], 0
    for g in gathered:
        offs.append(acc)
        acc += g["n_records"]
    return offs

"""Stream sharding across the GPUs of one box: streams are independent, so rank r owns a contiguous block of
stream ids and no data-path collective exists.  torch.distributed is used only to
gather finished transcripts / timings (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple


def shard_range(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) block of global stream ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(stream: int, n_streams: int, world: int) -> int:
    for r in range(world):
        lo, hi = shard_range(n_streams, world, r)
        if lo <= stream < hi:
            return r
    raise ValueError(stream)


def gather_results(local: Dict[int, Any], group=None) -> Dict[int, Any]:
    """All ranks receive {global stream id: result}.  The only exchange of a run (a few KB per stream)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    parts: List[Dict[int, Any]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    out: Dict[int, Any] = {}
    for p in parts:
        dup = set(out) & set(p)
        if dup:
            raise RuntimeError(f"streams owned by two ranks: {sorted(dup)[:5]}")
        out.update(p)
    return out

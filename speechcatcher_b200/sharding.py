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


def gather_beams(ctl, yseq, xpos, score, device=None, group=None):
    """The one exchange of a sharded run as fixed-width records: every rank contributes the
    beams of its S streams -- ctl [S, 2] int32 (n_hyp, len), yseq / xpos [S, B, L] int32, score [S, B] float64 -- and
    receives the same four arrays for all world * S streams, rank r's block at [r * S, (r + 1) * S).  Tensors travel
    as they are (NCCL all-gather over NVLink on CUDA tensors, gloo on CPU tensors); nothing is pickled."""
    import torch
    import torch.distributed as dist
    parts = [torch.as_tensor(ctl, dtype=torch.int32), torch.as_tensor(yseq, dtype=torch.int32),
             torch.as_tensor(xpos, dtype=torch.int32), torch.as_tensor(score, dtype=torch.float64)]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tuple(p.cpu().numpy() for p in parts)
    world = dist.get_world_size(group)
    out = []
    for p in parts:
        p = p.contiguous()
        if device is not None:
            p = p.to(device)
        full = torch.empty((world * p.shape[0],) + tuple(p.shape[1:]), dtype=p.dtype, device=p.device)
        dist.all_gather_into_tensor(full, p, group=group)
        out.append(full.cpu().numpy())
    return tuple(out)

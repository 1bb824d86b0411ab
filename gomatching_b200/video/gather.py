"""The one exchange step of the video path: per-chunk gather of frame records to the tracker rank.

One ``torch.distributed.gather`` of a (slots, stride) uint8 tensor per chunk -- NCCL over NVLink on the GPU box
(<= 0.49 MB per frame at 100 queries, latency-bound), gloo in the CPU tests.  No all-reduce, no all-to-all: the
spotting path itself has no collective.  world_size 1 degenerates to a no-op so the single-GPU path runs the same code.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .sharding import slot_of_frame, slots_per_rank


def gather_records(local: torch.Tensor, n_frames: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """local: (slots_per_rank, stride) uint8 on this rank.  Returns on ``dst`` the records in FRAME ORDER,
    shape (n_frames, stride); None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local[:n_frames]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    slots = slots_per_rank(n_frames, world)
    if local.shape[0] != slots:
        raise ValueError("rank %d contributes %d slots, expected %d" % (rank, local.shape[0], slots))
    local = local.contiguous()
    bucket: Optional[List[torch.Tensor]] = None
    out = None
    if rank == dst:
        out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        bucket = list(out.unbind(0))                       # the ranks' blocks land side by side: no stack afterwards
    dist.gather(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    return out.view(world * slots, -1).index_select(0, _frame_order(n_frames, world, slots, local.device))


_order_cache: dict = {}


def _frame_order(n_frames: int, world: int, slots: int, device) -> torch.Tensor:
    """Row of the (world * slots)-row gather buffer that holds frame t, for t = 0 .. n_frames-1 (cached per device:
    the tracker rank reorders every chunk with one index_select and no host work)."""
    key = (n_frames, world, slots, str(device))
    perm = _order_cache.get(key)
    if perm is None:
        rows = [r * slots + s for r, s in (slot_of_frame(t, world) for t in range(n_frames))]
        perm = _order_cache[key] = torch.tensor(rows, dtype=torch.long, device=device)
    return perm
